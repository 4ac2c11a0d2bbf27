set -x
timeout 300 python scripts/ncu_step.py 256 3 2>&1 | tail -1
bash scripts/launch_list.sh cur | grep -E "ubm|mfcc|splice"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
