set -x
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python scripts/bench_configs.py 2>&1 | grep "^{" | grep '"config": "4' | cut -c1-700
bash scripts/launch_list.sh cur | grep -E "ivec|cmvn"
