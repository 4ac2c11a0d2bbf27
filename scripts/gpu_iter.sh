set -x
timeout 300 python scripts/ncu_step.py 256 3 2>&1 | tail -1
bash scripts/launch_list.sh cur | grep -v "gemm_tc3_kernel<(int)-1\|<801\|<0, 0>\|<1, 0>"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
