timeout 300 python scripts/ncu_step.py 256 4 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_zamia.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
