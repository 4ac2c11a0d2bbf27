set -x
timeout 300 python scripts/ncu_step.py 256 3 2>&1 | tail -1
RS_B200_DECODE_PROFILE=1 timeout 300 python scripts/ncu_step.py 256 1 2>&1 | grep "decode_small phases" | head -2
timeout 900 python -m pytest tests/test_gpu_zamia.py tests/test_gpu_parity.py tests/test_gpu_nbest.py -m gpu -x -q 2>&1 | tail -3
