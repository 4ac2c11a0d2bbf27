# quick GPU iteration: stage timings, role profile of the tensor-core kernel, zamia parity tests
set -x
timeout 300 python scripts/ncu_step.py 256 3 2>&1 | tail -2
RS_B200_TC_PROFILE=1 timeout 300 python scripts/ncu_step.py 256 1 2>&1 | grep "tc2 profile" | sed -n 1,60p
timeout 600 python -m pytest tests/test_gpu_zamia.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
