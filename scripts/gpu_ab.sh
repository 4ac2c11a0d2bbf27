set -x
for v in v2 v3; do
RS_B200_TC=$v timeout 300 python scripts/ncu_step.py 256 3 2>&1 | tail -1
RS_B200_TC=$v bash scripts/launch_list.sh $v | grep gemm_tc
done
RS_B200_TC=v2 timeout 600 python -m pytest tests/test_gpu_zamia.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
