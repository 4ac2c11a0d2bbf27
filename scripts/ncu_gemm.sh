set -x
ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 8 -c 4 -o gpurun_out/r2_gemm2 python scripts/ncu_step.py 256 1 > gpurun_out/ncu_gemm2.log 2>&1
tail -3 gpurun_out/ncu_gemm2.log
