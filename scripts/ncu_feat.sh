set -x
ncu --set full --clock-control none --import-source on -k regex:"ubm_post|mfcc_kernel" -s 2 -c 2 -o gpurun_out/r2_feat python scripts/ncu_step.py 256 2 > gpurun_out/ncu_feat.log 2>&1
tail -2 gpurun_out/ncu_feat.log
