timeout 170 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
RS_B200_OVERLAP_STAGING=0 timeout 60 ncu --set full --clock-control none --import-source on -k regex:"mfcc_kernel|ubm_post|splice_lda4|decode_small" -s 4 -c 4 -o gpurun_out/r2g_others python scripts/ncu_step.py 256 2 > /dev/null 2>&1
ls -la gpurun_out/r2g_others.ncu-rep
