# End-of-round measurement set (one GPU): tests, bench line, reference arm, configs 3 / 4 / 5, n-best, launch list with
# DRAM bytes, full ncu captures of the hot kernels.  Outputs under gpurun_out/ (copied to profiles/ by hand).
set -x
T=${1:-r2f}
python -m pytest tests -m gpu -q 2>&1 | tail -3
python bench.py 2>&1 | tail -1 > gpurun_out/${T}_bench_full.json
python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 > gpurun_out/${T}_bench_reference_arm.json
python scripts/bench_configs.py 2>&1 | grep "^{" > gpurun_out/${T}_configs_3_4.jsonl
python scripts/bench_configs.py --surface 2>&1 | grep "^{" > gpurun_out/${T}_config4_python_surface.jsonl
python scripts/bench_configs.py --nbest 2>&1 | grep "^{" > gpurun_out/${T}_nbest_batch256.jsonl
python scripts/bench_config5.py 2048 2>&1 | grep "^{" > gpurun_out/${T}_config5_pool_1gpu.jsonl
RS_B200_OVERLAP_STAGING=0 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 42 -c 44 --csv --log-file gpurun_out/${T}_launches.csv python scripts/ncu_step.py 256 2 > /dev/null 2>&1
RS_B200_OVERLAP_STAGING=0 ncu --set full --clock-control none --import-source on -k regex:gemm_tc3 -s 8 -c 4 -o gpurun_out/${T}_gemm python scripts/ncu_step.py 256 1 > /dev/null 2>&1
RS_B200_OVERLAP_STAGING=0 ncu --set full --clock-control none --import-source on -k regex:"mfcc_kernel|ubm_post|splice_lda4|decode_small" -s 4 -c 4 -o gpurun_out/${T}_others python scripts/ncu_step.py 256 2 > /dev/null 2>&1
for f in gpurun_out/${T}_bench_full.json gpurun_out/${T}_bench_reference_arm.json gpurun_out/${T}_configs_3_4.jsonl gpurun_out/${T}_config4_python_surface.jsonl gpurun_out/${T}_nbest_batch256.jsonl gpurun_out/${T}_config5_pool_1gpu.jsonl; do echo "== $f"; cut -c1-700 $f; done
