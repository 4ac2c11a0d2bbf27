set -x
python -m pytest tests -m gpu -q 2>&1 | tail -3
python bench.py 2>&1 | tail -1 > gpurun_out/r2_bench_b_full.json
python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 > gpurun_out/r2_bench_b_reference_arm.json
python scripts/bench_configs.py 2>&1 | grep "^{" > gpurun_out/r2_configs_3_4.jsonl
python scripts/bench_configs.py --surface 2>&1 | grep "^{" > gpurun_out/r2_config4_python_surface.jsonl
python scripts/bench_configs.py --nbest 2>&1 | grep "^{" > gpurun_out/r2_nbest_batch256.jsonl
python scripts/bench_config5.py 2048 2>&1 | grep "^{" > gpurun_out/r2_config5_pool_1gpu.jsonl
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 42 -c 44 --csv --log-file gpurun_out/r2_launches_c.csv python scripts/ncu_step.py 256 2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:decode_small -s 1 -c 1 -o gpurun_out/r2_decode_small python scripts/ncu_step.py 256 2 > /dev/null 2>&1
for f in gpurun_out/r2_bench_b_full.json gpurun_out/r2_bench_b_reference_arm.json gpurun_out/r2_configs_3_4.jsonl gpurun_out/r2_config4_python_surface.jsonl gpurun_out/r2_nbest_batch256.jsonl gpurun_out/r2_config5_pool_1gpu.jsonl; do echo "== $f"; cut -c1-900 $f; done
