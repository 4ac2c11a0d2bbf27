set -x
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py 2>&1 | tail -1 > gpurun_out/bench_cur.json
cut -c1-1500 gpurun_out/bench_cur.json
