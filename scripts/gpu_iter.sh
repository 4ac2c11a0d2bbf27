timeout 60 python -m pytest tests/test_gpu_zamia.py tests/test_gpu_parity.py -m gpu -x -q -k "arpa" 2>&1 | tail -2
timeout 40 python scripts/config3_probe.py 2>&1 | tail -1 | tee gpurun_out/r2g_config3_strict_after.json
