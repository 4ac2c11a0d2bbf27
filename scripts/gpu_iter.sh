timeout 300 python scripts/overlap_probe.py 2>&1 | grep -E "^pinned"
RS_B200_HOST_PROFILE=1 timeout 300 python scripts/overlap_probe.py 2>&1 | grep -E "staging marks" | head -14 | tail -2
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
