for it in 4 6; do echo items $it; RS_B200_DIRECT_ITEMS=$it timeout 300 python scripts/overlap_probe.py 2>&1 | grep -E "^pinned|^pageable|rror"; done
RS_B200_HOST_PROFILE=1 RS_B200_DIRECT_ITEMS=4 timeout 300 python scripts/overlap_probe.py 2>&1 | grep -E "staging items" | head -16 | tail -4
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zamia.py tests/test_gpu_surface.py -m gpu -x -q 2>&1 | tail -3
