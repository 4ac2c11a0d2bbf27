for k in 3 6 10 14; do RS_B200_PACK_THREADS=$k timeout 300 python scripts/api_probe.py 2>&1 | tail -1; done
