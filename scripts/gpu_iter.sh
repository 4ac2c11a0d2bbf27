timeout 120 python -m pytest tests/test_gpu_zamia.py -m gpu -x -q -s -k "arpa" 2>&1 | grep -E "ARPA|passed|failed|rror" | tail -5
