set -x
RS_B200_HOST_PROFILE=1 timeout 300 python scripts/ncu_step.py 256 4 2>&1 | grep -E "host ms|total_ms" | tail -6
