for m in 1 3 5; do echo "== profile mode $m"
RS_B200_TC_PROFILE=$m timeout 300 python scripts/ncu_step.py 256 2 2>&1 | grep -E "tc2 profile|nnet_ms" | sed -n 1,12p
done
