for m in 32 64 96 128; do echo "== profile mode $m"
RS_B200_TC_PROFILE=$m bash scripts/launch_list.sh e$m | grep gemm_tc
done
