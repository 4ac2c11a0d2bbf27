set -x
for fs in 2 4; do
echo "=== FOLD_SHORT=$fs"
RS_B200_TC_FOLD_SHORT=$fs timeout 300 python scripts/ncu_step.py 256 3 2>&1 | tail -1
RS_B200_TC_FOLD_SHORT=$fs timeout 300 python scripts/debug_ll.py 2>&1 | grep "^utt" | cut -c1-260
done
echo "=== FOLD=4"
RS_B200_TC_FOLD=4 timeout 300 python scripts/ncu_step.py 256 3 2>&1 | tail -1
RS_B200_TC_FOLD=4 timeout 300 python scripts/debug_ll.py 2>&1 | grep "^utt" | cut -c1-260
