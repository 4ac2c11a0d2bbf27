set -x
timeout 300 python scripts/gemm_probe.py 32768 > gpurun_out/probe_v3.jsonl 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/probe_v3.jsonl'):
    if l.startswith('{'):
        x=json.loads(l); print(x['k'],x['n'],x['offsets'],x['stride'],'tc ms %.3f max %.2e rms %.2e | simt max %.2e'%(x['tc_split']['ms'],x['tc_split']['max_abs'],x['tc_split']['rms'],x['simt']['max_abs']))
    elif 'rror' in l or 'Traceback' in l: print(l.strip())
PY
timeout 300 python scripts/ncu_step.py 256 3 2>&1 | tail -1
bash scripts/launch_list.sh cur | grep "gemm_tc"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
