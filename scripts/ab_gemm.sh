# A/B of the two tensor-core kernels on a GPU box: layer probe (error + time), stage timings, parity tests
set -x
mkdir -p gpurun_out
RS_B200_TC=v1 timeout 300 python scripts/gemm_probe.py 32768 > gpurun_out/probe_v1.jsonl 2>&1
timeout 300 python scripts/gemm_probe.py 32768 > gpurun_out/probe_v2.jsonl 2>&1
python - <<'PY'
import json
def rd(f):
    out=[]
    for l in open(f):
        if l.startswith('{'): out.append(json.loads(l))
    return out
a,b=rd('gpurun_out/probe_v1.jsonl'),rd('gpurun_out/probe_v2.jsonl')
print("v1 lines",len(a),"v2 lines",len(b))
for x,y in zip(a,b):
    print(x['k'],x['n'],x['offsets'],x['stride'],'v1 ms %.3f max %.2e rms %.2e | v2 ms %.3f max %.2e rms %.2e | simt max %.2e'%(x['tc_split']['ms'],x['tc_split']['max_abs'],x['tc_split']['rms'],y['tc_split']['ms'],y['tc_split']['max_abs'],y['tc_split']['rms'],y['simt']['max_abs']))
PY
tail -5 gpurun_out/probe_v2.jsonl | cut -c1-400
RS_B200_TC=v1 timeout 300 python scripts/ncu_step.py 256 3 2>&1 | tail -2
timeout 300 python scripts/ncu_step.py 256 3 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
