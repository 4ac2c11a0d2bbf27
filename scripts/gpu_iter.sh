timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python scripts/bench_config5.py 2048 2>&1 | tail -1 | cut -c1-400
timeout 600 python bench.py 2>/dev/null | tail -1 > gpurun_out/bench_cur.json
python - <<PY
import json
x=json.loads(open('gpurun_out/bench_cur.json').read())
print('value',round(x['value']),'ms',round(x['ms_per_step'],3),'e2e',round(x['e2e']['value']),round(x['e2e']['ms_per_step'],3),'pageable',round(x['e2e_pageable']['value']),'2fl',round(x['e2e_two_in_flight']['value']),'api',round(x['e2e_api']['value']),'stages',{k:round(v,3) for k,v in x.get('stages_ms').items()}, 'parity', x.get('parity'), 'cpu', x.get('cpu_baseline',{}).get('value'))
PY
