# per-launch device times of one step (batch 256), after one warm-up step
ncu --metrics gpu__time_duration.sum --clock-control none -s 42 -c 44 --csv --log-file gpurun_out/launches_${1:-x}.csv python scripts/ncu_step.py 256 2 > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_${1:-x}.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
tot={}
for r in rows[1:]:
    try: v=float(r[vi].replace(',',''))
    except: continue
    k=r[ki][:60]; tot.setdefault(k,[]).append(v/1000 if v>5000 else v)
for k,v in tot.items(): print('%-62s n=%2d sum %8.1f us  each %s'%(k,len(v),sum(v),' '.join('%.0f'%x for x in v)))
PY
