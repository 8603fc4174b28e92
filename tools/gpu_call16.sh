#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout -s KILL 230 ncu --metrics gpu__time_duration.sum,launch__grid_size --clock-control none --csv --log-file gpurun_out/c16_launches.csv python tools/sched_probe.py gen 128 2 2 1000 1 0.05 > gpurun_out/c16_probe.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/c16_launches.csv')) if len(r)>10]
h=rows[0]; ki=h.index('Kernel Name'); mi=h.index('Metric Name'); vi=h.index('Metric Value'); ii=h.index('ID')
d=collections.defaultdict(dict)
for r in rows[1:]:
    d[r[ii]]['k']=r[ki].split('(')[0].replace('void ','').replace('sacb::','').replace('<unnamed>::','').replace('unnamed>::','')[:60]
    d[r[ii]][r[mi]]=float(r[vi].replace(',',''))
agg=collections.defaultdict(lambda:[0,0.0,0.0,0])
for v in d.values():
    a=agg[v['k']]; a[0]+=1; a[1]+=v.get('gpu__time_duration.sum',0)/1e6; a[2]+=v.get('gpu__time_duration.sum',0)/1e6*v.get('launch__grid_size',0); a[3]+=int(v.get('launch__grid_size',0))
print(f"{'kernel':60s} {'n':>5s} {'sum ms':>10s} {'sum ms*grid':>14s} {'CTAs':>8s}")
for k,a in sorted(agg.items(), key=lambda x:-x[1][2]): print(f"{k:60s} {a[0]:5d} {a[1]:10.1f} {a[2]:14.0f} {a[3]:8d}")
PY
tail -2 gpurun_out/c16_probe.log | cut -c1-600
