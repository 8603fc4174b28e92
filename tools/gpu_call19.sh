#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout -s KILL 200 ncu --metrics gpu__time_duration.sum,launch__grid_size --clock-control none --csv --log-file gpurun_out/c19_launches.csv python tools/grade_probe.py 20000 128 > gpurun_out/c19_probe.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/c19_launches.csv')) if len(r)>10]
h=rows[0]; ki=h.index('Kernel Name'); mi=h.index('Metric Name'); vi=h.index('Metric Value'); ii=h.index('ID')
d=collections.OrderedDict()
for r in rows[1:]:
    e=d.setdefault(r[ii],{}); e['k']=r[ki].split('(')[0].replace('void ','').replace('sacb::','').replace('<unnamed>::','').replace('unnamed>::','')[:48]
    e[r[mi]]=float(r[vi].replace(',',''))
for i,v in d.items(): print(i, v['k'], round(v.get('gpu__time_duration.sum',0)/1e6,1),'ms grid',int(v.get('launch__grid_size',0)))
PY
