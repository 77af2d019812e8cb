#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-inference --no-extra --soak-seconds 0"
$CMD > gpurun_out/plain_r2h.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 500 -c 480 --csv --log-file gpurun_out/launches_r2h.csv $CMD > gpurun_out/ncu_launches_r2h.log 2>&1
echo "launch list rc=$?"; tail -2 gpurun_out/ncu_launches_r2h.log
python - <<'PY'
import csv, collections
rows=list(csv.reader(open('gpurun_out/launches_r2h.csv')))
# find header
for i,r in enumerate(rows):
    if 'Kernel Name' in r: hdr=r; start=i+1; break
ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
agg=collections.OrderedDict(); n=collections.Counter()
for r in rows[start:]:
    if len(r)<=vi: continue
    name=r[ki].split('(')[0][:60]
    try: v=float(r[vi].replace(',',''))
    except: continue
    agg[name]=agg.get(name,0)+v; n[name]+=1
tot=sum(agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-kv[1]): print(f"{k:62s} {n[k]:4d} launches {v/1e3:9.1f} us  {100*v/tot:5.1f}%")
print('total us', tot/1e3)
PY
