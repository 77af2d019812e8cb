#!/bin/bash
mkdir -p gpurun_out
python scripts/adapter_bench.py > gpurun_out/plain_adapter.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 160 --csv --log-file gpurun_out/launches_adapter.csv python scripts/adapter_bench.py > gpurun_out/ncu_adapter.log 2>&1
echo "adapter launch list rc=$?"; tail -1 gpurun_out/plain_adapter.log
python - <<'PY'
import csv, collections
rows=list(csv.reader(open('gpurun_out/launches_adapter.csv')))
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
for k,v in sorted(agg.items(), key=lambda kv:-kv[1])[:22]: print(f"{k:62s} {n[k]:4d} launches {v/1e3:9.1f} us  {100*v/tot:5.1f}%")
print('total us', tot/1e3)
PY
python scripts/subsample_ncu.py > gpurun_out/plain_subsample.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:subsample_vec -s 12 -c 4 -o gpurun_out/prof_r2_subsample python scripts/subsample_ncu.py > gpurun_out/ncu_r2_subsample.log 2>&1
echo "subsample ncu rc=$?"
