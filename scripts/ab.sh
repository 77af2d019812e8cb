#!/bin/bash
# interleaved A/B of environment toggles: scripts/ab.sh "NAME1=V1" "NAME2=V2 NAME3=V3" ...  (3 rounds, 40 timed steps each)
mkdir -p gpurun_out
for round in 1 2 3; do
  i=0
  for cfg in "$@"; do
    i=$((i+1))
    env $cfg timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-extra --no-inference --soak-seconds 2 > gpurun_out/ab_${i}_${round}.json 2> /dev/null
  done
done
python - "$@" <<'PY'
import json, sys, statistics
cfgs=sys.argv[1:]
for i,c in enumerate(cfgs,1):
    ms=[]; clk=[]
    for r in (1,2,3):
        try:
            d=json.load(open(f'gpurun_out/ab_{i}_{r}.json')); ms.append(d['ms_per_step']); clk.append(d['clocks']['sm_mhz'])
        except Exception as e: pass
    print(f"{c:45s} ms/step {['%.3f'%m for m in ms]} median {statistics.median(ms):.3f} min {min(ms):.3f} clocks {clk}")
PY
