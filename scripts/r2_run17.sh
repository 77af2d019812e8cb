#!/bin/bash
mkdir -p gpurun_out
for ov in 0 1 0 1; do
  N2N_OVERLAP=$ov timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extra --no-inference > gpurun_out/bench_r2m_$ov.json 2> gpurun_out/bench_r2m_$ov.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_r2m_$ov.json')); print('overlap $ov', {k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['burst']['value'], d['clocks']['sm_mhz'], d['final_loss'])
except Exception as e: print('overlap $ov failed', e)
PY
done
tail -3 gpurun_out/bench_r2m_1.err
