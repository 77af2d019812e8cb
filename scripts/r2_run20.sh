#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 600 -k "adam" > gpurun_out/t20.log 2>&1; echo "adam tests rc=$?"; tail -2 gpurun_out/t20.log
for mt in 2 64 148 2; do
  N2N_PAIR_MIN_TILES=$mt python scripts/layer_times.py > gpurun_out/layers_pm_$mt.log 2>&1; echo "min_tiles=$mt: $(tail -1 gpurun_out/layers_pm_$mt.log)"
  N2N_PAIR_MIN_TILES=$mt timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extra --no-inference > gpurun_out/bench_pm_$mt.json 2> /dev/null
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_pm_$mt.json')); print('  bench', {k:d[k] for k in ('value','ms_per_step')}, d['burst']['ms_per_step'], d['clocks']['sm_mhz'])
except Exception as e: print('failed', e)
PY
done
python - <<'PY'
import json,sys
sys.path.insert(0,'.')
import torch, bench
r=bench.hbm_kernels(torch.device('cuda:0'))
for k,v in r.items(): print(k, round(v['us_per_launch'],2), round(v['frac'],3))
PY
