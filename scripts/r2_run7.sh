#!/bin/bash
mkdir -p gpurun_out
for f in test_gpu_upfuse test_gpu_network; do
  timeout 900 python -m pytest tests/$f.py -q -s -m gpu --timeout 600 > gpurun_out/$f.log 2>&1
  echo "$f rc=$?"; grep -E "passed|failed" gpurun_out/$f.log | tail -1; grep -E "^FAILED|^ERROR|^E  " gpurun_out/$f.log | head -20 | cut -c1-300
done
python scripts/layer_times.py > gpurun_out/layers_r2e.log 2>&1; head -22 gpurun_out/layers_r2e.log | cut -c1-70; tail -1 gpurun_out/layers_r2e.log
N2N_NO_DUAL_ISSUE=1 python scripts/layer_times.py > gpurun_out/layers_r2e_nodual.log 2>&1; sed -n 7,22p gpurun_out/layers_r2e_nodual.log | cut -c1-70
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/bench_r2e.json 2> gpurun_out/bench_r2e.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2e.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['burst']['value'], d['roofline']['frac'], d['roofline']['wgrad_kernel'])
print('infer', d['inference_704']['value'], d['inference_704_tiled']['value'])
PY
