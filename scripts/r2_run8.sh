#!/bin/bash
mkdir -p gpurun_out
for f in test_gpu_upfuse test_gpu_network test_gpu_kernels test_gpu_fullsize; do
  timeout 900 python -m pytest tests/$f.py -q -s -m gpu --timeout 600 > gpurun_out/$f.log 2>&1
  echo "$f rc=$?"; grep -E "passed|failed" gpurun_out/$f.log | tail -1; grep -E "^FAILED|^ERROR|^E  " gpurun_out/$f.log | head -20 | cut -c1-300
done
python scripts/layer_times.py > gpurun_out/layers_r2f.log 2>&1; cat gpurun_out/layers_r2f.log | cut -c1-70 | head -100; tail -1 gpurun_out/layers_r2f.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/bench_r2f.json 2> gpurun_out/bench_r2f.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2f.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['burst']['value'], d['roofline']['frac'], d['roofline']['wgrad_kernel'])
print('infer', d['inference_704']['value'], d['inference_704_tiled']['value'])
PY
