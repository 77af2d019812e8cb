#!/bin/bash
mkdir -p gpurun_out
python scripts/debug_nan.py > gpurun_out/debug_nan.log 2>&1; cat gpurun_out/debug_nan.log | tail -50
for f in test_gpu_upfuse test_gpu_live_step test_gpu_entry; do
  timeout 900 python -m pytest tests/$f.py -q -s -m gpu --timeout 600 > gpurun_out/$f.log 2>&1
  echo "$f rc=$?"; grep -E "passed|failed" gpurun_out/$f.log | tail -1; grep -E "^FAILED|^ERROR|^E  " gpurun_out/$f.log | head -30 | cut -c1-400
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2c.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['burst']['value'], d['roofline']['frac'], d['roofline']['wgrad_kernel'])
print('infer', d['inference_704']['value'], d['inference_704_tiled']['value'])
PY
tail -3 gpurun_out/bench_r2c.err
