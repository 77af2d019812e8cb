#!/bin/bash
# round-2 GPU pass 2: fused up-conv parity, network/fullsize suites, bench, per-layer table
mkdir -p gpurun_out
for f in test_gpu_upfuse test_gpu_live_step test_gpu_network test_gpu_fullsize test_gpu_baseline_shapes; do
  timeout 900 python -m pytest tests/$f.py -q -s -m gpu --timeout 600 -x > gpurun_out/$f.log 2>&1
  echo "$f rc=$?"; grep -E "passed|failed" gpurun_out/$f.log | tail -1; grep -E "^FAILED|^ERROR|^E  " gpurun_out/$f.log | head -30
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2b.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['burst']['value'], d['roofline']['frac'], d['roofline']['wgrad_kernel'])
print('infer', d['inference_704']['value'], d['inference_704_tiled']['value'], 'adapter', d['adapter_finetune']['value'], d['adapter_finetune']['ms_per_step'])
PY
tail -3 gpurun_out/bench_r2b.err
python scripts/layer_times.py > gpurun_out/layers_r2b.log 2>&1; head -60 gpurun_out/layers_r2b.log | cut -c1-70; tail -1 gpurun_out/layers_r2b.log
