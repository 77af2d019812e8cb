#!/bin/bash
mkdir -p gpurun_out
for f in test_gpu_upfuse test_gpu_network test_gpu_fullsize test_gpu_baseline_shapes; do
  timeout 900 python -m pytest tests/$f.py -q -s -m gpu --timeout 600 > gpurun_out/$f.log 2>&1
  echo "$f rc=$?"; grep -E "passed|failed" gpurun_out/$f.log | tail -1; grep -E "^FAILED|^ERROR|^E  |composite backward" gpurun_out/$f.log | head -30 | cut -c1-300
done
python scripts/layer_times.py > gpurun_out/layers_r2g.log 2>&1; sed -n 23,200p gpurun_out/layers_r2g.log | cut -c1-70; tail -1 gpurun_out/layers_r2g.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra --no-inference > gpurun_out/bench_r2g.json 2> gpurun_out/bench_r2g.err; echo "bench rc=$?"
N2N_NO_UPFUSE_TRAIN=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra --no-inference > gpurun_out/bench_r2g_nofuse.json 2> gpurun_out/bench_r2g_nofuse.err
python - <<'PY'
import json
for f in ('bench_r2g','bench_r2g_nofuse'):
    try:
        d=json.load(open(f'gpurun_out/{f}.json'))
        print(f, {k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['burst']['value'], d['roofline']['frac'], d['roofline']['wgrad_kernel']['ms_per_step'], d['final_loss'])
    except Exception as e: print(f, 'failed', e)
PY
tail -3 gpurun_out/bench_r2g.err
