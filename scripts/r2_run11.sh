#!/bin/bash
mkdir -p gpurun_out
for f in test_gpu_upfuse test_gpu_network test_gpu_fullsize; do
  timeout 900 python -m pytest tests/$f.py -q -s -m gpu --timeout 600 > gpurun_out/$f.log 2>&1
  echo "$f rc=$?"; grep -E "passed|failed" gpurun_out/$f.log | tail -1; grep -E "^FAILED|^ERROR|^E  |composite backward" gpurun_out/$f.log | head -30 | cut -c1-300
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra --no-inference > gpurun_out/bench_r2i.json 2> gpurun_out/bench_r2i.err; echo "bench rc=$?"
N2N_NO_UPFUSE_TRAIN=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra --no-inference > gpurun_out/bench_r2i_nofuse.json 2> gpurun_out/bench_r2i_nofuse.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra --no-inference > gpurun_out/bench_r2i_b.json 2> /dev/null
python - <<'PY'
import json
for f in ('bench_r2i','bench_r2i_nofuse','bench_r2i_b'):
    try:
        d=json.load(open(f'gpurun_out/{f}.json'))
        print(f, {k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['burst']['value'], d['roofline']['frac'], d['roofline']['wgrad_kernel']['ms_per_step'], d['final_loss'], d['clocks']['sm_mhz'])
    except Exception as e: print(f, 'failed', e)
PY
bash scripts/r2_run10.sh 2>&1 | tail -24
