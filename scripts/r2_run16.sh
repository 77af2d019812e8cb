#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_upfuse.py tests/test_gpu_network.py tests/test_gpu_kernels.py -q -m gpu --timeout 600 > gpurun_out/t16.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/t16.log | cut -c1-300
for mt in 256 0 1000; do
  N2N_UPFUSE_TRAIN_MIN_TILES=$mt timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extra --no-inference > gpurun_out/bench_r2l_$mt.json 2> /dev/null
done
python - <<'PY'
import json
for f in ('256','0','1000'):
    try:
        d=json.load(open(f'gpurun_out/bench_r2l_{f}.json'))
        print('min_tiles',f, {k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['burst']['value'], d['roofline']['frac'], d['clocks']['sm_mhz'])
    except Exception as e: print(f, 'failed', e)
PY
python scripts/infer_batch_sweep.py 2>&1 | tail -4
python scripts/adapter_bench.py 2>&1 | tail -1
