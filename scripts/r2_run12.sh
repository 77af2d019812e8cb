#!/bin/bash
mkdir -p gpurun_out
for f in test_gpu_upfuse test_gpu_network test_gpu_fullsize test_gpu_kernels; do
  timeout 900 python -m pytest tests/$f.py -q -s -m gpu --timeout 600 > gpurun_out/$f.log 2>&1
  echo "$f rc=$?"; grep -E "passed|failed" gpurun_out/$f.log | tail -1; grep -E "^FAILED|^ERROR|^E  " gpurun_out/$f.log | head -30 | cut -c1-300
done
python scripts/layer_times.py > gpurun_out/layers_r2j.log 2>&1; sed -n 1,23p gpurun_out/layers_r2j.log | cut -c1-70; tail -1 gpurun_out/layers_r2j.log
N2N_NO_DUAL_ISSUE=1 python scripts/layer_times.py > gpurun_out/layers_r2j_nodual.log 2>&1; sed -n 7,22p gpurun_out/layers_r2j_nodual.log | cut -c1-70; tail -1 gpurun_out/layers_r2j_nodual.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/bench_r2j.json 2> gpurun_out/bench_r2j.err; echo "bench rc=$?"
N2N_NO_DUAL_ISSUE=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra --no-inference > gpurun_out/bench_r2j_nodual.json 2> /dev/null
python - <<'PY'
import json
for f in ('bench_r2j','bench_r2j_nodual'):
    try:
        d=json.load(open(f'gpurun_out/{f}.json'))
        print(f, {k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['burst']['value'], d['roofline']['frac'], d['final_loss'], d['clocks']['sm_mhz'])
        if d.get('inference_704'): print('infer', d['inference_704']['value'], d['inference_704_tiled']['value'], d['inference_704_tiled']['psnr_first'], d['inference_704']['psnr_first'])
    except Exception as e: print(f, 'failed', e)
PY
tail -3 gpurun_out/bench_r2j.err
