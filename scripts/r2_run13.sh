#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_network.py tests/test_gpu_baseline_shapes.py tests/test_gpu_entry.py -q -m gpu --timeout 600 -x > gpurun_out/t13.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/t13.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2k.json 2> gpurun_out/bench_r2k.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2k.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['burst']['value'], d['roofline']['frac'], d['final_loss'], d['clocks'])
print('infer', d['inference_704']['value'], d['inference_704_tiled']['value'], d['inference_704_tiled']['psnr_first'], d['inference_704']['psnr_first'])
print('adapter', d['adapter_finetune']); print('torch', d['torch_gpu_baseline']); print('cpu', d['cpu_baseline'])
for k,v in d['hbm_kernels'].items(): print(k, round(v['us_per_launch'],2), round(v['frac'],3))
PY
tail -3 gpurun_out/bench_r2k.err
