#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_r2_8gpu.json 2> gpurun_out/bench_r2_8gpu.err; echo "bench8 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_8gpu.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['value'], d['burst']['value'], d['dp_parity'], d['dp_parity_detail']['gradient'], d['strong_scaling'], d['clocks'])
print('infer', d['inference_704']['value'], d['inference_704_tiled']['value'])
PY
tail -4 gpurun_out/bench_r2_8gpu.err | cut -c1-300
