#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multirank.py -q -s -m gpu --timeout 500 > gpurun_out/test_gpu_multirank.log 2>&1; echo "multirank rc=$?"; tail -3 gpurun_out/test_gpu_multirank.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_r2_2gpu.json 2> gpurun_out/bench_r2_2gpu.err; echo "bench2 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_2gpu.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['value'], d['dp_parity'], d['dp_parity_detail'], d['strong_scaling'])
print('infer', d['inference_704']['value'], d['inference_704_tiled']['value'])
PY
tail -5 gpurun_out/bench_r2_2gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 3 --warmup 1 --impl reference > gpurun_out/bench_r2_2gpu_ref.json 2> gpurun_out/bench_r2_2gpu_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_r2_2gpu_ref.json
