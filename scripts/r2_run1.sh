#!/bin/bash
# round-2 first GPU pass: new tests first, then the whole suite, then bench + per-layer timing
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
for f in test_gpu_baseline_shapes test_gpu_multirank; do
  timeout 1500 python -m pytest tests/$f.py -q -s -m gpu --timeout 900 > gpurun_out/$f.log 2>&1
  echo "$f rc=$?"; grep -E "passed|failed" gpurun_out/$f.log | tail -1; grep -E "^FAILED|^ERROR|^C1|^E  " gpurun_out/$f.log | head -40
done
RUN_BENCH=0 bash scripts/gpu_tests.sh
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; echo "bench rc=$?"
cat gpurun_out/bench_r2a.json; tail -5 gpurun_out/bench_r2a.err
python scripts/layer_times.py > gpurun_out/layers_r2a.log 2>&1; tail -1 gpurun_out/layers_r2a.log
