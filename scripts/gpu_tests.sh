#!/bin/bash
# Full GPU check on a B200 box: parity tests (one process per file so that a CUDA fault in one
# family cannot poison the others), smoke, and a short bench.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
rc=0
for f in test_gpu_umma_probe test_gpu_kernels test_gpu_network test_gpu_entry test_gpu_fullsize; do
  timeout 1200 python -m pytest tests/$f.py -q -s -m gpu --timeout 600 > gpurun_out/$f.log 2>&1
  r=$?; echo "$f rc=$r"; [ $r -ne 0 ] && rc=1
  grep -E "passed|failed" gpurun_out/$f.log | tail -1
  grep -E "^FAILED|^ERROR" gpurun_out/$f.log | head -20
done
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
if [ "${RUN_BENCH:-1}" = "1" ]; then
  timeout 900 python bench.py --steps ${BENCH_STEPS:-10} --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
  cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
fi
exit $rc
