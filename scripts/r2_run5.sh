#!/bin/bash
mkdir -p gpurun_out
for f in test_gpu_upfuse test_gpu_network; do
  timeout 900 python -m pytest tests/$f.py -q -s -m gpu --timeout 600 > gpurun_out/$f.log 2>&1
  echo "$f rc=$?"; grep -E "passed|failed" gpurun_out/$f.log | tail -1; grep -E "^FAILED|^ERROR|^E  " gpurun_out/$f.log | head -20 | cut -c1-300
done
export B=64 REPS=2
python scripts/fwd_only.py > gpurun_out/plain_fwd.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:slabgemm_umma -s 38 -c 3 -o gpurun_out/prof_r2_upconv python scripts/fwd_only.py > gpurun_out/ncu_r2_upconv.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_r2_upconv.log
