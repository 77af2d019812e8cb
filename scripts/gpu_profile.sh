#!/bin/bash
# ncu evidence for profiles/: (1) per-launch durations of the bench command, (2) full capture of the
# dominant kernel (dec_conv1a / dec_conv1b on the slab engine at the bench shape).
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-inference"
$CMD > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 420 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain_bench2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:slabgemm_umma -s 24 -c 2 -o gpurun_out/prof_slab $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"; tail -2 gpurun_out/ncu_full.log
