#!/bin/bash
# ncu evidence for profiles/: (1) per-launch durations of ~2 bench steps, (2) one --set full capture
# of the tap-GEMM kernel on the three biggest full-resolution layers (dec_conv1a, dec_conv1b, nin_a).
# Numbers printed by runs under ncu are never bench values.
mkdir -p gpurun_out
TAG=${TAG:-r01}
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_${TAG}.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s ${SKIP:-640} -c ${COUNT:-420} --csv \
    --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_launches_${TAG}.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2_${TAG}.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:${KERNEL:-tapgemm_umma} -s ${KSKIP:-35} -c 3 \
    -f -o gpurun_out/prof_${TAG} $CMD > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out/
