#!/bin/bash
# full single-GPU validation: every -m gpu test file in its own process, smoke, then the default bench
mkdir -p gpurun_out
rc=0
for f in tests/test_gpu_*.py; do
  b=$(basename $f .py)
  timeout 1200 python -m pytest $f -q -m gpu --timeout 900 > gpurun_out/$b.log 2>&1
  r=$?; echo "$b rc=$r $(grep -E 'passed|failed|skipped' gpurun_out/$b.log | tail -1)"; [ $r -ne 0 ] && rc=1
  grep -E "^FAILED|^ERROR" gpurun_out/$b.log | head -10
done
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_r2n.json 2> gpurun_out/bench_r2n.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2n.json'))
print({k:d[k] for k in ('value','ms_per_step','steps')}, d['e2e']['value'], d['burst']['value'], d['roofline']['frac'], d['roofline']['frac_burst'], d['clocks'])
print('infer', d['inference_704']['value'], d['inference_704_tiled']['value'], 'adapter', d['adapter_finetune']['ms_per_step'])
PY
exit $rc
