RUN_BENCH=0 bash scripts/gpu_tests.sh
python scripts/component_times.py 2>&1 | tail -8
