RUN_BENCH=0 bash scripts/gpu_tests.sh
python scripts/layer_times.py > gpurun_out/layers.log 2>&1; tail -1 gpurun_out/layers.log
python bench.py > gpurun_out/bench_dev.json 2> gpurun_out/bench_dev.err; tail -c 2500 gpurun_out/bench_dev.json; tail -3 gpurun_out/bench_dev.err
