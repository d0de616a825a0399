#!/bin/bash
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 900 python -m pytest tests/test_graphs_gpu.py -m gpu -q -x -rf -s --no-header -p no:cacheprovider > gpurun_out/r2l_test_graphs.log 2>&1
echo "== graph tests rc=$?"; tail -n 40 gpurun_out/r2l_test_graphs.log
timeout 900 python bench.py --steps 22 --warmup 11 > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err; tail -5 gpurun_out/r2l_bench.err; cat gpurun_out/r2l_bench.json
timeout 900 python bench.py --steps 22 --warmup 11 --graphs 0 --no-cpu-baseline > gpurun_out/r2l_bench_eager.json 2> gpurun_out/r2l_bench_eager.err; tail -3 gpurun_out/r2l_bench_eager.err; cat gpurun_out/r2l_bench_eager.json
