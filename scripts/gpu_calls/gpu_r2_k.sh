#!/bin/bash
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 300 python scripts/bench_attn.py > gpurun_out/r2k_bench_attn.log 2>&1; head -4 gpurun_out/r2k_bench_attn.log
timeout 900 python bench.py --steps 22 --warmup 11 > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; tail -3 gpurun_out/r2k_bench.err; cat gpurun_out/r2k_bench.json
timeout 1800 python -m pytest tests -m gpu -q -x -rf --no-header -p no:cacheprovider > gpurun_out/r2k_test_all.log 2>&1
echo "== all gpu tests rc=$?"; tail -n 25 gpurun_out/r2k_test_all.log
