#!/bin/bash
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -rf --no-header -p no:cacheprovider -k "layernorm or colsum" > gpurun_out/r2n_test_rows.log 2>&1
echo "== row tests rc=$?"; tail -n 4 gpurun_out/r2n_test_rows.log
timeout 300 python scripts/bench_hbm.py > gpurun_out/r2n_bench_hbm.log 2>&1; cat gpurun_out/r2n_bench_hbm.log
BEVBERT_BENCH_VERBOSE=1 timeout 900 python bench.py --steps 22 --warmup 11 --no-cpu-baseline > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err; tail -3 gpurun_out/r2n_bench.err; cut -c1-1400 gpurun_out/r2n_bench.json
timeout 900 python bench.py --config rxr --steps 11 --warmup 5 --no-cpu-baseline > gpurun_out/r2n_bench_rxr.json 2> gpurun_out/r2n_bench_rxr.err; tail -5 gpurun_out/r2n_bench_rxr.err; cut -c1-900 gpurun_out/r2n_bench_rxr.json
timeout 900 python bench.py --config reverie --steps 12 --warmup 4 --no-cpu-baseline > gpurun_out/r2n_bench_rvr.json 2> gpurun_out/r2n_bench_rvr.err; tail -5 gpurun_out/r2n_bench_rvr.err; cut -c1-900 gpurun_out/r2n_bench_rvr.json
