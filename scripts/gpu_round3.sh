#!/bin/bash
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -rf --no-header -p no:cacheprovider -k "native or linear_helpers or layernorm" > gpurun_out/test_kernels_sel.log 2>&1
echo "== kernel sel rc=$?"; tail -n 15 gpurun_out/test_kernels_sel.log | cut -c1-200
timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -q -rf -s --no-header -p no:cacheprovider > gpurun_out/test_model_gpu.log 2>&1
echo "== model tests rc=$?"; grep -E "task=|passed|failed" gpurun_out/test_model_gpu.log | cut -c1-220 | tail -n 30
timeout 600 python scripts/prof_host.py > gpurun_out/prof_host.log 2>&1; echo "== prof rc=$?"; head -n 45 gpurun_out/prof_host.log | cut -c1-160
BEVBERT_BENCH_VERBOSE=1 timeout 900 python bench.py --steps 22 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "== bench rc=$?"; grep gemm-shape gpurun_out/bench.err; tail -n 3 gpurun_out/bench.err; cat gpurun_out/bench.json
