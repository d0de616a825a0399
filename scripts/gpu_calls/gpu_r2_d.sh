#!/bin/bash
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -rf --no-header -p no:cacheprovider -k "tcgen05 or flash or project_bev" > gpurun_out/r2d_test_attn.log 2>&1
echo "== attn tests rc=$?"; tail -n 40 gpurun_out/r2d_test_attn.log
timeout 300 python scripts/bench_attn.py > gpurun_out/r2d_bench_attn.log 2>&1; cat gpurun_out/r2d_bench_attn.log
