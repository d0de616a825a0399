#!/bin/bash
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -rf --no-header -p no:cacheprovider -k "tcgen05 or flash" > gpurun_out/r2h_test_attn.log 2>&1
echo "== attn tests rc=$?"; tail -n 6 gpurun_out/r2h_test_attn.log
timeout 300 python scripts/bench_attn.py > gpurun_out/r2h_bench_attn.log 2>&1; cat gpurun_out/r2h_bench_attn.log
timeout 1200 python -m pytest tests/test_model_gpu.py -m gpu -q -rf -s --no-header -p no:cacheprovider -k "fp32_verification" > gpurun_out/r2h_test_fp32.log 2>&1
echo "== fp32 arm tests rc=$?"; grep -E "fp32 arm|passed|failed|Error|error" gpurun_out/r2h_test_fp32.log | tail -20
