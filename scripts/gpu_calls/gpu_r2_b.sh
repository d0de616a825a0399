#!/bin/bash
# round-2 call B: tcgen05 attention forward (parity + timing), optimizer kernel tests
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -rf --no-header -p no:cacheprovider -k "tcgen05 or flash" > gpurun_out/r2b_test_attn.log 2>&1
echo "== attn tests rc=$?"; tail -n 40 gpurun_out/r2b_test_attn.log
timeout 600 python -m pytest tests/test_optim_gpu.py -m gpu -q -x -rf -s --no-header -p no:cacheprovider > gpurun_out/r2b_test_optim.log 2>&1
echo "== optim tests rc=$?"; tail -n 30 gpurun_out/r2b_test_optim.log
BB_ATTN_TC=1 timeout 300 python scripts/bench_attn.py > gpurun_out/r2b_bench_attn_tc.log 2>&1; cat gpurun_out/r2b_bench_attn_tc.log
