#!/bin/bash
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -rf --no-header -p no:cacheprovider -k "tcgen05 or flash" > gpurun_out/r2e_test_attn.log 2>&1
echo "== attn tests rc=$?"; tail -n 5 gpurun_out/r2e_test_attn.log
ATTN_TRACE=1 timeout 300 python scripts/bench_attn.py > gpurun_out/r2e_bench_attn.log 2>&1; cat gpurun_out/r2e_bench_attn.log
ATTN_ONLY="bev self 441" timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_tc_bwd -c 2 -f -o gpurun_out/r2e_ncu_attn_tc_bwd_441 python scripts/bench_attn.py > gpurun_out/r2e_ncu1.log 2>&1; tail -2 gpurun_out/r2e_ncu1.log
