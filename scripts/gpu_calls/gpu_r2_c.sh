#!/bin/bash
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
ATTN_TRACE=1 timeout 300 python scripts/bench_attn.py > gpurun_out/r2c_bench_attn_trace.log 2>&1; cat gpurun_out/r2c_bench_attn_trace.log
ATTN_ONLY="bev self 441" timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_tc_fwd -c 1 -f -o gpurun_out/r2c_ncu_attn_tc_fwd_441 python scripts/bench_attn.py > gpurun_out/r2c_ncu1.log 2>&1; tail -2 gpurun_out/r2c_ncu1.log
