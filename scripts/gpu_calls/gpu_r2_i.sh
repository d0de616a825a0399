#!/bin/bash
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
ATTN_ONLY="bev self 441" timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_tc_fwd2 -c 1 -f -o gpurun_out/r2i_ncu_fwd2_441 python scripts/bench_attn.py > gpurun_out/r2i_ncu1.log 2>&1; tail -2 gpurun_out/r2i_ncu1.log
