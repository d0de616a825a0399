#!/bin/bash
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
ATTN_ONLY="bev self 441 p=0" timeout 600 ncu --set full --clock-control none --import-source on -k regex:flash -o gpurun_out/ncu_flash441 -f python scripts/bench_attn.py > gpurun_out/ncu_flash441.log 2>&1
echo "== ncu rc=$?"; tail -5 gpurun_out/ncu_flash441.log
