#!/bin/bash
# ncu launch list of the bench command (per-launch gpu__time_duration, clocks untouched); one 11-step task cycle after warm-up
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -s 7000 -c 7000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 11 --warmup 11 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
echo "== ncu launches rc=$?"; wc -l gpurun_out/launches.csv
