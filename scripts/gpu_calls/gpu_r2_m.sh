#!/bin/bash
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 900 python -m pytest tests/test_graphs_gpu.py -m gpu -q -rf -s --no-header -p no:cacheprovider > gpurun_out/r2m_test_graphs.log 2>&1
echo "== graph tests rc=$?"; grep -E "lr=0|train |passed|failed|Error" gpurun_out/r2m_test_graphs.log | cut -c1-600
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -s 7400 -c 6700 --csv --log-file gpurun_out/r2m_launches.csv \
    python bench.py --steps 11 --warmup 11 --no-cpu-baseline --graphs 0 > gpurun_out/r2m_bench_under_ncu.log 2>&1
echo "== ncu launches rc=$?"; wc -l gpurun_out/r2m_launches.csv
