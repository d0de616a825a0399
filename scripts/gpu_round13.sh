#!/bin/bash
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/test_gpu_all.log 2>&1; echo "== pytest gpu rc=$?"; tail -5 gpurun_out/test_gpu_all.log
timeout 300 python scripts/bench_attn.py 2>&1 | tee gpurun_out/bench_attn.log
timeout 300 python scripts/prof_phases.py 2>&1 | tail -40 | tee gpurun_out/prof_phases.log
