#!/bin/bash
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "flash or native" > gpurun_out/test_flash.log 2>&1; echo "== flash rc=$?"; tail -5 gpurun_out/test_flash.log
timeout 300 python scripts/bench_attn.py 2>&1 | tee gpurun_out/bench_attn.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_flash2.json 2> gpurun_out/bench_flash2.err; echo "== bench rc=$?"; cut -c1-300 gpurun_out/bench_flash2.json
