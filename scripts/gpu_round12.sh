#!/bin/bash
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "flash" > gpurun_out/test_flash.log 2>&1; echo "== flash rc=$?"; tail -30 gpurun_out/test_flash.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/test_gpu_all.log 2>&1; echo "== pytest gpu rc=$?"; tail -15 gpurun_out/test_gpu_all.log
BEVBERT_BENCH_VERBOSE=1 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_flash.json 2> gpurun_out/bench_flash.err; echo "== bench rc=$?"; cut -c1-300 gpurun_out/bench_flash.json
BB_FLASH=0 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_noflash.json 2> gpurun_out/bench_noflash.err; echo "== bench noflash rc=$?"; cut -c1-300 gpurun_out/bench_noflash.json
