#!/bin/bash
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -rf --no-header -p no:cacheprovider -k "native or layernorm or fused" > gpurun_out/test_kernels_sel.log 2>&1
echo "== kernel sel rc=$?"; tail -n 12 gpurun_out/test_kernels_sel.log | cut -c1-220
timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -q -rf -s --no-header -p no:cacheprovider > gpurun_out/test_model_gpu.log 2>&1
echo "== model tests rc=$?"; grep -E "passed|failed" gpurun_out/test_model_gpu.log | tail -n 3
timeout 600 python scripts/prof_host.py > gpurun_out/prof_host.log 2>&1; echo "== prof rc=$?"; head -n 40 gpurun_out/prof_host.log | cut -c1-160
timeout 900 python bench.py --steps 22 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "== bench rc=$?"; tail -n 3 gpurun_out/bench.err; cat gpurun_out/bench.json
