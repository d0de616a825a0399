#!/bin/bash
# PDL + step arena + static-buffer e2e: correctness first, then A/B bench.
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
( cd vln-bevbert_b200/csrc/build && timeout 300 ./selftest_gemm > ../../../gpurun_out/selftest_pdl.log 2>&1; echo "== selftest rc=$?"; tail -4 ../../../gpurun_out/selftest_pdl.log )
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/test_gpu_all.log 2>&1; echo "== pytest gpu rc=$?"; tail -3 gpurun_out/test_gpu_all.log
BB_PDL=0 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_pdl0.json 2> gpurun_out/bench_pdl0.err; echo "== pdl0 rc=$?"; cut -c1-400 gpurun_out/bench_pdl0.json
BB_PDL=1 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_pdl1.json 2> gpurun_out/bench_pdl1.err; echo "== pdl1 rc=$?"; cut -c1-400 gpurun_out/bench_pdl1.json
