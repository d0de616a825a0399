#!/bin/bash
# 2-CTA GEMM: correctness + perf, every case in its own process with a hard timeout
mkdir -p gpurun_out
cd vln-bevbert_b200/csrc/build
for mode in 1 0; do
  echo "===== BB_GEMM_2CTA=$mode"
  for c in $(./selftest_gemm list); do
    BB_GEMM_2CTA=$mode timeout -s KILL 60 ./selftest_gemm $c 2>&1 | tail -2 || echo "CASE $c exit=$?"
  done
done 2>&1 | tee ../../../gpurun_out/selftest_2cta.log | grep -E "=====|FAIL|exit=|perf_|Killed|error" 
