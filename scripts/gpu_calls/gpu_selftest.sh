#!/bin/bash
# Runs the standalone GEMM self-test on the GPU box, one case per process (a hang cannot mask other cases).
mkdir -p gpurun_out
LOG=gpurun_out/selftest_gemm.log
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv > $LOG 2>&1
cd vln-bevbert_b200/csrc/build
for c in $(./selftest_gemm list); do
  timeout -s KILL 90 ./selftest_gemm $c >> ../../../$LOG 2>&1 || echo "CASE $c exit=$?" >> ../../../$LOG
done
cd ../../..
cat $LOG
