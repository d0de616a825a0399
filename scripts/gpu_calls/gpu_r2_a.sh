#!/bin/bash
# round-2 call A: staged-GELU validation, HBM-kernel microbench + ncu --set full, baseline attention microbench
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv
cd vln-bevbert_b200/csrc/build
for c in nt_gelu_aux nt_dgelu_add nt_drelu perf_ffn1 perf_lang_ffn1 perf_lang_dx perf_qkv; do
  echo "--- default $c"; timeout -s KILL 90 ./selftest_gemm $c 2>&1 | tail -3
  echo "--- staged $c"; BB_GEMM_GELU_STAGED=1 timeout -s KILL 90 ./selftest_gemm $c 2>&1 | tail -3
done > ../../../gpurun_out/r2a_selftest_staged.log 2>&1
cd ../../..
cat gpurun_out/r2a_selftest_staged.log
timeout 300 python scripts/bench_hbm.py > gpurun_out/r2a_bench_hbm.log 2>&1; cat gpurun_out/r2a_bench_hbm.log
timeout 300 python scripts/bench_attn.py > gpurun_out/r2a_bench_attn.log 2>&1; cat gpurun_out/r2a_bench_attn.log
for k in scatter layernorm colsum cast; do
  HBM_ONLY=$k timeout 600 ncu --set full --clock-control none --import-source on -k regex:'scatter_|layernorm_|colsum|cast_' \
     -c 6 -f -o gpurun_out/r2a_ncu_hbm_$k python scripts/bench_hbm.py > gpurun_out/r2a_ncu_hbm_$k.log 2>&1
  tail -2 gpurun_out/r2a_ncu_hbm_$k.log
done
ls -la gpurun_out/*.ncu-rep
