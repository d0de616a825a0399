#!/bin/bash
cd vln-bevbert_b200/csrc/build
for c in perf_qkv perf_lang_ffn1 perf_lang_dx; do
  BB_GEMM_2CTA=0 BB_GEMM_TRACE=1 timeout -s KILL 40 ./selftest_gemm $c
done 2>&1 | tee ../../../gpurun_out/gemm_trace.log
for c in nt_gelu_aux nt_dgelu_add perf_ffn1 perf_lang_ffn1; do
  BB_FAST_GELU=1 timeout -s KILL 40 ./selftest_gemm $c
  BB_FAST_GELU=0 timeout -s KILL 40 ./selftest_gemm $c
done 2>&1 | tee ../../../gpurun_out/gemm_fastgelu.log
