#!/bin/bash
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
cd vln-bevbert_b200/csrc/build
for c in perf_lang_ffn1 perf_lang_dx perf_lang_dw perf_attn_s441; do timeout 120 ./selftest_gemm $c; done
for c in perf_lang_ffn1 perf_lang_dx perf_attn_s441 perf_qkv; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 3 -c 1 -o ../../../gpurun_out/ncu_$c -f ./selftest_gemm $c > ../../../gpurun_out/ncu_$c.log 2>&1
  echo "== ncu $c rc=$?"
done
cd ../../..
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -s 9000 -c 5000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 11 --warmup 11 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
echo "== ncu launches rc=$?"; wc -l gpurun_out/launches.csv
