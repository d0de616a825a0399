#!/bin/bash
# round-2 evidence, third pass: final bench line with the per-shape GEMM table, ncu --set full of the 441-key attention
# kernels and the BEV scatter kernels, a 4-step launch list, GEMM calibration point.  Output < 64 MiB.
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
BEVBERT_BENCH_VERBOSE=1 timeout 200 python bench.py --steps 22 --warmup 11 > gpurun_out/r02_bench_final_n1.json 2> gpurun_out/r02_bench_final_n1.err; echo "== bench rc=$?"
ATTN_ONLY="bev self 441" timeout 100 ncu --set full --clock-control none --import-source on -k regex:attn_tc -c 3 -f -o gpurun_out/r02_ncu_attn_shape1 python scripts/bench_attn.py > gpurun_out/r02_ncu_attn_shape1.log 2>&1; echo "shape 1 rc=$?"
HBM_ONLY=scatter timeout 100 ncu --set full --clock-control none -k regex:'scatter_' -c 3 -f -o gpurun_out/r02_ncu_hbm_scatter python scripts/bench_hbm.py > gpurun_out/r02_ncu_hbm_scatter.log 2>&1; echo "hbm scatter rc=$?"
timeout 170 ncu --metrics gpu__time_duration.sum --clock-control none -s 7400 -c 2400 --csv --log-file gpurun_out/r02_launches_final.csv \
  python bench.py --steps 2 --warmup 12 --no-cpu-baseline --graphs 0 --side-stream 0 > gpurun_out/r02_launches_final.log 2>&1
echo "== ncu launches rc=$?"; wc -l gpurun_out/r02_launches_final.csv
(cd vln-bevbert_b200/csrc/build && timeout 60 ncu --set full --clock-control none -k regex:gemm_tc -s 3 -c 1 -f -o ../../../gpurun_out/r02_ncu_gemm_sq8k ./selftest_gemm perf_sq8k > ../../../gpurun_out/r02_ncu_gemm_sq8k.log 2>&1; echo "gemm sq8k rc=$?")
while [ "$(du -sm gpurun_out | cut -f1)" -gt 58 ]; do big=$(ls -S gpurun_out/*.ncu-rep | head -1); echo "dropping $big"; rm -f "$big"; done
du -sm gpurun_out
