#!/bin/bash
# round-2 evidence: ncu --set full of the attention kernels per shape, the HBM kernels, and a GEMM calibration point
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
i=0
for shape in "bev self 441" "bev->lang 441x80" "lang->bev 80x441" "lang self 80"; do
  i=$((i+1))
  ATTN_ONLY="$shape" timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_tc -c 3 -f -o gpurun_out/r02_ncu_attn_shape$i python scripts/bench_attn.py > gpurun_out/r02_ncu_attn_shape$i.log 2>&1
  echo "shape $i ($shape) rc=$?"
done
for k in scatter layernorm colsum; do
  HBM_ONLY=$k timeout 300 ncu --set full --clock-control none --import-source on -k regex:'scatter_|layernorm_|colsum' -c 4 -f -o gpurun_out/r02_ncu_hbm_$k python scripts/bench_hbm.py > gpurun_out/r02_ncu_hbm_$k.log 2>&1
  echo "hbm $k rc=$?"
done
cd vln-bevbert_b200/csrc/build
timeout 200 ncu --set full --clock-control none -k regex:gemm_tc -s 3 -c 1 -f -o ../../../gpurun_out/r02_ncu_gemm_sq8k ./selftest_gemm perf_sq8k > ../../../gpurun_out/r02_ncu_gemm_sq8k.log 2>&1; echo "gemm sq8k rc=$?"
cd ../../..
timeout 300 python scripts/bench_attn.py > gpurun_out/r02_bench_attn_final.log 2>&1; cat gpurun_out/r02_bench_attn_final.log
timeout 300 python scripts/bench_hbm.py > gpurun_out/r02_bench_hbm_final.log 2>&1; cat gpurun_out/r02_bench_hbm_final.log
ls -la gpurun_out/r02_*.ncu-rep
