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
# launch list of the final bench step (eager launches on one stream so that every kernel is listed and serial)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 7400 -c 6700 --csv --log-file gpurun_out/r02_launches_final.csv \
  python bench.py --steps 11 --warmup 11 --no-cpu-baseline --graphs 0 --side-stream 0 > gpurun_out/r02_launches_final.log 2>&1
echo "== ncu launches rc=$?"; wc -l gpurun_out/r02_launches_final.csv
# final bench line (graph replay, side stream), every GEMM shape to stderr
BEVBERT_BENCH_VERBOSE=1 timeout 600 python bench.py --steps 22 --warmup 11 > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; echo "== bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_final.json')); r=d['roofline']; print('value %.0f (%.2f ms) e2e %.0f (%.2f ms) launches %d gemm %.0f TF/s frac %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['gpu_launches'], r['achieved'], r['frac']))"
