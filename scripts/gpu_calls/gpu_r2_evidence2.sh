#!/bin/bash
# round-2 evidence, second (size-bounded) pass: ncu --set full of the attention kernels at the cross-attention and
# language shapes, and of the rewritten LayerNorm / column-sum kernels.  Output must stay below gpurun's 64 MiB.
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
i=1
for shape in "bev->lang 441x80" "lang->bev 80x441" "lang self 80"; do
  i=$((i+1))
  ATTN_ONLY="$shape" timeout 200 ncu --set full --clock-control none --import-source on -k regex:attn_tc -c 3 -f -o gpurun_out/r02_ncu_attn_shape$i python scripts/bench_attn.py > gpurun_out/r02_ncu_attn_shape$i.log 2>&1
  echo "shape $i ($shape) rc=$?"
done
for k in layernorm colsum; do
  HBM_ONLY=$k timeout 200 ncu --set full --clock-control none -k regex:'layernorm_|colsum' -c 3 -f -o gpurun_out/r02_ncu_hbm_$k python scripts/bench_hbm.py > gpurun_out/r02_ncu_hbm_$k.log 2>&1
  echo "hbm $k rc=$?"
done
du -sm gpurun_out
# keep the merge-back below the limit whatever happened above
while [ "$(du -sm gpurun_out | cut -f1)" -gt 58 ]; do big=$(ls -S gpurun_out/*.ncu-rep | head -1); echo "dropping $big"; rm -f "$big"; done
ls -la gpurun_out
