#!/bin/bash
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -q -rf -s --no-header -p no:cacheprovider > gpurun_out/test_model_gpu.log 2>&1
echo "== model tests rc=$?"; grep -E "task=| ours .* autocast|passed|failed" gpurun_out/test_model_gpu.log | cut -c1-220 | tail -n 70
timeout 600 python scripts/prof_host.py > gpurun_out/prof_host.log 2>&1; echo "== prof rc=$?"; head -n 60 gpurun_out/prof_host.log | cut -c1-200
BEVBERT_BENCH_VERBOSE=1 timeout 900 python bench.py --steps 22 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "== bench rc=$?"; grep gemm-shape gpurun_out/bench.err; cat gpurun_out/bench.json
cd vln-bevbert_b200/csrc/build
for c in perf_qkv perf_ffn2 attn_qk attn_pv perf_dw; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 1 -o ../../../gpurun_out/ncu_$c -f ./selftest_gemm $c > ../../../gpurun_out/ncu_$c.log 2>&1
  echo "== ncu $c rc=$?"
done
