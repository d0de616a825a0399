#!/bin/bash
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
bash scripts/gpu_selftest.sh > /dev/null 2>&1; echo "selftest PASS count: $(grep -c PASS gpurun_out/selftest_gemm.log)"; grep -E "FAIL|ERROR|exit=|perf_" gpurun_out/selftest_gemm.log | cut -c1-200
cd vln-bevbert_b200/csrc/build
for st in 2 3 4; do for c in perf_lang_ffn1 perf_qkv perf_sq8k; do echo -n "stages=$st "; BB_GEMM_STAGES=$st timeout 120 ./selftest_gemm $c | cut -c1-40,100-140; done; done
cd ../../..
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -rf --no-header -p no:cacheprovider -k "fused or native or attention_core or softmax" > gpurun_out/test_kernels_sel.log 2>&1
echo "== kernel sel rc=$?"; tail -n 12 gpurun_out/test_kernels_sel.log | cut -c1-220
timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -q -rf -s --no-header -p no:cacheprovider > gpurun_out/test_model_gpu.log 2>&1
echo "== model tests rc=$?"; grep -E "passed|failed" gpurun_out/test_model_gpu.log | tail -n 3
BEVBERT_BENCH_VERBOSE=1 timeout 900 python bench.py --steps 22 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "== bench rc=$?"; grep gemm-shape gpurun_out/bench.err | head -16; tail -n 3 gpurun_out/bench.err; cat gpurun_out/bench.json
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -s 8000 -c 8000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 11 --warmup 11 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
echo "== ncu launches rc=$?"; wc -l gpurun_out/launches.csv
