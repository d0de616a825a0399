#!/bin/bash
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -rf --no-header -p no:cacheprovider -k "fused or native or attention_core or softmax or dropout or cast" > gpurun_out/test_kernels_sel.log 2>&1
echo "== kernel sel rc=$?"; tail -n 12 gpurun_out/test_kernels_sel.log | cut -c1-220
timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -q -rf -s --no-header -p no:cacheprovider > gpurun_out/test_model_gpu.log 2>&1
echo "== model tests rc=$?"; grep -E "passed|failed" gpurun_out/test_model_gpu.log | tail -n 3
for fm in 512 128 0; do
  BB_FUSED_SCORES_MAX=$fm timeout 900 python bench.py --steps 22 --warmup 5 --no-cpu-baseline > gpurun_out/bench_fm$fm.json 2> gpurun_out/bench_fm$fm.err; echo "== bench fused_max=$fm rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/bench_fm$fm.json')); print('fused_max=$fm', round(d['value'],1), 'samples/s', round(d['ms_per_step'],2), 'ms/step', 'e2e', round(d['e2e']['value'],1), 'gemm ms/step', round(d['roofline']['gemm_ms_per_step'],2), 'gemm TF/s', round(d['roofline']['achieved'],1))"
done
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -s 8000 -c 8000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 11 --warmup 11 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
echo "== ncu launches rc=$?"; wc -l gpurun_out/launches.csv
