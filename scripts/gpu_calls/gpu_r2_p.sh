#!/bin/bash
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -rf --no-header -p no:cacheprovider -k "layernorm or colsum or native" > gpurun_out/r2p_test_rows.log 2>&1
echo "== row tests rc=$?"; tail -n 4 gpurun_out/r2p_test_rows.log
HBM_ONLY= timeout 300 python scripts/bench_hbm.py 2>&1 | grep -E "layernorm|colsum" 
timeout 900 python bench.py --steps 22 --warmup 11 --no-cpu-baseline > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; tail -3 gpurun_out/r2p_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r2p_bench.json')); print('value %.0f (%.2f ms) e2e %.0f (%.2f ms)' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step']))"
timeout 1200 python -m pytest tests/test_model_gpu.py -m gpu -q -x -rf --no-header -p no:cacheprovider > gpurun_out/r2p_test_model.log 2>&1
echo "== model tests rc=$?"; tail -n 5 gpurun_out/r2p_test_model.log
