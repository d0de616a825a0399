#!/bin/bash
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 500 python -m pytest tests/test_graphs_gpu.py tests/test_kernels_gpu.py -m gpu -q -rf --no-header -p no:cacheprovider -k "graph or side_stream or masksem or layernorm" > gpurun_out/r2r_tests.log 2>&1
echo "== tests rc=$?"; tail -n 8 gpurun_out/r2r_tests.log | cut -c1-300
for ss in 1 0; do timeout 400 python bench.py --steps 22 --warmup 11 --no-cpu-baseline --side-stream $ss > gpurun_out/r2r_bench_ss$ss.json 2> gpurun_out/r2r_bench_ss$ss.err; tail -2 gpurun_out/r2r_bench_ss$ss.err | cut -c1-300; python -c "
import json; d=json.load(open('gpurun_out/r2r_bench_ss$ss.json')); print('side-stream $ss: value %.0f (%.2f ms) e2e %.0f (%.2f ms) launches %d' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['gpu_launches']))"; done
