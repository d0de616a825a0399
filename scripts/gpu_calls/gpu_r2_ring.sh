#!/bin/bash
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 110 python -m pytest tests/test_optim_gpu.py tests/test_graphs_gpu.py -q -x -rf --no-header -p no:cacheprovider > gpurun_out/r2z_test.log 2>&1
echo "== tests rc=$?"; tail -n 6 gpurun_out/r2z_test.log | cut -c1-300
timeout 70 python bench.py --steps 11 --warmup 11 --no-cpu-baseline > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; echo "== bench rc=$?"; tail -2 gpurun_out/r2z_bench.err | cut -c1-300; python -c "
import json; d=json.load(open('gpurun_out/r2z_bench.json')); print('value %.0f (%.2f ms) e2e %.0f (%.2f ms) launches %d' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['gpu_launches']))"
