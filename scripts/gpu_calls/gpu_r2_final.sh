#!/bin/bash
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 1500 python -m pytest tests -m gpu -q -x -rf --no-header -p no:cacheprovider > gpurun_out/r2v_test_all.log 2>&1
echo "== all gpu tests rc=$?"; tail -n 12 gpurun_out/r2v_test_all.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2v_smoke.log 2>&1; echo "== smoke rc=$?"; tail -2 gpurun_out/r2v_smoke.log | cut -c1-600
timeout 600 python bench.py --steps 22 --warmup 11 > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err; echo "== bench rc=$?"; tail -2 gpurun_out/r2v_bench.err | cut -c1-300; python -c "
import json; d=json.load(open('gpurun_out/r2v_bench.json')); r=d['roofline']; print('value %.0f (%.2f ms) e2e %.0f (%.2f ms) launches %d gemm %.0f TF/s frac %.3f cpu %.2f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['gpu_launches'], r['achieved'], r['frac'], d['cpu_baseline']['value']))"
