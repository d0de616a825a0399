#!/bin/bash
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
python scripts/h2d_bw.py 2>&1 | tee gpurun_out/r2o_h2d.log
timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -q -rf -s --no-header -p no:cacheprovider -k "wire_format" > gpurun_out/r2o_test_wire.log 2>&1
echo "== wire tests rc=$?"; grep -E "wire format|passed|failed" gpurun_out/r2o_test_wire.log
for i in 1 2; do timeout 900 python bench.py --steps 22 --warmup 11 --no-cpu-baseline > gpurun_out/r2o_bench_$i.json 2> gpurun_out/r2o_bench_$i.err; python -c "
import json; d=json.load(open('gpurun_out/r2o_bench_$i.json')); print('run $i value %.0f (%.2f ms) e2e %.0f (%.2f ms, %d MB)' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['h2d_bytes_per_step']/1e6))"; done
BB_PDL=0 timeout 900 python bench.py --steps 22 --warmup 11 --no-cpu-baseline > gpurun_out/r2o_bench_pdl0.json 2> gpurun_out/r2o_bench_pdl0.err; python -c "
import json; d=json.load(open('gpurun_out/r2o_bench_pdl0.json')); print('PDL=0 value %.0f (%.2f ms) e2e %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value']))"
