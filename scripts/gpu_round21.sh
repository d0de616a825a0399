#!/bin/bash
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_r21.json 2> gpurun_out/bench_r21.err; echo "== bench rc=$?"
python - <<'PY'
import json
for f in ['bench_r21']:
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().split('\n')[-1])
        print(f, round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'roof', round(d['roofline']['achieved']), round(d['roofline']['gemm_ms_per_step'],2), 'launches', d['gpu_launches'])
    except Exception as e: print(f, 'ERR', e)
PY
tail -3 gpurun_out/bench_r21.err
