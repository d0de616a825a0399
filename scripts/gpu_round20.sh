#!/bin/bash
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
BEVBERT_BENCH_VERBOSE=1 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_r20.json 2> gpurun_out/bench_r20.err; echo "== bench rc=$?"
BB_GEMM_2CTA=0 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_r20_1cta.json 2> gpurun_out/bench_r20_1cta.err; echo "== bench 1cta rc=$?"
python - <<'PY'
import json
for f in ['bench_r20','bench_r20_1cta']:
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().split('\n')[-1])
        print(f, round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'roof', round(d['roofline']['achieved']), round(d['roofline']['gemm_ms_per_step'],2), 'launches', d['gpu_launches'])
    except Exception as e: print(f, 'ERR', e)
PY
grep gemm-shape gpurun_out/bench_r20.err | head -45
tail -3 gpurun_out/bench_r20.err
