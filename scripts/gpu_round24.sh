#!/bin/bash
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
for i in 1 2; do
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_r24_$i.json 2> gpurun_out/bench_r24_$i.err; echo "== bench $i rc=$?"
done
python - <<'PY'
import json
for f in ['bench_r24_1','bench_r24_2']:
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().split('\n')[-1])
        print(f, round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'roof', round(d['roofline']['achieved']), round(d['roofline']['gemm_ms_per_step'],2), 'launches', d['gpu_launches'])
    except Exception as e: print(f, 'ERR', e)
PY
timeout 300 python scripts/prof_host.py > gpurun_out/prof_host2.log 2>&1; head -6 gpurun_out/prof_host2.log
