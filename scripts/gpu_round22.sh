#!/bin/bash
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 600 python -m pytest tests -m gpu -x -q -k "scatter_mean or bf16_wire or gemm or native" > gpurun_out/test_sel.log 2>&1; echo "== pytest sel rc=$?"; tail -3 gpurun_out/test_sel.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_r22.json 2> gpurun_out/bench_r22.err; echo "== bench rc=$?"
timeout 600 python bench.py --no-cpu-baseline --wire fp32 > gpurun_out/bench_r22_fp32.json 2> gpurun_out/bench_r22_fp32.err; echo "== bench fp32 wire rc=$?"
python - <<'PY'
import json
for f in ['bench_r22','bench_r22_fp32']:
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().split('\n')[-1])
        print(f, round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), d['e2e']['h2d_bytes_per_step'], 'roof', round(d['roofline']['achieved']), round(d['roofline']['gemm_ms_per_step'],2), 'launches', d['gpu_launches'])
    except Exception as e: print(f, 'ERR', e)
PY
tail -3 gpurun_out/bench_r22.err
