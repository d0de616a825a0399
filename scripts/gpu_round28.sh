#!/bin/bash
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
cd vln-bevbert_b200/csrc/build
for c in $(./selftest_gemm list); do
  timeout -s KILL 30 ./selftest_gemm $c 2>&1 | tail -1 || echo "CASE $c exit=$?"
done 2>&1 | tee ../../../gpurun_out/selftest_epi.log | grep -E "FAIL|exit=|perf_|Killed|error|gelu|drelu"
grep -c PASS ../../../gpurun_out/selftest_epi.log
cd ../../..
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/test_gpu_all.log 2>&1; echo "== pytest gpu rc=$?"; tail -3 gpurun_out/test_gpu_all.log
BEVBERT_BENCH_VERBOSE=1 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_r28.json 2> gpurun_out/bench_r28.err; echo "== bench rc=$?"
python - <<'PY'
import json
for f in ['bench_r28']:
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().split('\n')[-1])
        print(f, round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'roof', round(d['roofline']['achieved']), round(d['roofline']['gemm_ms_per_step'],2), 'launches', d['gpu_launches'])
    except Exception as e: print(f, 'ERR', e)
PY
grep gemm-shape gpurun_out/bench_r28.err | head -16
