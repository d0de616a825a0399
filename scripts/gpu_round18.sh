#!/bin/bash
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
cd vln-bevbert_b200/csrc/build
for mode in 1 0; do
  echo "===== BB_GEMM_2CTA=$mode"
  for c in $(./selftest_gemm list); do
    BB_GEMM_2CTA=$mode timeout -s KILL 30 ./selftest_gemm $c 2>&1 | tail -2 || echo "CASE $c exit=$?"
  done
done 2>&1 | tee ../../../gpurun_out/selftest_2cta_b.log | grep -E "=====|FAIL|exit=|perf_|Killed|error"
cd ../../..
grep -c PASS gpurun_out/selftest_2cta_b.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_2cta.json 2> gpurun_out/bench_2cta.err; echo "== bench rc=$?"; cut -c1-200 gpurun_out/bench_2cta.json
BENCH_E2E_NOCOPY=1 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_nocopy.json 2> gpurun_out/bench_nocopy.err; echo "== nocopy rc=$?"
BENCH_E2E_NOITEM=1 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_noitem.json 2> gpurun_out/bench_noitem.err; echo "== noitem rc=$?"
python - <<'PY'
import json
for f in ['bench_2cta','bench_nocopy','bench_noitem']:
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().split('\n')[-1])
        print(f, round(d['value']), 'e2e', round(d['e2e']['value']), 'roof', round(d['roofline']['achieved']))
    except Exception as e: print(f, 'ERR', e)
PY
