#!/bin/bash
# One GPU-box visit: model parity tests, smoke, bench (N=1), ncu launch list of a short bench run.
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -q -rf -s --no-header -p no:cacheprovider > gpurun_out/test_model_gpu.log 2>&1
echo "== model tests rc=$?"; grep -E "task=|    bert|    .*ours|passed|failed" gpurun_out/test_model_gpu.log | tail -n 70
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "== smoke rc=$?"; tail -n 5 gpurun_out/smoke.log
timeout 900 python bench.py --steps 22 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "== bench rc=$?"; tail -n 3 gpurun_out/bench.err; cat gpurun_out/bench.json
if [ "$1" == "ncu" ]; then
  timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
  echo "== ncu rc=$?"; wc -l gpurun_out/launches.csv
fi
