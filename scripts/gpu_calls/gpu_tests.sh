#!/bin/bash
# GPU-box test driver: kernel parity, model parity (each file in its own process), logs under gpurun_out/.
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
for f in test_kernels_gpu test_model_gpu; do
  timeout 900 python -m pytest tests/$f.py -m gpu -q -rf -s --no-header -p no:cacheprovider "$@" > gpurun_out/$f.log 2>&1
  echo "== $f rc=$?"; tail -n 60 gpurun_out/$f.log
done
