#!/bin/bash
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-cpu-baseline > gpurun_out/final_n2.json 2> gpurun_out/final_n2.err; echo "== n2 rc=$? stdout lines: $(wc -l < gpurun_out/final_n2.json)"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/final_n2.json').read().strip().split('\n')[-1])
    print('N2', round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), d['config']['workload'][-80:])
except Exception as e: print('ERR', e)
PY
tail -3 gpurun_out/final_n2.err
