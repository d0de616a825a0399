#!/bin/bash
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-cpu-baseline > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "== n2 rc=$?"
cat gpurun_out/bench_n2.json | cut -c1-250
python - <<'PY'
import json
try:
    lines=[l for l in open('gpurun_out/bench_n2.json').read().strip().split('\n') if l.startswith('{')]
    print("json lines:", len(lines), "total lines:", len(open('gpurun_out/bench_n2.json').read().strip().split('\n')))
    d=json.loads(lines[-1])
    print('N2', round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']))
except Exception as e: print('ERR', e)
PY
tail -5 gpurun_out/bench_n2.err
