#!/bin/bash
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-cpu-baseline > gpurun_out/final_n2.json 2> gpurun_out/final_n2.err; echo "== n2 rc=$? stdout lines: $(wc -l < gpurun_out/final_n2.json)"
timeout 900 python bench.py > gpurun_out/final_n1.json 2> gpurun_out/final_n1.err; echo "== n1 rc=$? stdout lines: $(wc -l < gpurun_out/final_n1.json)"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python - <<'PY'
import json
for f in ['final_n2','final_n1']:
    try:
        d=json.loads(open('gpurun_out/%s.json'%f).read().strip().split('\n')[-1])
        print(f, round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'roof', d['roofline'] and round(d['roofline']['achieved']), 'cpu', d.get('cpu_baseline') and d['cpu_baseline']['value'], 'launches', d['gpu_launches'], d['clocks'])
    except Exception as e: print(f, 'ERR', e)
PY
