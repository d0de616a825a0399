#!/bin/bash
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1 BEVBERT_NCCL_TIMEOUT_S=60
timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 11 --warmup 5 --no-cpu-baseline > gpurun_out/r2t_bench_n2.json 2> gpurun_out/r2t_bench_n2.err
echo "== bench n2 rc=$?"; tail -4 gpurun_out/r2t_bench_n2.err | cut -c1-300; python -c "
import json; d=json.load(open('gpurun_out/r2t_bench_n2.json')); print('N=2 value %.0f (%.2f ms) e2e %.0f (%.2f ms) mode %s numa %s launches %d' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['step_mode'], d.get('numa_node'), d['gpu_launches']))"
