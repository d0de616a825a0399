#!/bin/bash
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1 BEVBERT_NCCL_TIMEOUT_S=60
timeout 300 python -m pytest tests/test_ddp_gpu.py -m gpu -q -rf -s --no-header -p no:cacheprovider > gpurun_out/r2s_test_ddp.log 2>&1
echo "== ddp tests rc=$?"; grep -E "2-rank|passed|failed|Error" gpurun_out/r2s_test_ddp.log | cut -c1-400 | head -12
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 11 --warmup 5 --no-cpu-baseline > gpurun_out/r2s_bench_n2.json 2> gpurun_out/r2s_bench_n2.err
echo "== bench n2 rc=$?"; tail -3 gpurun_out/r2s_bench_n2.err | cut -c1-300; python -c "
import json; d=json.load(open('gpurun_out/r2s_bench_n2.json')); print('N=2 value %.0f (%.2f ms) e2e %.0f (%.2f ms) mode %s numa %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['step_mode'], d.get('numa_node')))"
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 11 --warmup 5 --no-cpu-baseline --graphs 0 > gpurun_out/r2s_bench_n2_eager.json 2> gpurun_out/r2s_bench_n2_eager.err
echo "== bench n2 eager rc=$?"; tail -3 gpurun_out/r2s_bench_n2_eager.err | cut -c1-300; python -c "
import json; d=json.load(open('gpurun_out/r2s_bench_n2_eager.json')); print('N=2 eager value %.0f (%.2f ms) e2e %.0f' % (d['value'], d['ms_per_step'], d['e2e']['value']))"
