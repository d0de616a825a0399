#!/bin/bash
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
nvidia-smi -L
timeout 600 python scripts/prof_host.py > gpurun_out/prof_host.log 2>&1; echo "== prof rc=$?"; head -n 5 gpurun_out/prof_host.log | cut -c1-160
timeout 900 python bench.py --steps 22 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "== bench n1 rc=$?"; tail -n 2 gpurun_out/bench_n1.err; cut -c1-400 gpurun_out/bench_n1.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 22 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "== bench n2 rc=$?"; tail -n 5 gpurun_out/bench_n2.err | cut -c1-300; cut -c1-400 gpurun_out/bench_n2.json
