#!/bin/bash
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
BEVBERT_BENCH_VERBOSE=1 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_v.json 2> gpurun_out/bench_v.err; echo "== bench rc=$?"
grep gemm-shape gpurun_out/bench_v.err | head -70
timeout 120 python - <<'PY'
import torch, time
x = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
d = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3): d.copy_(x, non_blocking=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): d.copy_(x, non_blocking=True)
e1.record(); torch.cuda.synchronize()
print("H2D pinned GB/s", 10 * 0.268435456 / (e0.elapsed_time(e1) * 1e-3))
e0.record()
for _ in range(10): x.copy_(d, non_blocking=True)
e1.record(); torch.cuda.synchronize()
print("D2H pinned GB/s", 10 * 0.268435456 / (e0.elapsed_time(e1) * 1e-3))
PY
cd vln-bevbert_b200/csrc/build
for c in perf_qkv perf_ffn1 perf_ffn2 perf_sq8k perf_lang_ffn1 perf_lang_dx perf_lang_dw perf_attn_s441 perf_dw; do timeout 120 ./selftest_gemm $c; done
