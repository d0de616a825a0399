"""pinned host -> device copy bandwidth of the box (the e2e leg of bench.py moves ~126-150 MB per step)."""
import torch
for mb in (16, 128, 512):
    h = torch.empty(mb << 20, dtype=torch.uint8).pin_memory()
    d = torch.empty(mb << 20, dtype=torch.uint8, device="cuda")
    for _ in range(2):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        d.copy_(h, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    print("H2D pinned %4d MB: %.1f GB/s" % (mb, 5 * mb / 1024 / (e0.elapsed_time(e1) * 1e-3)))
