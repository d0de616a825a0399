"""Host-side (Python / launch) overhead profile of one training step on the GPU box (cProfile, top by self time)."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import MIX, full_config  # noqa: E402
from bevbert_b200 import synth  # noqa: E402
from bevbert_b200.model.pretrain_cmt import GlocalTextPathCMTPreTraining  # noqa: E402

dev = torch.device("cuda", 0)
model = synth.det_init_(GlocalTextPathCMTPreTraining(full_config()), seed=3).to(dev).train()
from bevbert_b200.optim import FusedAdamW  # noqa: E402
from bevbert_b200.parallel import direct_param_grads  # noqa: E402
direct_param_grads(True)
opt = FusedAdamW(model.parameters(), lr=5e-5)
import gc  # noqa: E402
gc.collect(); gc.freeze(); gc.disable()
from bevbert_b200.model.ops import prepare_batch  # noqa: E402
batches = {t: synth.batch_to(prepare_batch(synth.make_batch(synth.SynthConfig(batch_size=32), seed=1, task=t)), dev) for t in set(MIX)}


def step(t):
    model(batches[t], t).mean().backward()
    opt.step()


for i in range(6):
    step(MIX[i % 11])
torch.cuda.synchronize()
for task in ("sap", "mlm", "masksem", "mix"):
    n = 11
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n):
        step(MIX[i % 11] if task == "mix" else task)
    t1 = time.perf_counter()          # host time to ENQUEUE the steps (pure CPU cost when the task has no host sync)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("%-8s host enqueue %.2f ms/step, until GPU idle %.2f ms/step" % (task, (t1 - t0) / n * 1e3, (t2 - t0) / n * 1e3))
pr = cProfile.Profile()
pr.enable()
for i in range(11):
    step(MIX[i % 11])
torch.cuda.synchronize()
pr.disable()
ps = pstats.Stats(pr)
ps.sort_stats("tottime").print_stats(30)
ps.sort_stats("cumtime").print_stats(40)
