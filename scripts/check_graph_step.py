"""Quick GPU check of graphs.GraphedTrainStep (subset of tests/test_graphs_gpu.py, sized for a one-minute slot)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch
import test_graphs_gpu as T
from bevbert_b200.graphs import GraphedTrainStep
from bevbert_b200.parallel import direct_param_grads

t0 = time.time()
direct_param_grads(True)
seq = ["mlm", "sap", "mlm", "sap", "masksem", "mlm", "sap", "mlm", "sap"]
b = T._batches()
e0, _ = T._run(seq, b, 0.0, False)
g0, _ = T._run(seq, b, 0.0, True)
print("lr=0 eager  ", e0)
print("lr=0 graphed", g0)
assert all(abs(a - g) <= 2e-4 * abs(a) for a, g in zip(e0, g0))
print("replay == eager OK  %.1fs" % (time.time() - t0), flush=True)
m, o = T._setup(drop=0.1)
o.param_groups[0]["lr"] = o.param_groups[1]["lr"] = 0.0
step = GraphedTrainStep(m, o, warmup=1)
losses = [float(step(b["sap"], "sap")) for _ in range(6)]
assert len(set(round(x, 6) for x in losses[2:])) >= 3, losses
print("fresh masks OK", losses, flush=True)
m, o = T._setup()
step = GraphedTrainStep(m, o, warmup=1)
ls = [step(b["mlm"], "mlm").clone() for _ in range(40)]  # no host sync between replays (the loss tensor is static)
torch.cuda.synchronize()
ls = [float(x) for x in ls]
assert all(x == x for x in ls) and ls[-1] < ls[2], ls
print("40 back-to-back replays OK: loss %.4f -> %.4f  (%.1fs)" % (ls[2], ls[-1], time.time() - t0))
print("CHECK PASSED")
