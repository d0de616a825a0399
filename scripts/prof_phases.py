"""Host-side enqueue time of the phases of one SAP / MLM training step (perf_counter, no cProfile)."""
import collections
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import MIX, full_config  # noqa: E402
from bevbert_b200 import synth  # noqa: E402
from bevbert_b200.model.ops import prepare_batch  # noqa: E402
from bevbert_b200.model.pretrain_cmt import GlocalTextPathCMTPreTraining  # noqa: E402

dev = torch.device("cuda", 0)
model = synth.det_init_(GlocalTextPathCMTPreTraining(full_config()), seed=3).to(dev).train()
opt = torch.optim.AdamW(model.parameters(), lr=5e-5, fused=True)
batches = {t: synth.batch_to(prepare_batch(synth.make_batch(synth.SynthConfig(batch_size=32), seed=1, task=t)), dev) for t in set(MIX)}
T = collections.defaultdict(float)


def timed(obj, name, label):
    fn = getattr(obj, name)

    def w(*a, **k):
        t0 = time.perf_counter()
        r = fn(*a, **k)
        T[label] += time.perf_counter() - t0
        return r
    setattr(obj, name, w)


timed(model, "lift_splat", "fwd.lift_splat")
timed(model.bert.embeddings, "forward", "fwd.text_embed")
timed(model.bert.lang_encoder, "forward", "fwd.lang_encoder")
timed(model.bert.img_embeddings, "forward", "fwd.img_embeddings(+pano)")
timed(model.bert.global_encoder, "forward", "fwd.global_encoder")
timed(model.bert.global_encoder, "gmap_input_embedding", "fwd.  gmap_input_embedding")
timed(model.bert.local_encoder, "forward", "fwd.local_encoder")
timed(model.bert.local_encoder, "bev_input_embedding", "fwd.  bev_input_embedding")
timed(model, "forward_sap", "fwd.forward_sap(total)")
timed(model, "forward_mlm", "fwd.forward_mlm(total)")


def step(t):
    t0 = time.perf_counter()
    loss = model(batches[t], t).mean()
    t1 = time.perf_counter()
    loss.backward()
    t2 = time.perf_counter()
    opt.step()
    opt.zero_grad(set_to_none=True)
    t3 = time.perf_counter()
    T["forward(total)"] += t1 - t0
    T["backward"] += t2 - t1
    T["optimizer+zero_grad"] += t3 - t2


for task in ("sap", "mlm"):
    for i in range(4):
        step(task)
    torch.cuda.synchronize()
    T.clear()
    n = 10
    for i in range(n):
        step(task)
        torch.cuda.synchronize()      # isolate pure host cost: GPU idle at the start of every step
    print("== %s: host ms per step (GPU drained between steps)" % task)
    for k, v in sorted(T.items()):
        print("   %-32s %7.2f" % (k, v / n * 1e3))
