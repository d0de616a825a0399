"""Whole-step CUDA-graph replay (bevbert_b200/graphs.py) reproduces the eager training loop: same losses over several
optimizer steps on the same static batches (dropout off: identical kernels, only atomics order may differ), fresh
dropout masks on every replay (device-resident salt), and the eager fallback for the tasks that need a host sync."""
import pytest
import torch

from bevbert_b200 import synth
from bevbert_b200.graphs import GraphedTrainStep
from bevbert_b200.model.ops import prepare_batch
from bevbert_b200.model.pretrain_cmt import GlocalTextPathCMTPreTraining
from bevbert_b200.optim import AdamW, build_param_groups
from bevbert_b200.parallel import direct_param_grads
from helpers import small_config, small_synth

pytestmark = pytest.mark.gpu


def _setup(drop=0.0):
    cfg = small_config(hidden_dropout_prob=drop, attention_probs_dropout_prob=drop, feat_dropout=0.0)
    model = synth.det_init_(GlocalTextPathCMTPreTraining(cfg), seed=3).cuda().train()
    opt = AdamW(build_param_groups(model, 0.01), lr=1e-4, betas=(0.9, 0.98), max_grad_norm=5.0, runtime=model.rt)
    return model, opt


def _batches():
    scfg = small_synth()
    return {t: synth.batch_to(prepare_batch(synth.make_batch(scfg, seed=40 + i, task=t)), "cuda")
            for i, t in enumerate(["mlm", "sap", "masksem"])}


def _run(seq, b, lr, graphed):
    m, o = _setup()
    for g in o.param_groups:
        g["lr"] = lr
    w0 = m.bert.lang_encoder.layer[0].attention.self.query.weight.detach().clone()
    losses = []
    if graphed:
        step = GraphedTrainStep(m, o, warmup=1)
        losses = [float(step(b[t], t)) for t in seq]
        assert step.launches(b["mlm"], "mlm") and step.launches(b["sap"], "sap")
    else:
        for t in seq:
            loss = m(b[t], t).mean()
            loss.backward()
            o.step()
            losses.append(float(loss.detach()))
    return losses, m.bert.lang_encoder.layer[0].attention.self.query.weight.detach() - w0


def test_graph_replay_matches_eager_training():
    direct_param_grads(True)
    try:
        seq = ["mlm", "sap", "mlm", "sap", "masksem", "mlm", "sap", "mlm", "sap"]
        b = _batches()
        # (1) frozen weights: every replay must reproduce the eager forward / backward exactly (up to atomics order)
        e0, _ = _run(seq, b, 0.0, False)
        g0, _ = _run(seq, b, 0.0, True)
        print("lr=0 eager  ", e0)
        print("lr=0 graphed", g0)
        for a, g in zip(e0, g0):
            assert abs(a - g) <= 2e-4 * abs(a), (e0, g0)
        # (2) training: the optimizer inside the graph follows the eager trajectory.  Adam's first steps are sign-like
        # (m / sqrt(v) ~ +-1), so split-K atomics noise moves individual weights by +-lr and the steep SAP loss amplifies
        # it: two EAGER runs differ by the same few per cent, which sets the bar.
        e1, d1 = _run(seq, b, 1e-4, False)
        e2, d2 = _run(seq, b, 1e-4, False)
        g1, dg = _run(seq, b, 1e-4, True)
        noise = max(abs(a - c) / abs(a) for a, c in zip(e1, e2))
        err = max(abs(a - c) / abs(a) for a, c in zip(e1, g1))
        print("train eager  ", e1)
        print("train eager2 ", e2)
        print("train graphed", g1, "eager-vs-eager %.3e graphed-vs-eager %.3e" % (noise, err))
        # chaotic metric (two eager runs differ by 1.6-4 %): the exactness check is the lr = 0 part above; here the graph's
        # optimizer must follow the eager trajectory -- same direction of the accumulated update, losses in the same band
        assert err <= max(0.3, 6.0 * noise), (err, noise)       # observed: noise 1.6-4 %, err 0.2-11 %
        cos = float((dg * d1).sum() / (dg.norm() * d1.norm()))
        cos_ee = float((d2 * d1).sum() / (d2.norm() * d1.norm()))
        print("update cosine graphed-vs-eager %.4f eager-vs-eager %.4f" % (cos, cos_ee))
        assert cos > min(0.85, cos_ee - 0.1), (cos, cos_ee)
    finally:
        direct_param_grads(False)


def test_graph_replay_draws_fresh_dropout_masks():
    direct_param_grads(True)
    try:
        b = _batches()
        m, o = _setup(drop=0.1)
        o.param_groups[0]["lr"] = o.param_groups[1]["lr"] = 0.0      # frozen weights: loss differences = mask differences
        step = GraphedTrainStep(m, o, warmup=1)
        losses = [float(step(b["sap"], "sap")) for _ in range(6)]
        assert step.launches(b["sap"], "sap")
        assert len(set(round(x, 6) for x in losses[2:])) >= 3, losses   # replays 3..6 use different masks
    finally:
        direct_param_grads(False)


def test_masksem_sync_free_mean_equals_per_item_mean():
    """masksem in the graph-capturable mode (mean over all cells weighted by the device-side selection) gives the value
    and the gradients of `forward(...).mean()` on the reference's variable-length per-item loss."""
    b = _batches()
    m, _ = _setup()
    l1 = m(b["masksem"], "masksem").mean()
    l1.backward()
    g1 = {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}
    m.zero_grad(set_to_none=True)
    m.sync_free_mean = True
    l2 = m(b["masksem"], "masksem")
    m.sync_free_mean = False
    assert l2.dim() == 0 and abs(float(l1) - float(l2)) <= 1e-4 * abs(float(l1))
    l2.backward()
    num = sum(float((p.grad - g1[n]).norm()) ** 2 for n, p in m.named_parameters() if p.grad is not None)
    den = sum(float(g.norm()) ** 2 for g in g1.values())
    assert (num / den) ** 0.5 < 2e-2, (num / den) ** 0.5        # two bf16 runs with different row sets in the head GEMMs


def test_side_stream_weight_gradients_match_single_stream():
    """parallel.enable_side_stream: weight-gradient GEMMs / bias column sums on a second stream (fork per call, one join
    after backward, buffers kept alive, second contributions deferred) give the single-stream gradients."""
    from bevbert_b200 import blocks
    from bevbert_b200.parallel import enable_side_stream
    b = _batches()
    grads = {}
    for side in (False, True):
        m, _ = _setup()
        direct_param_grads(True)
        if side:
            enable_side_stream(True)
        try:
            for t in ("mlm", "sap"):
                m.zero_grad(set_to_none=True)
                m(b[t], t).mean().backward()
                blocks.join_side()
                torch.cuda.synchronize()
                grads[(side, t)] = {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}
        finally:
            if side:
                enable_side_stream(False)
            direct_param_grads(False)
    for t in ("mlm", "sap"):
        a, c = grads[(False, t)], grads[(True, t)]
        assert set(a) == set(c)
        num = sum(float((c[n] - a[n]).norm()) ** 2 for n in a)
        den = sum(float(a[n].norm()) ** 2 for n in a)
        assert (num / den) ** 0.5 < 1e-3, (t, (num / den) ** 0.5)
