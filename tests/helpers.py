"""Shared helpers for the parity tests."""
import torch

from bevbert_b200 import synth
from bevbert_b200.config import make_config

SMALL = dict(num_l_layers=2, num_x_layers=2, num_pano_layers=1, vocab_size=3000, bev_dim=11, bev_res=1.0,
             hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, feat_dropout=0.0)


def small_config(**kw):
    d = dict(SMALL)
    d.update(kw)
    return make_config(**d)


def small_synth(**kw):
    d = dict(batch_size=2, bev_dim=11, bev_res=1.0, pano_min=4, pano_max=6, gmap_min=8, gmap_max=8, vocab_hi=2000)
    d.update(kw)
    return synth.SynthConfig(**d)


def rel_l2(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def grad_report(named_grads, ref_grads, floor_frac=1e-4):
    """max over parameters of ||g - g_ref|| / max(||g_ref||, floor) where floor = floor_frac * largest
    reference-gradient norm (gradients that are mathematically zero -- e.g. the key bias -- are noise)."""
    top = max(float(g.norm()) for g in ref_grads.values() if g is not None)
    worst, worst_name = 0.0, None
    for n, g in named_grads.items():
        r = ref_grads.get(n)
        if r is None and g is None:
            continue
        assert r is not None and g is not None, "gradient presence differs for %s" % n
        e = float((g.detach().float().cpu() - r.detach().float().cpu()).norm()) / max(float(r.norm()), floor_frac * top)
        if e > worst:
            worst, worst_name = e, n
    return worst, worst_name


def grad_errors(named_grads, ref_grads, floor_frac=1e-2):
    """per-parameter relative errors (same definition as grad_report) + the global relative L2 error over all
    gradients concatenated."""
    top = max(float(g.norm()) for g in ref_grads.values() if g is not None)
    errs, num, den = {}, 0.0, 0.0
    for n, g in named_grads.items():
        r = ref_grads.get(n)
        if r is None and g is None:
            continue
        assert r is not None and g is not None, "gradient presence differs for %s" % n
        d = float((g.detach().float().cpu() - r.detach().float().cpu()).norm())
        errs[n] = d / max(float(r.norm()), floor_frac * top)
        num += d * d
        den += float(r.norm()) ** 2
    return errs, (num / max(den, 1e-60)) ** 0.5
