"""Agent-side per-step API on the CUDA kernels (GlocalTextPathNavCMT.forward(mode, batch),
map_nav_src/models/vilmodel.py:744-912): language / panorama / navigation modes against the fp32 oracle
(oracle.nav_forward is pinned to the unmodified reference class by tests/test_nav_api_cpu.py), REVERIE object tokens
included, plus the per-step latency at the fine-tuning batch size (B=4, scripts/ft_r2r.bash:40)."""
import time

import pytest
import torch

from bevbert_b200 import synth
from bevbert_b200.model.nav_vilmodel import GlocalTextPathNavCMT
from helpers import rel_l2
from oracle import bevbert_ref as R
from test_nav_api_cpu import nav_batches, nav_config

pytestmark = pytest.mark.gpu


def _to(d, dev):
    return {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in d.items()}


@pytest.mark.parametrize("obj", [0, 5])
def test_nav_modes_on_gpu_match_oracle(obj):
    kw = dict(obj_feat_size=768) if obj else {}
    cfg = nav_config(**kw)
    model = synth.det_init_(GlocalTextPathNavCMT(cfg), seed=4).cuda().eval()
    sd = {k: v.detach().float().cpu().clone() for k, v in model.state_dict().items()}
    ocfg = R.OracleConfig(cfg)
    lang, pano, nav = nav_batches(obj=obj)
    with torch.no_grad():
        t = model("language", _to(lang, "cuda"))
        rt_ = R.nav_forward(sd, "language", lang, ocfg)
        assert rel_l2(t, rt_) < 1e-2
        pe, pm = model("panorama", _to(pano, "cuda"))
        rpe, rpm = R.nav_forward(sd, "panorama", pano, ocfg)
        assert torch.equal(pm.cpu(), rpm) and rel_l2(pe, rpe) < 1e-2
        nav["txt_embeds"] = rt_
        ref = R.nav_forward(sd, "navigation", nav, ocfg)
        out = model("navigation", _to(nav, "cuda"))
    for k in ("gmap_embeds", "global_logits", "local_logits", "fused_logits"):
        fin = torch.isfinite(ref[k])
        assert torch.equal(torch.isfinite(out[k].cpu()), fin), k
        assert rel_l2(out[k].cpu()[fin], ref[k][fin]) < 2e-2, (k, rel_l2(out[k].cpu()[fin], ref[k][fin]))


def test_nav_step_gradients_and_latency():
    """Fine-tuning shape (B=4, 21x21 BEV, full depth): the navigation step is differentiable on the CUDA path and its
    latency is reported (printed; `pytest -s`)."""
    from bevbert_b200.config import make_config
    cfg = make_config(hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, feat_dropout=0.0, fix_lang_embedding=False,
                      fix_pano_embedding=False, fix_local_branch=False)
    model = synth.det_init_(GlocalTextPathNavCMT(cfg), seed=4).cuda().train()
    lang, pano, nav = nav_batches(B=3, G=12, K=4, D=21)
    lang, pano, nav = _to(lang, "cuda"), _to(pano, "cuda"), _to(nav, "cuda")
    t = model("language", lang)
    nav["txt_embeds"] = t
    out = model("navigation", nav)
    fl = out["fused_logits"]
    fl[torch.isfinite(fl)].sum().backward()
    g = model.bert.lang_encoder.layer[0].attention.self.query.weight.grad if hasattr(model, "bert") else None
    assert g is None or torch.isfinite(g).all()
    assert model.global_sap_head.net[0].weight.grad is not None
    model.eval()
    with torch.no_grad():
        for _ in range(3):
            model("navigation", nav)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            model("navigation", nav)
        torch.cuda.synchronize()
    print("navigation step (B=3, 21x21, G=12, full depth): %.2f ms" % ((time.perf_counter() - t0) * 100))
