"""Model-level parity on the GPU: GlocalTextPathCMTPreTraining on the sm_100a kernels (bf16 activations)
against the fp32 CPU oracle, same seeded weights and synthetic batches; losses, logits and every parameter
gradient.  Tolerances (north_star): 1e-2 relative for the bf16 path."""
import os

import pytest
import torch

from bevbert_b200 import synth
from bevbert_b200.config import make_config
from bevbert_b200.model.pretrain_cmt import GlocalTextPathCMTPreTraining
from helpers import grad_report, rel_l2, small_config, small_synth
from oracle import bevbert_ref as R

pytestmark = pytest.mark.gpu


def _compare(task, cfg, scfg, seed=7, loss_tol=1e-2, grad_tol=3e-2):
    model = synth.det_init_(GlocalTextPathCMTPreTraining(cfg), seed=3).cuda().train()
    sd = {k: v.detach().float().cpu().clone().requires_grad_(True) for k, v in model.state_dict().items()}
    b = synth.make_batch(scfg, seed=seed, task=task)
    out = model(synth.batch_to(b, "cuda"), task, compute_loss=True)
    out.mean().backward()
    ref = R.forward(sd, synth.clone_batch(b), task, R.OracleConfig(cfg))
    ref.mean().backward()
    assert out.shape == ref.shape
    le = rel_l2(out, ref)
    worst, name = grad_report({n: p.grad for n, p in model.named_parameters()},
                              {n: sd[n].grad for n, _ in model.named_parameters()}, floor_frac=1e-2)
    print("task=%s loss rel-L2=%.3e worst grad rel err=%.3e (%s)" % (task, le, worst, name))
    assert le < loss_tol, le
    assert worst < grad_tol, (worst, name)
    return le, worst


@pytest.mark.parametrize("task", ["mlm", "sap", "masksem"])
def test_small_config_matches_oracle(task):
    _compare(task, small_config(), small_synth())


@pytest.mark.parametrize("task", ["mlm", "mrc", "sap", "og"])
def test_reverie_object_tokens(task):
    cfg = small_config(obj_feat_size=768, obj_prob_size=100, pretrain_tasks=["mlm", "mrc", "sap", "og"])
    _compare(task, cfg, small_synth(obj_feat_size=768, obj_max=5, obj_prob_size=100))


@pytest.mark.parametrize("task", ["mlm", "sap", "masksem"])
def test_baseline_config1_full_depth(task):
    """BASELINE.json configs[0]: B=2, 80 tokens, 36 views x 768, 11x11 BEV, 8 topo nodes, full 9/2/4/4 layers."""
    cfg = make_config(bev_dim=11, bev_res=1.0, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0,
                      feat_dropout=0.0)
    _compare(task, cfg, small_synth(), grad_tol=5e-2)


def test_bev_inputs_exact_and_logits():
    """The BEV tensors handed to the encoder are bit-exact; compute_loss=False returns the logits tuple."""
    cfg = small_config()
    model = synth.det_init_(GlocalTextPathCMTPreTraining(cfg), seed=3).cuda().eval()
    b = synth.make_batch(small_synth(), seed=11, task="sap")
    bb = model.lift_splat(dict(synth.batch_to(b, "cuda")))
    rb = R.lift_splat(dict(synth.clone_batch(b)), cfg.bev_dim, cfg.bev_res)
    assert torch.equal(bb["bev_cell_idx"].cpu().long(), rb["bev_cell_idx"])
    assert torch.equal(bb["bev_fts"].cpu(), rb["bev_fts"])
    assert torch.equal(bb["bev_sems"].cpu(), rb["bev_sems"])
    assert torch.equal(bb["bev_sem_masks"].cpu(), rb["bev_sem_masks"])
    assert torch.equal(bb["bev_pos_fts"].cpu(), rb["bev_pos_fts"])
    sd = {k: v.detach().float().cpu() for k, v in model.state_dict().items()}
    with torch.no_grad():
        gl, ll, fl, _, _ = model(synth.batch_to(b, "cuda"), "sap", compute_loss=False)
        rgl, rll, rfl, _, _ = R.forward(sd, synth.clone_batch(b), "sap", R.OracleConfig(cfg), compute_loss=False)
    for a, r in ((gl, rgl), (ll, rll), (fl, rfl)):
        fin = torch.isfinite(r)
        assert torch.equal(torch.isfinite(a.cpu()), fin)
        assert rel_l2(a.cpu()[fin], r[fin]) < 2e-2


def test_dropout_training_step_runs_and_is_reproducible():
    cfg = small_config(hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1, feat_dropout=0.4)
    model = synth.det_init_(GlocalTextPathCMTPreTraining(cfg), seed=3).cuda().train()
    b = synth.batch_to(synth.make_batch(small_synth(), seed=5, task="sap"), "cuda")
    model.rt.calls = 10
    l1 = model(b, "sap").mean()
    l1.backward()
    g1 = model.bert.lang_encoder.layer[0].attention.self.query.weight.grad.clone()
    model.zero_grad()
    model.rt.calls = 10
    l2 = model(b, "sap").mean()
    l2.backward()
    assert torch.isfinite(l1) and float(l1) == float(l2)
    assert torch.equal(g1, model.bert.lang_encoder.layer[0].attention.self.query.weight.grad) or \
        rel_l2(model.bert.lang_encoder.layer[0].attention.self.query.weight.grad, g1) < 1e-3   # split-K atomics
    l3 = model(b, "sap").mean()
    assert float(l3) != float(l1)
