"""Pins the oracle to the real reference (only where /root/reference exists, i.e. the build container):
the unmodified reference modules are imported through oracle/ref_shim.py and run on the same weights
and batches as the restatement."""
import pytest
import torch

from bevbert_b200 import synth
from bevbert_b200.model.pretrain_cmt import GlocalTextPathCMTPreTraining
from helpers import small_config, small_synth
from oracle import bevbert_ref as R
from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present on this machine")


def _ref_model(cfg, sd):
    ref = ref_shim.build_reference_model(cfg, sd).train()
    for m in ref.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
        if isinstance(m, torch.nn.MultiheadAttention):
            m.dropout = 0.0
    return ref


@pytest.mark.parametrize("task", ["mlm", "sap", "masksem"])
def test_restatement_equals_reference(task):
    cfg, scfg = small_config(), small_synth(ragged_txt=True)
    sd = {k: v.detach().clone() for k, v in synth.det_init_(GlocalTextPathCMTPreTraining(cfg), seed=5).state_dict().items()}
    ref = _ref_model(cfg, sd)
    b = synth.make_batch(scfg, seed=21, task=task)
    want = ref(synth.clone_batch(b), task, compute_loss=True)
    got = R.forward(sd, synth.clone_batch(b), task, R.OracleConfig(cfg))
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-6), float((got - want).abs().max())


@pytest.mark.parametrize("case", ["longest_trajectory", "empty_bev"])
def test_restatement_equals_reference_on_edge_cases(case):
    """SURVEY section 4 property cases, pinned on the unmodified reference: 21 panoramas per trajectory (TRAIN_MAX_STEP + 1,
    dataset.py:185-187) and depth maps that are all zero (every point dropped, all BEV cells empty)."""
    cfg = small_config()
    scfg = small_synth(pano_min=21, pano_max=21) if case == "longest_trajectory" else small_synth(depth_zero_frac=1.0)
    sd = {k: v.detach().clone() for k, v in synth.det_init_(GlocalTextPathCMTPreTraining(cfg), seed=5).state_dict().items()}
    ref = _ref_model(cfg, sd)
    for task in ("sap", "mlm"):
        b = synth.make_batch(scfg, seed=23, task=task)
        want = ref(synth.clone_batch(b), task, compute_loss=True)
        got = R.forward(sd, synth.clone_batch(b), task, R.OracleConfig(cfg))
        assert torch.isfinite(want).all()
        assert torch.allclose(got, want, rtol=1e-5, atol=1e-6), (task, float((got - want).abs().max()))


def test_bev_projection_matches_reference_bit_for_bit():
    """lift + project_bev of the reference (torch bmm / matmul + scatter_mean stub) vs the fixed-order oracle:
    cell indices, pooled features, semantic maps and masks."""
    _, pretrain_cmt = ref_shim.load_pretrain_modules(21, 0.5)
    projector, _ = pretrain_cmt.build_projector()
    mismatched = 0
    total = 0
    for seed in (1, 2, 3):
        b = synth.make_batch(synth.SynthConfig(batch_size=4), seed=seed)
        bs = 4
        depths = (b["depths"] * 10).reshape(-1, 1, 14, 14)
        pc, mask = projector.forward(depths, b["T_c2w"].reshape(-1, 4, 4))
        pc = pc.reshape(bs, -1, 3) - b["S_w2c"]
        pc1 = torch.cat([pc, torch.ones(bs, pc.shape[1], 1)], -1) @ b["T_w2c"].squeeze(1).transpose(1, 2)
        bev, ob, sem, sem_mask = projector.project_bev(pc1[:, :, :3], mask.reshape(bs, -1), b["rgbs"].reshape(bs, -1, 768),
                                                       b["sems"].reshape(bs, -1, 40))
        ob_ = dict(b)
        R.lift_splat(ob_, 21, 0.5)
        same = torch.equal(ob_["bev_fts"], bev.reshape(bs, -1, 768))
        # the reference leaves the order of the 4-term dot products to bmm/matmul; count index flips instead of failing
        opc, nod = R.lift_points(b["depths"], b["T_c2w"], b["S_w2c"], b["T_w2c"])
        ref_idx = R.cell_index(pc1[:, :, :3], mask.reshape(bs, -1), 21, 0.5)
        mismatched += int((ref_idx != ob_["bev_cell_idx"]).sum())
        total += ref_idx.numel()
        if same:
            assert torch.equal(ob_["bev_sems"], sem.reshape(bs, -1, 40))
            assert torch.equal(ob_["bev_sem_masks"], sem_mask.reshape(bs, -1))
            assert torch.equal(ob_["bev_ob_masks"], ob.reshape(bs, -1))
    assert mismatched <= total * 1e-5, "cell-index disagreement with the reference: %d of %d" % (mismatched, total)
