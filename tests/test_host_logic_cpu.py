"""Host-side logic (block forward/backward composition, index builders, model glue) on CPU: the kernels are
replaced by their torch emulation (tests/emu_kernels.py) and the model is compared with the fp32 oracle."""
import pytest
import torch

from bevbert_b200 import synth
from bevbert_b200.model.pretrain_cmt import GlocalTextPathCMTPreTraining
from helpers import grad_report, small_config, small_synth
from oracle import bevbert_ref as R


def _run(task, cfg, scfg, seed=7):
    model = synth.det_init_(GlocalTextPathCMTPreTraining(cfg), seed=3).train()
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in model.state_dict().items()}
    b = synth.make_batch(scfg, seed=seed, task=task)
    out = model(synth.clone_batch(b), task, compute_loss=True)
    out.mean().backward()
    ref = R.forward(sd, synth.clone_batch(b), task, R.OracleConfig(cfg))
    ref.mean().backward()
    assert out.shape == ref.shape
    assert torch.allclose(out, ref, rtol=1e-4, atol=1e-5), float((out - ref).abs().max())
    worst, name = grad_report({n: p.grad for n, p in model.named_parameters()},
                              {n: sd[n].grad for n, _ in model.named_parameters()})
    assert worst < 1e-3, (worst, name)


@pytest.mark.parametrize("task", ["mlm", "sap", "masksem"])
def test_r2r_tasks_match_oracle(emu, task):
    _run(task, small_config(), small_synth())


def test_ragged_text_and_sem_modes(emu):
    for mode in ("sattn", "embed"):
        _run("masksem", small_config(sem_pred_token=mode), small_synth(ragged_txt=True))
    _run("sem", small_config(pretrain_tasks=["mlm", "sap", "sem"]), small_synth(ragged_txt=True))


@pytest.mark.parametrize("task", ["mlm", "mrc", "sap", "og"])
def test_reverie_object_tokens(emu, task):
    cfg = small_config(obj_feat_size=768, obj_prob_size=100, pretrain_tasks=["mlm", "mrc", "sap", "og"])
    _run(task, cfg, small_synth(obj_feat_size=768, obj_max=5, obj_prob_size=100))


def test_product_refuses_cpu_tensors():
    """No CPU fallback: the real kernel wrappers reject CPU tensors / a missing CUDA runtime."""
    import bevbert_b200.kernels as K
    with pytest.raises(Exception):
        K.cast_to_act(torch.zeros(8))


def test_prepare_batch_gives_identical_results(emu):
    """Collate-time host index building (model.ops.prepare_batch) changes nothing but where the host work happens."""
    from bevbert_b200.model.ops import prepare_batch
    cfg, scfg = small_config(), small_synth()
    model = synth.det_init_(GlocalTextPathCMTPreTraining(cfg), seed=3).train()
    for task in ("sap", "mlm"):
        b = synth.make_batch(scfg, seed=13, task=task)
        a = model(synth.clone_batch(b), task)
        c = model(prepare_batch(synth.clone_batch(b)), task)
        assert torch.equal(a, c)


def test_step_arena_gives_identical_gradients(emu):
    """The per-step zero arena (blocks.ARENA: one fill per step for all accumulate-into gradient buffers) is used from
    the second step of a task on and changes no gradient; gradients of an earlier step stay valid after the next."""
    from bevbert_b200 import blocks
    cfg, scfg = small_config(), small_synth()
    model = synth.det_init_(GlocalTextPathCMTPreTraining(cfg), seed=3).train()
    b = synth.make_batch(scfg, seed=13, task="sap")
    grads = []
    for step in range(3):
        model.zero_grad(set_to_none=True)
        model(synth.clone_batch(b), "sap").mean().backward()
        if step:
            assert blocks.ARENA.buf is not None and blocks.ARENA.off > 0
        grads.append({n: p.grad for n, p in model.named_parameters() if p.grad is not None})
    for n, g in grads[0].items():
        assert torch.equal(g, grads[1][n]) and torch.equal(g, grads[2][n]), n


def test_direct_param_grads_equal_autograd_accumulation(emu):
    """parallel.direct_param_grads(): blocks write p.grad themselves; same gradients (shared word embeddings
    accumulate), also over two accumulated backward passes."""
    from bevbert_b200 import parallel
    cfg, scfg = small_config(), small_synth()
    b = synth.make_batch(scfg, seed=13, task="mlm")
    grads = {}
    for mode in (False, True):
        parallel.direct_param_grads(mode)
        try:
            model = synth.det_init_(GlocalTextPathCMTPreTraining(cfg), seed=3).train()
            for _ in range(2):
                model(synth.clone_batch(b), "mlm").mean().backward()
            grads[mode] = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
        finally:
            parallel.direct_param_grads(False)
    assert set(grads[False]) == set(grads[True])
    for n, g in grads[False].items():
        assert torch.allclose(g, grads[True][n], rtol=1e-6, atol=1e-7), n


def test_forward_sees_weights_updated_behind_the_version_counter(emu):
    """Optimizers write `p.data` (torch._fused_adamw_, the reference's `p.data.add_`, optim/adamw.py:103-110) without
    bumping `p._version`, so the operand shadows kept by blocks.WeightCache cannot rely on versions (round-1 advisor
    finding): a training forward after such an update must use the new weights, here against the oracle on CPU."""
    cfg, scfg = small_config(), small_synth()
    model = synth.det_init_(GlocalTextPathCMTPreTraining(cfg), seed=3).train()
    b = synth.make_batch(scfg, seed=7, task="mlm")
    out0 = model(synth.clone_batch(b), "mlm").detach()
    touched = [model.bert.lang_encoder.layer[0].attention.self.query.weight,        # stacked Q|K|V shadow
               model.bert.lang_encoder.layer[0].intermediate.dense.weight,          # plain shadow
               model.bert.lang_encoder.layer[0].output.LayerNorm.weight,            # fp32 vector entry
               model.bert.lang_encoder.layer[0].attention.self.key.bias]            # packed bias
    versions = [p._version for p in touched]
    with torch.no_grad():
        for p in touched:
            p.data.mul_(1.25).add_(0.01)
    assert [p._version for p in touched] == versions      # the premise: nothing for a version check to see
    out1 = model(synth.clone_batch(b), "mlm").detach()
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    ref = R.forward(sd, synth.clone_batch(b), "mlm", R.OracleConfig(cfg)).detach()
    assert not torch.allclose(out1, out0, rtol=1e-3, atol=1e-4)
    assert torch.allclose(out1, ref, rtol=1e-4, atol=1e-5), float((out1 - ref).abs().max())


def test_edge_cases_longest_trajectory_and_empty_bev(emu):
    """SURVEY section 4 property cases on the host logic: the longest trajectory the loaders produce (TRAIN_MAX_STEP + 1 = 21
    panoramas, dataset.py:185-187), and a batch whose depth maps are all zero (every point dropped: an empty BEV, all
    cells `ob_mask` False, `bev_utils.py:393-430`)."""
    _run("sap", small_config(), small_synth(pano_min=21, pano_max=21))
    for task in ("sap", "mlm"):
        _run(task, small_config(), small_synth(depth_zero_frac=1.0))


def test_reverie_batch_contains_viewpoints_without_objects(emu):
    """zero-object viewpoints (and a last viewpoint without objects -> label -100, pretrain_cmt.py:380-389) are part of the
    REVERIE parity cases above, not an accident of the seed"""
    scfg = small_synth(obj_feat_size=768, obj_max=5, obj_prob_size=100, batch_size=4)
    cfg = small_config(obj_feat_size=768, obj_prob_size=100, pretrain_tasks=["mlm", "mrc", "sap", "og"])
    for seed in range(7, 40):
        b = synth.make_batch(scfg, seed=seed, task="og")
        if int((b["traj_vp_obj_lens"] == 0).sum()) > 0 and int((b["obj_labels"] == -100).sum()) > 0 \
                and int((b["obj_labels"] >= 0).sum()) > 0:
            break
    else:
        raise AssertionError("no seed with an object-free last viewpoint")
    _run("og", cfg, scfg, seed=seed)
