"""TEST INFRASTRUCTURE ONLY -- imports the *unmodified* reference modules from /root/reference.

Only usable in the build container (the GPU box has no /root/reference).  It validates the torch
restatement in `oracle/bevbert_ref.py` and generates the golden fixtures under tests/golden/
(tests/golden/make_golden.py).  Nothing in the product package imports this file.

The five shims are the ones listed in SURVEY.md 8(c); they patch library drift (transformers 5.x,
missing torch_scatter, CPU-only container), never the reference's arithmetic.
"""
import importlib
import math
import os
import sys
import types

import torch

REF_ROOT = os.environ.get("BEVBERT_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "pretrain_src", "model"))


def _scatter_mean(src, index, dim=0, dim_size=None):
    """torch-scatter 2.0.9 scatter_mean restated: scatter_sum, count (clamped to >=1), true divide.
    (third-party dependency, environment.yaml:246; call sites pretrain_src/model/bev_utils.py:407,417)"""
    assert dim == 0
    out = torch.zeros((dim_size,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    out.index_add_(0, index, src)
    cnt = torch.zeros(dim_size, dtype=src.dtype, device=src.device)
    cnt.index_add_(0, index, torch.ones(index.shape[0], dtype=src.dtype, device=src.device))
    cnt = cnt.clamp(min=1)
    return out / cnt.view(-1, *([1] * (src.dim() - 1)))


def load_pretrain_modules(bev_dim: int = 21, bev_res: float = 0.5):
    """Returns (vilmodel, pretrain_cmt) reference modules ready to run on CPU."""
    if not available():
        raise RuntimeError("reference tree not found at %s" % REF_ROOT)
    sys.dont_write_bytecode = True
    src = os.path.join(REF_ROOT, "pretrain_src")
    if src not in sys.path:
        sys.path.insert(0, src)
    if "torch_scatter" not in sys.modules:
        stub = types.ModuleType("torch_scatter")
        stub.scatter_mean = _scatter_mean
        stub.scatter_max = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError("scatter_max stub"))
        sys.modules["torch_scatter"] = stub
    vilmodel = importlib.import_module("model.vilmodel")
    vilmodel.BertPreTrainedModel.init_weights = lambda self: None
    vilmodel.BertPreTrainedModel._tie_or_clone_weights = lambda self, out, inp: setattr(out, "weight", inp.weight)
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
    pretrain_cmt = importlib.import_module("model.pretrain_cmt")
    bev_utils = importlib.import_module("model.bev_utils")
    pretrain_cmt.BEV_DIM = bev_dim
    pretrain_cmt.BEV_RES = bev_res
    dev = torch.device("cuda" if torch.cuda.is_available() else "cpu")

    def build_projector():
        projector = bev_utils.PointCloud(math.radians(90), 1, feature_map_height=14, feature_map_width=14,
                                         map_dim=pretrain_cmt.BEV_DIM, map_res=pretrain_cmt.BEV_RES,
                                         world_shift_origin=torch.zeros(3, device=dev), z_clip_threshold=0.5,
                                         device=dev)
        bev_pos = bev_utils.bevpos_polar(pretrain_cmt.BEV_DIM).to(dev)
        return projector, bev_pos.reshape(pretrain_cmt.BEV_DIM * pretrain_cmt.BEV_DIM, 3)[None, :, :]

    pretrain_cmt.build_projector = build_projector
    return vilmodel, pretrain_cmt


def build_reference_model(config, state_dict=None):
    """Constructs the reference GlocalTextPathCMTPreTraining on CPU and loads `state_dict` (fp32)."""
    _, pretrain_cmt = load_pretrain_modules(config.bev_dim, getattr(config, "bev_res", 0.5))
    model = pretrain_cmt.GlocalTextPathCMTPreTraining(config)
    if state_dict is not None:
        missing, unexpected = model.load_state_dict(state_dict, strict=False)
        missing = [k for k in missing if not k.endswith("decoder.weight")]
        assert not missing and not unexpected, (missing, unexpected)
        model.tie_weights()
    return model


def load_nav_module():
    """map_nav_src/models/vilmodel.py (GlocalTextPathNavCMT) of the unmodified reference."""
    if not available():
        raise RuntimeError("reference tree not found at %s" % REF_ROOT)
    sys.dont_write_bytecode = True
    src = os.path.join(REF_ROOT, "map_nav_src")
    if src not in sys.path:
        sys.path.insert(0, src)
    nav = importlib.import_module("models.vilmodel")
    nav.BertPreTrainedModel.init_weights = lambda self: None
    return nav
