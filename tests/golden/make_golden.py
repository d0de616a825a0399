"""Generates tests/golden/*.pt by running the UNMODIFIED reference (/root/reference, via oracle/ref_shim.py)
on seeded weights and synthetic batches.  Run in the build container only:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Weights (synth.det_init_) and batches (synth.make_batch) are regenerated bit-identically from their seeds by
the tests, so only reference OUTPUTS are stored: per-item losses, logits, per-parameter gradient norms and a
few full gradients.  The reference has no tests or golden vectors of its own."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from bevbert_b200 import synth  # noqa: E402
from bevbert_b200.config import make_config  # noqa: E402
from bevbert_b200.model.pretrain_cmt import GlocalTextPathCMTPreTraining  # noqa: E402
from helpers import small_config, small_synth  # noqa: E402
from oracle import ref_shim  # noqa: E402

FULL_GRADS = ["bert.embeddings.LayerNorm.weight", "bert.lang_encoder.layer.0.attention.self.query.bias",
              "bert.local_encoder.encoder.x_layers.0.visual_attention.att.query.bias",
              "bert.global_encoder.sprel_linear.weight", "bert.global_encoder.sprel_linear.bias",
              "bert.img_embeddings.loc_linear.weight", "bert.local_encoder.bev_pos_embeddings.0.weight"]

CASES = {
    "small_r2r": (lambda: small_config(), lambda: small_synth(), ["mlm", "sap", "masksem"]),
    "small_reverie": (lambda: small_config(obj_feat_size=768, obj_prob_size=100, pretrain_tasks=["mlm", "mrc", "sap", "og"]),
                      lambda: small_synth(obj_feat_size=768, obj_max=5, obj_prob_size=100), ["mlm", "mrc", "sap", "og"]),
    "config1_full_depth": (lambda: make_config(bev_dim=11, bev_res=1.0, hidden_dropout_prob=0.0,
                                               attention_probs_dropout_prob=0.0, feat_dropout=0.0),
                           lambda: small_synth(), ["mlm", "sap", "masksem"]),
}


def main():
    out_dir = os.path.dirname(os.path.abspath(__file__))
    for name, (mk_cfg, mk_synth, tasks) in CASES.items():
        cfg, scfg = mk_cfg(), mk_synth()
        ours = synth.det_init_(GlocalTextPathCMTPreTraining(cfg), seed=3)
        sd = {k: v.detach().clone() for k, v in ours.state_dict().items()}
        ref = ref_shim.build_reference_model(cfg, sd).train()
        for m in ref.modules():
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0
            if isinstance(m, torch.nn.MultiheadAttention):
                m.dropout = 0.0
        assert set(ref.state_dict().keys()) == set(sd.keys()), "state_dict key sets differ"
        for k, v in ref.state_dict().items():
            assert v.shape == sd[k].shape, k
        gold = {"state_dict_keys": sorted(sd.keys()), "shapes": {k: tuple(v.shape) for k, v in sd.items()}}
        for task in tasks:
            b = synth.make_batch(scfg, seed=7, task=task)
            ref.zero_grad()
            loss = ref(synth.clone_batch(b), task, compute_loss=True)
            loss.mean().backward()
            with torch.no_grad():
                logits = ref(synth.clone_batch(b), task, compute_loss=False)
            logits = logits if isinstance(logits, tuple) else (logits,)
            grads = {n: p.grad for n, p in ref.named_parameters() if p.grad is not None}
            gold[task] = {
                "loss": loss.detach().clone(),
                "logits": [t.detach().clone() for t in logits[:3]],
                "grad_norms": {n: float(g.norm()) for n, g in grads.items()},
                "grads": {n: grads[n].clone() for n in FULL_GRADS if n in grads},
            }
            print(name, task, "loss mean %.6f" % float(loss.mean()), "n_grads", len(grads))
        torch.save(gold, os.path.join(out_dir, name + ".pt"))


if __name__ == "__main__":
    main()
