"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's optimizer step for the parity tests of
bevbert_b200.optim.AdamW (csrc/optim.cu).  Pinned to the unmodified reference by tests/test_oracle_vs_reference.py
(`pretrain_src/optim/adamw.py` imported read-only in the build container).

Follows pretrain_src/optim/adamw.py:53-112 (AdamW.step) and torch.nn.utils.clip_grad_norm_ as called at
pretrain_src/train_r2r.py:295-300.  Plain torch fp32 ops on CPU, one tensor at a time, like the reference.
"""
import math

import torch


def clip_grad_norm_(grads, max_norm):
    """total L2 norm over all gradients; every gradient *= min(1, max_norm / (norm + 1e-6)).  -> norm"""
    total = torch.sqrt(sum((g.detach().double() ** 2).sum() for g in grads)).float()
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    for g in grads:
        g.mul_(coef)
    return total


class AdamWRef:
    """state per parameter index: step, exp_avg, exp_avg_sq (adamw.py:76-83)."""

    def __init__(self, params, lr, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.0, correct_bias=True):
        self.params = list(params)
        self.lr, self.betas, self.eps, self.correct_bias = lr, betas, eps, correct_bias
        self.wd = weight_decay if isinstance(weight_decay, (list, tuple)) else [weight_decay] * len(self.params)
        self.state = [None] * len(self.params)

    @torch.no_grad()
    def step(self, grads):
        b1, b2 = self.betas
        for i, (p, g) in enumerate(zip(self.params, grads)):
            if g is None:                                            # adamw.py:67-68
                continue
            if self.state[i] is None:
                self.state[i] = [0, torch.zeros_like(p), torch.zeros_like(p)]
            st = self.state[i]
            st[0] += 1
            st[1].mul_(b1).add_(g, alpha=1.0 - b1)                   # adamw.py:89
            st[2].mul_(b2).addcmul_(g, g, value=1.0 - b2)            # adamw.py:90
            denom = st[2].sqrt().add_(self.eps)                      # adamw.py:91
            step_size = self.lr
            if self.correct_bias:                                    # adamw.py:94-97
                step_size = step_size * math.sqrt(1.0 - b2 ** st[0]) / (1.0 - b1 ** st[0])
            p.addcdiv_(st[1], denom, value=-step_size)               # adamw.py:99
            if self.wd[i] > 0.0:                                     # adamw.py:110
                p.add_(p, alpha=-self.lr * self.wd[i])
