"""AdamW driver for the step loop: the same `torch._fused_adamw_` multi-tensor kernel as
`torch.optim.AdamW(fused=True)`, with the per-step Python bookkeeping removed.

torch's optimizer walks every parameter of every group on each step (state lookup, list building, grouping by device
and dtype: ~1.2 ms of host time for the 430 parameters of this model), which matters on a path whose whole step takes
~15 ms and is host-bound.  Here the parameter / state lists are cached per "gradient signature" (the set of parameters
that received a gradient: one signature per pre-training task), so a step is two foreach calls.

Semantics follow torch.optim.AdamW: decoupled weight decay, bias correction from a per-parameter step counter,
parameters without a gradient are skipped entirely (no decay, no moment update), state created lazily as zeros.
The optimizer of the reference (pretrain_src/optim/) is out of scope; this is bench / training-loop plumbing.
"""
import torch


class FusedAdamW:
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        self.params = [p for p in params if p.requires_grad]
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.state = {}       # index -> (exp_avg, exp_avg_sq, step)
        self._lists = {}      # signature -> (params, exp_avgs, exp_avg_sqs, steps)

    def _state(self, i):
        st = self.state.get(i)
        if st is None:
            p = self.params[i]
            st = (torch.zeros_like(p, memory_format=torch.preserve_format),
                  torch.zeros_like(p, memory_format=torch.preserve_format),
                  torch.zeros((), dtype=torch.float32, device=p.device))
            self.state[i] = st
        return st

    @torch.no_grad()
    def step(self, set_to_none=True):
        sig = tuple(i for i, p in enumerate(self.params) if p.grad is not None)
        if not sig:
            return
        lists = self._lists.get(sig)
        if lists is None:
            sts = [self._state(i) for i in sig]
            lists = ([self.params[i] for i in sig], [s[0] for s in sts], [s[1] for s in sts], [s[2] for s in sts])
            self._lists[sig] = lists
        plist, m, v, steps = lists
        grads = [p.grad for p in plist]
        torch._foreach_add_(steps, 1)
        torch._fused_adamw_(plist, grads, m, v, [], steps, lr=self.lr, beta1=self.betas[0], beta2=self.betas[1],
                            weight_decay=self.weight_decay, eps=self.eps, amsgrad=False, maximize=False)
        if set_to_none:
            for p in plist:
                p.grad = None

    def zero_grad(self, set_to_none=True):
        for p in self.params:
            if p.grad is not None:
                if set_to_none:
                    p.grad = None
                else:
                    p.grad.zero_()
