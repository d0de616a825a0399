"""bevbert_b200.optim.FusedAdamW (cached parameter lists around torch._fused_adamw_) takes exactly the steps of
torch.optim.AdamW(fused=True), including parameters that get no gradient in some steps."""
import torch

from bevbert_b200.optim import FusedAdamW


def _params():
    torch.manual_seed(1)
    return [torch.nn.Parameter(torch.randn(7, 5)), torch.nn.Parameter(torch.randn(11)), torch.nn.Parameter(torch.randn(3, 3))]


def test_fused_adamw_matches_torch():
    a, b = _params(), _params()
    oa = torch.optim.AdamW(a, lr=1e-2, weight_decay=0.05, fused=True)
    ob = FusedAdamW(b, lr=1e-2, weight_decay=0.05)
    g = torch.Generator().manual_seed(2)
    for step in range(7):
        grads = [torch.randn(p.shape, generator=g) for p in a]
        skip = step % 3                                  # one parameter without a gradient per step
        for i, (pa, pb) in enumerate(zip(a, b)):
            pa.grad = None if i == skip else grads[i].clone()
            pb.grad = None if i == skip else grads[i].clone()
        oa.step()
        oa.zero_grad(set_to_none=True)
        ob.step()
        assert all(p.grad is None for p in b)
    for pa, pb in zip(a, b):
        assert torch.equal(pa.detach(), pb.detach())
