"""Host half of bevbert_b200.optim.AdamW (no kernel launch, runs without a GPU): the launch table the update kernel
reads -- pointers, chunk offsets, per-parameter step counters, bias-corrected step sizes (`optim/adamw.py:93-99`),
decoupled decay `lr * wd` (`:110`), parameters without a gradient skipped (`:63-64`), table rebuilt when a parameter's
storage moves."""
import math

import torch

from bevbert_b200 import _lib
from bevbert_b200.optim import AdamW, build_param_groups


class _Net(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.dense = torch.nn.Linear(7, 5)
        self.LayerNorm = torch.nn.LayerNorm(5)
        self.unused = torch.nn.Linear(3, 3)


def _opt():
    torch.manual_seed(0)
    net = _Net()
    opt = AdamW(build_param_groups(net, 0.01), lr=1e-3, betas=(0.9, 0.98), eps=1e-6, max_grad_norm=5.0)
    for n, p in net.named_parameters():
        if not n.startswith("unused"):
            p.grad = torch.randn_like(p)
    return net, opt


def test_param_groups_follow_the_reference_no_decay_rule():
    net = _Net()
    g = build_param_groups(net, 0.01)
    decay = {id(p) for p in g[0]["params"]}
    assert id(net.dense.weight) in decay and id(net.unused.weight) in decay
    assert all(id(p) not in decay for p in (net.dense.bias, net.LayerNorm.weight, net.LayerNorm.bias, net.unused.bias))
    assert g[0]["weight_decay"] == 0.01 and g[1]["weight_decay"] == 0.0


def test_launch_table_rows_and_step_sizes():
    net, opt = _opt()
    sig = tuple(i for i, (_, p) in enumerate(opt.flat) if p.grad is not None)
    assert len(sig) == 4 and len(opt.flat) == 6          # the two `unused` parameters have no gradient: skipped entirely
    opt._init_state()
    tab, idx, stale, _ = opt._table(sig)
    chunk = int(_lib.load().bb_mt_chunk_elems())
    for t in (1, 2, 3):
        assert opt._fill(tab, idx, sig, True)
        c0 = 0
        for r, i in enumerate(idx):
            gi, p = opt.flat[i]
            row = tab.np[r]
            assert int(row["p"]) == p.data_ptr() and int(row["g"]) == p.grad.data_ptr()
            assert int(row["m"]) == opt.m[i].data_ptr() and int(row["v"]) == opt.v[i].data_ptr()
            assert int(row["n"]) == p.numel() and int(row["chunk0"]) == c0 and int(row["p16"]) == 0
            c0 += (p.numel() + chunk - 1) // chunk
            want = 1e-3 * math.sqrt(1.0 - 0.98 ** t) / (1.0 - 0.9 ** t)
            assert abs(float(row["step_size"]) - want) <= 1e-6 * want
            wd = opt.param_groups[gi]["weight_decay"]
            assert abs(float(row["decay"]) - 1e-3 * wd) <= 1e-12
        assert tab.chunks == c0
    assert [opt.steps[i] for i in sig] == [3] * 4 and all(opt.steps[i] == 0 for i in range(6) if i not in sig)
    # a scheduler writes group["lr"]: picked up by the next fill; correct_bias=False uses the plain lr (adamw.py:93-97)
    for g in opt.param_groups:
        g["lr"], g["correct_bias"] = 5e-4, False
    assert opt._fill(tab, idx, sig, True)
    assert all(abs(float(tab.np[r]["step_size"]) - 5e-4) < 1e-10 for r in range(len(idx)))


def test_table_is_rebuilt_when_a_parameter_moves():
    net, opt = _opt()
    sig = tuple(i for i, (_, p) in enumerate(opt.flat) if p.grad is not None)
    opt._init_state()
    tab, idx, _, _ = opt._table(sig)
    assert opt._fill(tab, idx, sig, True)
    moved = opt.flat[idx[2]][1]
    moved.data = moved.data.clone()                        # what module.to() / bias re-homing do
    before = list(opt.steps)
    assert not opt._fill(tab, idx, sig, True)              # detected; counters of the rows already advanced are rolled back
    assert opt.steps == before and sig not in opt._tables
    tab2, idx2, _, _ = opt._table(sig)
    assert tab2 is not tab and int(tab2.np[2]["p"]) == moved.data_ptr()
    assert opt._fill(tab2, idx2, sig, True)
