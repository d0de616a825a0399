"""Data-parallel gradient exchange for the hot path: one NCCL all-reduce over NVLink per step on a flat fp32
bucket of the gradients.

The model also runs unchanged under `torch.nn.parallel.DistributedDataParallel(find_unused_parameters=True)`
(what `pretrain_src/utils/misc.py:64-77` does; tests/test_ddp_cpu.py).  This helper is the lighter path used by
bench.py: DDP's per-parameter hooks and per-step unused-parameter graph search cost several milliseconds of host
time per step on a path that is already launch-bound, while every rank runs the same task in a step
(`data/loader.py:50-61` broadcasts the task id), so the set of parameters that received a gradient is identical
on all ranks and one coalesced all-reduce suffices.
"""
import os

import torch
import torch.distributed as dist

_CHECK_LAYOUT = os.environ.get("BEVBERT_CHECK_ARENA") == "1"


class FlatGradAllReduce:
    """all-reduce (average) the gradients of `params` with one NCCL call per step.

    From the second step of a task on, the blocks carve every gradient buffer from one zero-filled fp32 arena per step
    (blocks._ZeroArena), in the same order on every rank, so the arena itself is the flat bucket and is reduced in
    place.  Gradients that live elsewhere (the first step of a task, the few small heads that run on torch autograd)
    go through a second, small flat buffer."""

    def __init__(self, params, world_size=None):
        self.params = [p for p in params if p.requires_grad]
        self.world = world_size or (dist.get_world_size() if dist.is_initialized() else 1)
        self.buf = None

    @torch.no_grad()
    def __call__(self):
        if self.world == 1:
            return
        from . import blocks
        arena = blocks.ARENA.buf
        grads = [p.grad for p in self.params if p.grad is not None]
        if not grads:
            return
        inv = 1.0 / self.world
        if arena is not None and blocks.ARENA.off > 0:
            if _CHECK_LAYOUT:      # debug (BEVBERT_CHECK_ARENA=1): every rank must have carved the same number of floats
                n = torch.tensor([blocks.ARENA.off, -blocks.ARENA.off], dtype=torch.int64, device=arena.device)
                dist.all_reduce(n, op=dist.ReduceOp.MAX)
                if int(n[0]) != -int(n[1]):
                    raise RuntimeError("gradient arena layouts differ across ranks (%d..%d floats)" % (-int(n[1]), int(n[0])))
            base = arena.untyped_storage().data_ptr()
            rest = [g for g in grads if g.untyped_storage().data_ptr() != base]
            used = arena[:blocks.ARENA.off]
            used.mul_(inv)
            dist.all_reduce(used)
        else:
            rest = grads
        if not rest:
            return
        n = sum(g.numel() for g in rest)
        if self.buf is None or self.buf.numel() < n:
            self.buf = torch.empty(n, dtype=torch.float32, device=rest[0].device)
        flat = self.buf[:n]
        views = list(torch.split(flat, [g.numel() for g in rest]))
        torch._foreach_copy_(views, [g.reshape(-1) for g in rest])
        flat.mul_(inv)
        dist.all_reduce(flat)
        torch._foreach_copy_([g.view(-1) if g.is_contiguous() else g.reshape(-1) for g in rest], views)


def broadcast_parameters(model, src=0):
    """rank `src` -> all, once (the DDP constructor's initial broadcast, utils/misc.py:71-72)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    for t in list(model.parameters()) + list(model.buffers()):
        dist.broadcast(t.data, src)


def direct_param_grads(enable=True):
    """Let every encoder block write `p.grad` of its parameters itself at the end of its backward instead of returning
    the gradients to autograd (one AccumulateGrad node per parameter, ~430 per step).  Same values in `p.grad`
    (shared parameters are accumulated); tensor hooks on parameters -- and therefore torch DDP's bucket hooks -- do
    not fire in this mode, so use it with FlatGradAllReduce or on a single GPU."""
    from . import blocks
    blocks.DIRECT_PARAM_GRADS = bool(enable)
