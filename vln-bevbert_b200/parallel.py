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
    """all-reduce (average) the gradients of `params` over the data-parallel group, overlapped with backward.

    From the second step of a task on, the blocks carve every gradient buffer from one zero-filled fp32 arena per step
    (blocks._ZeroArena), in the same order on every rank, so the arena itself is the flat bucket and is reduced in
    place -- no flatten / unflatten copies.  The arena fills in backward order (heads first, language encoder last);
    on CUDA the reduction is issued in `chunks` pieces on a communication stream while backward is still running:
    after each block's backward a hook checks how far the arena has been carved and launches the NCCL all-reduce
    (ReduceOp.AVG, so no separate scaling pass) of the finished prefix; only the last piece (the language-encoder and
    embedding gradients, written last) is exposed after backward.  Parameters that receive a second gradient
    contribution later in backward (the tied word-embedding / MLM-decoder matrix) keep both arena slices apart until
    the reductions are done (the average of the sum is the sum of the averages), see blocks.PENDING_ADDS.
    Gradients that live outside the arena (the first step of a task, the few small heads that run on torch autograd)
    go through a second, small flat buffer.  With graphs.GraphedTrainStep the graph holds forward + backward only and
    this object is called once after the replay (one all-reduce of the whole arena): collectives stay out of the
    capture (graphs.py explains why), so the chunked overlap is what the EAGER loop gets."""

    def __init__(self, params, world_size=None, chunks=3):
        self.params = [p for p in params if p.requires_grad]
        self.world = world_size or (dist.get_world_size() if dist.is_initialized() else 1)
        self.buf = None
        self.chunks = max(1, int(chunks))
        self.comm = None
        self._arena_id = None
        self._done = 0
        self._launched = 0
        self.enabled = True      # set False around rank-local forward/backward passes (no collective may be issued)
        if self.world > 1:
            from . import blocks
            blocks.AFTER_BLOCK_BWD = self._after_block
            blocks.DEFER_GRAD_ADDS = True

    # ------------------------------------------------------------------ overlap with backward
    def _overlap_ok(self, arena):
        return arena is not None and arena.is_cuda and dist.get_backend() == "nccl"

    def _reduce_range(self, arena, lo, hi):
        if hi <= lo:
            return
        if self.comm is None:
            self.comm = torch.cuda.Stream()
        ev = torch.cuda.Event()
        ev.record()
        from . import kernels as K
        side = K.side_stream()
        with torch.cuda.stream(self.comm):
            self.comm.wait_event(ev)
            if side is not None:            # weight gradients of the finished prefix may still be on the side stream
                self.comm.wait_stream(side)
            dist.all_reduce(arena[lo:hi], op=dist.ReduceOp.AVG)

    def _after_block(self):
        """called by blocks._BlockFn.backward after every block (direct-gradient mode)"""
        from . import blocks
        A = blocks.ARENA
        arena = A.buf
        if self.world == 1 or not self.enabled or not self._overlap_ok(arena):
            return
        if A.step_id != self._arena_id:                     # first block of a new step
            self._arena_id, self._done, self._launched = A.step_id, 0, 0
        if self._launched >= self.chunks - 1:
            return
        step = arena.numel() // self.chunks
        if A.off - self._done >= step:
            self._reduce_range(arena, self._done, A.off)
            self._done = A.off
            self._launched += 1

    @torch.no_grad()
    def __call__(self):
        from . import blocks
        if blocks.side_active():
            from . import kernels as K
            K.side_join()
            blocks.SIDE_KEEP.clear()
        if self.world == 1:
            blocks.apply_pending_adds()
            return
        A = blocks.ARENA
        arena = A.buf
        grads = [p.grad for p in self.params if p.grad is not None]
        if not grads:
            return
        inv = 1.0 / self.world
        if arena is not None and A.off > 0:
            if _CHECK_LAYOUT:      # debug (BEVBERT_CHECK_ARENA=1): every rank must have carved the same number of floats
                n = torch.tensor([A.off, -A.off], dtype=torch.int64, device=arena.device)
                dist.all_reduce(n, op=dist.ReduceOp.MAX)
                if int(n[0]) != -int(n[1]):
                    raise RuntimeError("gradient arena layouts differ across ranks (%d..%d floats)" % (-int(n[1]), int(n[0])))
            base = arena.untyped_storage().data_ptr()
            rest = [g for g in grads if g.untyped_storage().data_ptr() != base]
            if self._overlap_ok(arena):
                if A.step_id != self._arena_id:
                    self._arena_id, self._done, self._launched = A.step_id, 0, 0
                self._reduce_range(arena, self._done, A.off)          # the tail, written last by backward
                self._done = A.off
                torch.cuda.current_stream().wait_stream(self.comm)
            else:
                used = arena[:A.off]
                used.mul_(inv)
                dist.all_reduce(used)
        else:
            rest = grads
        pend = [g for _, g in blocks.PENDING_ADDS]
        rest = rest + [g for g in pend if arena is None or g.untyped_storage().data_ptr() != arena.untyped_storage().data_ptr()]
        if rest:
            n = sum(g.numel() for g in rest)
            if self.buf is None or self.buf.numel() < n:
                self.buf = torch.empty(n, dtype=torch.float32, device=rest[0].device)
            flat = self.buf[:n]
            views = list(torch.split(flat, [g.numel() for g in rest]))
            torch._foreach_copy_(views, [g.reshape(-1) for g in rest])
            flat.mul_(inv)
            dist.all_reduce(flat)
            torch._foreach_copy_([g.view(-1) if g.is_contiguous() else g.reshape(-1) for g in rest], views)
        blocks.apply_pending_adds()


def broadcast_parameters(model, src=0):
    """rank `src` -> all, once (the DDP constructor's initial broadcast, utils/misc.py:71-72)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    for t in list(model.parameters()) + list(model.buffers()):
        dist.broadcast(t.data, src)


def enable_side_stream(enable=True):
    """Run the weight-gradient GEMMs and bias column sums of the native backward executors on a second CUDA stream
    (kernels.set_side_stream): nothing in backward depends on them, and the 2560-row language-encoder layers fill only
    0.4-1.6 waves of GEMM tiles, so they overlap the dX chain.  Requires direct parameter gradients and an end-of-backward
    join: `bevbert_b200.optim.AdamW.step`, `FlatGradAllReduce.__call__` and `graphs.GraphedTrainStep` call
    `blocks.join_side()` themselves; a hand-written loop calls it after `loss.backward()`."""
    from . import kernels as K
    if enable:
        direct_param_grads(True)
        if K.side_stream() is None:
            K.set_side_stream(torch.cuda.Stream())
    else:
        from . import blocks
        blocks.join_side()
        K.set_side_stream(None)


def direct_param_grads(enable=True):
    """Let every encoder block write `p.grad` of its parameters itself at the end of its backward instead of returning
    the gradients to autograd (one AccumulateGrad node per parameter, ~430 per step).  Same values in `p.grad`
    (shared parameters are accumulated); tensor hooks on parameters -- and therefore torch DDP's bucket hooks -- do
    not fire in this mode, so use it with FlatGradAllReduce or on a single GPU."""
    from . import blocks
    blocks.DIRECT_PARAM_GRADS = bool(enable)
    if not enable:
        blocks.DEFER_GRAD_ADDS = False
        blocks.AFTER_BLOCK_BWD = None
