"""N > 1 path on CPU (gloo, world_size 2): the model under torch DistributedDataParallel
(find_unused_parameters=True, as pretrain_src/utils/misc.py:70), and under the flat gradient all-reduce of
bevbert_b200/parallel.py, gives the gradients of a single process on the concatenated batch.  Kernels are emulated (tests/emu_kernels.py)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, port, task, ret, mode="ddp"):
    sys.path.insert(0, os.path.dirname(HERE))
    sys.path.insert(0, HERE)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), BEVBERT_CHECK_ARENA="1")
    torch.set_num_threads(2)
    import emu_kernels
    emu_kernels.install()
    from bevbert_b200 import synth
    from bevbert_b200.model.pretrain_cmt import GlocalTextPathCMTPreTraining
    from helpers import small_config, small_synth
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = small_config()
    full = synth.make_batch(small_synth(batch_size=4), seed=9, task=task)
    shard = synth.split_batch(full, world)[rank]
    model = synth.det_init_(GlocalTextPathCMTPreTraining(cfg), seed=3).train()
    if mode == "ddp":
        ddp = torch.nn.parallel.DistributedDataParallel(model, find_unused_parameters=True)
        ddp(shard, task).mean().backward()
    else:       # bevbert_b200.parallel: one flat all-reduce after backward (what bench.py runs at N > 1)
        from bevbert_b200.parallel import FlatGradAllReduce, broadcast_parameters
        broadcast_parameters(model)
        reducer = FlatGradAllReduce(model.parameters(), world)
        for _ in range(2):      # the second step runs on the step arena (reduced in place), the first on the fallback
            model.zero_grad(set_to_none=True)
            model(synth.clone_batch(shard), task).mean().backward()
            reducer()
        from bevbert_b200 import blocks
        assert blocks.ARENA.buf is not None and blocks.ARENA.off > 0
    grads = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
    if rank == 0:
        single = synth.det_init_(GlocalTextPathCMTPreTraining(cfg), seed=3).train()
        single(full, task).mean().backward()
        ref = {n: p.grad for n, p in single.named_parameters() if p.grad is not None}
        top = max(float(g.norm()) for g in ref.values())
        worst = 0.0
        assert set(ref) == set(grads), set(ref) ^ set(grads)
        for n, g in ref.items():
            worst = max(worst, float((grads[n] - g).norm()) / max(float(g.norm()), 1e-4 * top))
        ret["worst"] = worst
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("task,mode", [("sap", "ddp"), ("mlm", "ddp"), ("sap", "flat"), ("mlm", "flat")])
# (masksem averages over a per-shard count of masked cells, so mean-of-shard-means != global mean in the reference too)
def test_ddp_gradients_equal_single_process(task, mode):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, task, ret, mode), nprocs=2, join=True)
    assert ret["worst"] < 1e-3, ret["worst"]
