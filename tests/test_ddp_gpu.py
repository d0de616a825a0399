"""2-rank NCCL gradient parity on real GPUs: the CUDA path under the overlapped flat
all-reduce of bevbert_b200/parallel.py (direct parameter gradients, chunked ReduceOp.AVG on a communication stream while
backward runs, deferred second contributions to the tied embedding matrix) reproduces the gradients of ONE process on
the concatenated batch; then the same through whole-step CUDA-graph replay (NCCL captured in the graph).
Skipped on boxes with fewer than two GPUs."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
pytestmark = pytest.mark.gpu


def _worker(rank, world, port, task, ret):
    sys.path.insert(0, os.path.dirname(HERE))
    sys.path.insert(0, HERE)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    from bevbert_b200 import blocks, synth
    from bevbert_b200.graphs import GraphedTrainStep
    from bevbert_b200.model.ops import prepare_batch
    from bevbert_b200.model.pretrain_cmt import GlocalTextPathCMTPreTraining
    from bevbert_b200.optim import AdamW, build_param_groups
    from bevbert_b200.parallel import FlatGradAllReduce, broadcast_parameters, direct_param_grads
    from helpers import small_config, small_synth
    import datetime
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank),
                            timeout=datetime.timedelta(seconds=90))
    cfg = small_config()
    full = synth.make_batch(small_synth(batch_size=4), seed=9, task=task)
    shard = synth.batch_to(prepare_batch(synth.split_batch(full, world)[rank]), "cuda")
    model = synth.det_init_(GlocalTextPathCMTPreTraining(cfg), seed=3).cuda().train()
    broadcast_parameters(model)
    direct_param_grads(True)
    reducer = FlatGradAllReduce(model.parameters(), world, chunks=3)
    for _ in range(3):          # from the second step on the arena is reduced in place, in overlapped chunks
        model.zero_grad(set_to_none=True)
        model(shard, task).mean().backward()
        reducer()
    torch.cuda.synchronize()
    assert blocks.ARENA.buf is not None and reducer._launched >= 1, "the overlapped path did not run"
    grads = {n: p.grad.detach().float().cpu() for n, p in model.named_parameters() if p.grad is not None}
    # graph replay with the collective inside: three optimizer steps must track the eager loop of rank-identical models
    opt = AdamW(build_param_groups(model, 0.01), lr=0.0, max_grad_norm=5.0, runtime=model.rt)
    step = GraphedTrainStep(model, opt, reducer, warmup=1)
    losses = [float(step(shard, task)) for _ in range(4)]
    graphed = step.launches(shard, task)
    if rank == 0:
        direct_param_grads(False)
        single = synth.det_init_(GlocalTextPathCMTPreTraining(cfg), seed=3).cuda().train()
        single(synth.batch_to(full, "cuda"), task).mean().backward()
        ref = {n: p.grad.detach().float().cpu() for n, p in single.named_parameters() if p.grad is not None}
        top = max(float(g.norm()) for g in ref.values())
        assert set(ref) == set(grads), set(ref) ^ set(grads)
        worst = max(float((grads[n] - g).norm()) / max(float(g.norm()), 1e-2 * top) for n, g in ref.items())
        num = sum(float((grads[n] - g).norm()) ** 2 for n, g in ref.items())
        den = sum(float(g.norm()) ** 2 for g in ref.values())
        ret.update(worst=worst, glob=(num / den) ** 0.5, losses=losses, graphed=graphed)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("task", ["sap", "mlm"])
def test_two_rank_nccl_gradients_equal_single_process(task):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(2, port, task, ret), nprocs=2, join=True)
    print("2-rank %s: global grad rel err %.3e worst %.3e losses %s graph launches %s" % (
        task, ret["glob"], ret["worst"], ret["losses"], ret["graphed"]))
    # two bf16 runs with different batch splits: the bf16 bars of the single-GPU parity tests apply
    assert ret["glob"] < (1.3e-1 if task == "sap" else 3e-2), dict(ret)
    assert ret["graphed"], "the step with the NCCL all-reduce was not captured"
    assert max(ret["losses"]) - min(ret["losses"]) <= 1e-3 * abs(ret["losses"][0])      # lr = 0: replays reproduce the loss
