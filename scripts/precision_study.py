"""CPU precision study (test infrastructure): runs the block composition of blocks.py on the emulated kernels with
bf16 STORAGE at exactly the points where the CUDA kernels store bf16, and compares parameter gradients with the fp32
oracle and with PyTorch's bf16 autocast of the oracle.  Used to locate where the bf16 path loses precision
(SAP/OG gradients) before changing the CUDA kernels.   python scripts/precision_study.py sap full"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import emu_kernels
from bevbert_b200 import synth
from bevbert_b200.config import make_config
from bevbert_b200.model.pretrain_cmt import GlocalTextPathCMTPreTraining
from helpers import grad_errors, rel_l2, small_config, small_synth
from oracle import bevbert_ref as R

task = sys.argv[1] if len(sys.argv) > 1 else "sap"
depth = sys.argv[2] if len(sys.argv) > 2 else "full"
ALL = "gemm_in gemm_out gemm_dx ln_y ln_dx ln_dres flash_p flash_ds flash_o flash_dqkv other".split()
emu_kernels.install()
if "NOROUND" in os.environ:      # study mode: every site rounds to bf16 except the ones listed in NOROUND
    act = torch.float64
    emu_kernels.set_act_dtype(act, [s_ for s_ in ALL if s_ not in os.environ["NOROUND"].split(",")])
else:
    act = torch.bfloat16 if os.environ.get("ACT", "bf16") == "bf16" else torch.float32
    emu_kernels.set_act_dtype(act)
torch.set_num_threads(int(os.environ.get("NT", "8")))

if depth == "full":
    cfg = make_config(bev_dim=11, bev_res=1.0, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, feat_dropout=0.0)
else:
    cfg = small_config()
scfg = small_synth()
model = synth.det_init_(GlocalTextPathCMTPreTraining(cfg), seed=3).train()
sd = {k: v.detach().float().clone().requires_grad_(True) for k, v in model.state_dict().items()}
b = synth.make_batch(scfg, seed=7, task=task)
out = model(synth.batch_to(b, "cpu"), task, compute_loss=True)
out.float().mean().backward()


def oracle(autocast):
    for v in sd.values():
        v.grad = None
    if autocast:
        with torch.autocast("cpu", dtype=torch.bfloat16):
            o = R.forward(sd, synth.clone_batch(b), task, R.OracleConfig(cfg))
    else:
        o = R.forward(sd, synth.clone_batch(b), task, R.OracleConfig(cfg))
    o.float().mean().backward()
    return o.detach().float(), {n: v.grad.clone() for n, v in sd.items() if v.grad is not None}


ref, rg = oracle(False)
aco, ag = oracle(True)
names = [n for n, _ in model.named_parameters()]
mine = {n: p.grad for n, p in model.named_parameters()}
errs, glob = grad_errors(mine, {n: rg.get(n) for n in names})
errs_ac, glob_ac = grad_errors({n: ag.get(n) for n in names}, {n: rg.get(n) for n in names})
print("NOROUND=%s " % os.environ.get("NOROUND"), end="")
print("task=%s depth=%s act=%s  loss err ours %.3e autocast %.3e | grads ours %.3e autocast %.3e" % (
    task, depth, act, rel_l2(out, ref), rel_l2(aco, ref), glob, glob_ac))
for n, e in sorted(errs.items(), key=lambda kv: -kv[1])[:6]:
    print("    %-75s ours %.3e autocast %.3e" % (n, e, errs_ac[n]))
