"""Per-parameter gradient comparison of the CUDA path against the fp64 oracle for one golden case."""
import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import learner_torch as LT
from tests.test_learner_gpu import build_case, to_cuda
from torchbeast_b200 import learner

fname = sys.argv[1] if len(sys.argv) > 1 else "learn_atari_T40_B6_clip10.npz"
g, model, actor, batch, params, state, opt, sched = build_case(fname)
p64 = {k: v.double() for k, v in params.items()}
o = LT.learner_step(p64, batch, tuple(s.double() for s in state), net="atari", update=False)
cb = to_cuda(batch)
out = model.learner_forward(cb, tuple(s.cuda() for s in state))
loss = learner.impala_loss_fwd_bwd(cb["policy_logits"][1:], out.policy_logits[:-1], cb["action"][1:], cb["reward"][1:],
                                   cb["done"][1:], out.baseline[:-1], out.baseline[-1])
model.learner_backward(loss.grad_logits, loss.grad_values)
print("loss", float(loss.losses[3]), float(o["total_loss"]))
print("logits maxerr", float((out.policy_logits.cpu().double() - o["policy_logits"]).abs().max()))
tot = 0
for n, p in model.named_parameters():
    ref = o["grads"][n]; got = p.grad.cpu().double()
    tot += float((got ** 2).sum())
    print("%-22s norm got %.7f ref %.7f  maxabs err %.3e (ref max %.3e)" % (n, got.norm(), ref.norm(), (got - ref).abs().max(), ref.abs().max()))
print("total norm got %.6f ref %.6f" % (tot ** 0.5, float(o["grad_norm"])))
opt.step(max_grad_norm=float(g["clip"]))
torch.cuda.synchronize()
print("kernel norm", float(opt.grad_norm), "sumsq", float(opt._sumsq))
