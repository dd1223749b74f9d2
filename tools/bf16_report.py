"""Print per-parameter relative error / cosine of the bf16 backend's backward for fixed cotangents."""
import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import learner_torch as LT
from tests.test_learner_gpu import build_case, to_cuda
for fname in sys.argv[1:] or ["learn_atari_T20_B4.npz", "learn_atari_lstm_T20_B4.npz"]:
    for prec in ("fp32", "bf16"):
        g, model, actor, batch, params, state, opt, sched = build_case(fname, precision=prec)
        A = int(g["meta"][2])
        p64 = {k: v.double().requires_grad_(True) for k, v in params.items()}
        ol, ob, _ = LT.atarinet_forward(p64, batch["frame"], batch["reward"], batch["done"], batch["last_action"], tuple(s.double() for s in state), A)
        rs = np.random.RandomState(0)
        w1 = torch.from_numpy(rs.randn(*ol.shape)); w2 = torch.from_numpy(rs.randn(*ob.shape))
        names = list(p64)
        ref = dict(zip(names, torch.autograd.grad((ol * w1).sum() + (ob * w2).sum(), [p64[n] for n in names])))
        cb = to_cuda(batch)
        out = model.learner_forward(cb, tuple(s.cuda() for s in state))
        rel = lambda a, b: float((a - b).norm() / b.norm())
        print(fname, prec, "logits rel %.2e baseline rel %.2e" % (rel(out.policy_logits.cpu().double(), ol.detach()), rel(out.baseline.cpu().double(), ob.detach())))
        model.learner_backward(w1.float().cuda().contiguous(), w2.float().cuda().contiguous())
        for n, p in model.named_parameters():
            got = p.grad.cpu().double()
            cos = float((got * ref[n]).sum() / (got.norm() * ref[n].norm()))
            print("   %-22s rel %.3e cos %.6f" % (n, rel(got, ref[n]), cos))
