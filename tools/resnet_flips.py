"""How much of the split-bf16 ResNet's gradient difference from the fp32 backend is discrete decisions flipping
(max-pool argmax, ReLU sign) rather than rounding: runs the same T=4,B=2 case on both backends, reads the section
buffers out of the two workspaces (layout of res_ws in csrc/resnet.cu) and counts the decisions that differ."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from tests.test_resnet_gpu import build

SEC_S, SEC_SO, SEC_CH = (84, 42, 21), (42, 21, 11), (16, 32, 32)


def sections(ws, N):
    off = 0
    out = []
    def take(nbytes):
        nonlocal off
        r = off
        off += (nbytes + 255) & ~255
        return r
    for i in range(3):
        big, small = N * SEC_S[i] ** 2 * SEC_CH[i], N * SEC_SO[i] ** 2 * SEC_CH[i]
        d = {}
        d["P"] = ws[take(big * 4):][:big * 4].view(torch.float32)
        for k in ("X0", "Y1", "X1", "Y2", "X2"):
            d[k] = ws[take(small * 4):][:small * 4].view(torch.float32)
        d["arg"] = ws[take(small):][:small]
        out.append({k: v.clone() for k, v in d.items()})
    return out


def run(fname, precision):
    g, model, actor, batch, params, state, opt, sched, flags = build(fname, precision)
    cb = {k: v.cuda() for k, v in batch.items()}
    out = model.learner_forward(cb, tuple(s.cuda() for s in state))
    N = cb["frame"].shape[0] * cb["frame"].shape[1]
    secs = sections(model._ws, N)
    rs = np.random.RandomState(0)
    w1 = torch.from_numpy(rs.randn(*out.policy_logits.shape)).float().cuda()
    w2 = torch.from_numpy(rs.randn(*out.baseline.shape)).float().cuda()
    model.learner_backward(w1.contiguous(), w2.contiguous())
    grads = {n: p.grad.detach().cpu().double().clone() for n, p in model.named_parameters()}
    return secs, grads, out


for fname in ("learn_resnet_T4_B2.npz",):
    sa, ga, oa = run(fname, "fp32")
    sb, gb, ob = run(fname, "bf16x3")
    print("outputs rel:", float((oa.policy_logits - ob.policy_logits).norm() / oa.policy_logits.norm()))
    for i in range(3):
        arg = int((sa[i]["arg"] != sb[i]["arg"]).sum())
        tot = sa[i]["arg"].numel()
        flips = {k: int(((sa[i][k] > 0) != (sb[i][k] > 0)).sum()) for k in ("X0", "Y1", "X1", "Y2", "X2")}
        relP = float((sa[i]["P"] - sb[i]["P"]).norm() / sa[i]["P"].norm())
        print("section %d: argmax differs %d / %d, relu sign flips %s, conv output rel diff %.2e" % (i, arg, tot, flips, relP))
    for n in ga:
        print("  %-28s rel diff vs fp32 backend %.2e" % (n, float((ga[n] - gb[n]).norm() / ga[n].norm().clamp_min(1e-30))))
