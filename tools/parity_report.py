"""Parity of one learn() step at a golden fixture's size, per backend: prints the worst errors of the learner outputs,
vs / pg_advantages, the four losses, every gradient tensor (strided samples) and the updated parameters against the
reference-generated fixture.  Usage: python tools/parity_report.py [fixture.npz ...] [--precisions fp32,bf16x3,bf16]"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from tests.common import sample_index  # noqa: E402
from tests.test_learner_gpu import build_case, flags_for, to_cuda  # noqa: E402


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    precs = ["fp32", "bf16x3", "bf16"]
    for a in sys.argv[1:]:
        if a.startswith("--precisions="):
            precs = a.split("=", 1)[1].split(",")
    files = args or ["learn_atari_lstm_T80_B32.npz", "learn_atari_T80_B32.npz"]
    from torchbeast_b200 import learner, monobeast
    for fname in files:
        for prec in precs:
            g, model, actor, batch, params, state, opt, sched = build_case(fname, precision=prec)
            flags = flags_for(g)
            cb = to_cuda(batch)
            st = tuple(s.cuda() for s in state)
            out = model.learner_forward(cb, st)
            loss = learner.impala_loss_fwd_bwd(cb["policy_logits"][1:], out.policy_logits[:-1], cb["action"][1:], cb["reward"][1:],
                                               cb["done"][1:], out.baseline[:-1], out.baseline[-1])
            line = {}
            for k, v in (("policy_logits", out.policy_logits), ("baseline", out.baseline), ("vs", loss.vs),
                         ("pg_advantages", loss.pg_advantages)):
                if k in g.files:
                    d = np.abs(v.cpu().numpy() - g[k])
                    line[k] = "%.2e" % d.max()
            torch.cuda.synchronize(); t0 = time.time()
            stats = monobeast.learn(flags, actor, model, cb, st, opt, sched)
            torch.cuda.synchronize(); dt = time.time() - t0
            for k in ("total_loss", "pg_loss", "baseline_loss", "entropy_loss"):
                line[k] = "%.2e" % (abs(stats[k] - float(g[k])) / max(abs(float(g[k])), 1e-12))
            print("== %s  precision=%s  (learn %.1f ms, eager, first call)" % (fname, prec, dt * 1e3))
            print("   max|err| outputs:", {k: line[k] for k in line if not k.endswith("loss")})
            print("   rel err losses  :", {k: line[k] for k in line if k.endswith("loss")})
            worst = []
            for n, p in model.named_parameters():
                gr = p.grad.detach().cpu().flatten()
                if ("grad_sample/" + n) in g.files:
                    idx = torch.from_numpy(sample_index(gr.numel()))
                    ref = torch.from_numpy(g["grad_sample/" + n]).double()
                    got = gr[idx].double()
                    pref = torch.from_numpy(g["param_sample/" + n]).double()
                    pgot = p.detach().cpu().flatten()[idx].double()
                else:
                    ref = torch.from_numpy(g["grad_head/" + n]).double(); got = gr[:16].double()
                    pref = torch.from_numpy(g["param_head/" + n]).double(); pgot = p.detach().cpu().flatten()[:16].double()
                rel = float((got - ref).norm() / ref.norm().clamp_min(1e-30))
                mx = float((got - ref).abs().max() / ref.abs().max().clamp_min(1e-30))
                nrm = abs(float(gr.double().norm()) - float(g["grad_stats/" + n][2])) / max(float(g["grad_stats/" + n][2]), 1e-30)
                pmx = float((pgot - pref).abs().max())
                worst.append((n, rel, mx, nrm, pmx))
            for n, rel, mx, nrm, pmx in worst:
                print("   grad %-24s relL2 %.2e  max/|max| %.2e  norm rel %.2e   param max|err| %.2e" % (n, rel, mx, nrm, pmx))


if __name__ == "__main__":
    main()
