"""Parity of one polybeast_learner.learn() step of the IMPALA ResNet (+LSTM) at T=80, B=8 (one GPU's shard of BASELINE
configs[3]) per backend, against the fixture the reference itself produced (tests/golden/learn_resnet_lstm_T80_B8.npz):
worst errors of the learner outputs, V-trace targets, losses, and per gradient tensor (4096 strided samples)."""
import sys
from unittest import mock

import numpy as np
import torch

sys.path.insert(0, ".")
from tests.common import sample_index  # noqa: E402
from tests.test_resnet_gpu import build, queue_of  # noqa: E402
from torchbeast_b200 import learner, polybeast_learner  # noqa: E402

for prec in (sys.argv[1:] or ["fp32", "bf16x3", "bf16"]):
    g, model, actor, batch, params, state, opt, sched, flags = build("learn_resnet_lstm_T80_B8.npz", prec)
    cb = {k: v.cuda() for k, v in batch.items()}
    out = model.learner_forward(cb, tuple(s.cuda() for s in state))
    loss = learner.impala_loss_fwd_bwd(cb["policy_logits"][1:], out.policy_logits[:-1], cb["action"][1:], cb["reward"][1:],
                                       cb["done"][1:], out.baseline[:-1], out.baseline[-1])
    errs = {k: "%.2e" % np.abs(v.cpu().numpy() - g[k]).max() for k, v in (("policy_logits", out.policy_logits), ("baseline", out.baseline),
                                                                         ("vs", loss.vs), ("pg_advantages", loss.pg_advantages))}
    stats = {}
    polybeast_learner.learn(flags, queue_of(batch, state), model, actor, opt, sched, stats, mock.Mock())
    lerr = {k: "%.2e" % (abs(stats[k] - float(g[k])) / max(abs(float(g[k])), 1e-12)) for k in ("total_loss", "pg_loss", "baseline_loss", "entropy_loss")}
    print("== learn_resnet_lstm_T80_B8.npz  precision=%s" % prec)
    print("   max|err| outputs:", errs)
    print("   rel err losses  :", lerr)
    rels = []
    for n, p in model.named_parameters():
        gr = p.grad.detach().cpu().flatten()
        idx = torch.from_numpy(sample_index(gr.numel()))
        ref = torch.from_numpy(g["grad_sample/" + n]).double()
        got = gr[idx].double()
        rel = float((got - ref).norm() / ref.norm().clamp_min(1e-30))
        mx = float((got - ref).abs().max() / ref.abs().max().clamp_min(1e-30))
        perr = float((p.detach().cpu().flatten()[idx].double() - torch.from_numpy(g["param_sample/" + n]).double()).abs().max())
        rels.append(rel)
        print("   grad %-26s relL2 %.2e  max/|max| %.2e   param max|err| %.2e" % (n, rel, mx, perr))
    print("   gradient relL2: median %.2e  worst %.2e" % (float(np.median(rels)), max(rels)))
