"""Compare parameter gradients of the two-layer backward wavefront kernel with the per-layer kernels."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from tests.test_learner_gpu import build_case, to_cuda
fname = sys.argv[1] if len(sys.argv) > 1 else "learn_atari_lstm_T4_B2.npz"
res = {}
for mode in ("0", "1"):
    os.environ["TB_LSTM_WAVE_BWD"] = mode
    g, model, actor, batch, params, state, opt, sched = build_case(fname, precision="bf16")
    cb = to_cuda(batch)
    out = model.learner_forward(cb, tuple(s.cuda() for s in state))
    rs = np.random.RandomState(0)
    w1 = torch.from_numpy(rs.randn(*out.policy_logits.shape)).float().cuda(); w2 = torch.from_numpy(rs.randn(*out.baseline.shape)).float().cuda()
    model.learner_backward(w1.contiguous(), w2.contiguous())
    torch.cuda.synchronize()
    res[mode] = {n: p.grad.detach().clone() for n, p in model.named_parameters()}
for n in res["0"]:
    a, b = res["0"][n].double(), res["1"][n].double()
    print(f"{n:28s} rel {float((a-b).norm()/a.norm().clamp_min(1e-30)):.3e}  norm {float(a.norm()):.3e}")
