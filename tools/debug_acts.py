"""Compare the CUDA path's saved activations (workspace) with the fp64 oracle's, incl. ReLU sign flips."""
import sys, os
import numpy as np, torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.test_learner_gpu import build_case, to_cuda

fname = sys.argv[1] if len(sys.argv) > 1 else "learn_atari_T40_B6_clip10.npz"
g, model, actor, batch, params, state, opt, sched = build_case(fname)
cb = to_cuda(batch)
out = model.learner_forward(cb, ())
torch.cuda.synchronize()
T1, B = batch["frame"].shape[:2]
N = T1 * B
ws = model._ws
def al(n): return (n + 255) // 256 * 256
off = 0
M1, M2, M3 = N * 400, N * 81, N * 49
off += al(M1 * 256)
def takef(n):
    global off
    t = ws[off:off + 4 * n].view(torch.float32); off += al(4 * n); return t
act1 = takef(M1 * 32).view(N, 20, 20, 32); col2 = takef(M2 * 512); act2 = takef(M2 * 64).view(N, 9, 9, 64)
col3 = takef(M3 * 576); act3 = takef(N * 3136).view(N, 7, 7, 64)
p = {k: v.double() for k, v in params.items()}
x = batch["frame"].reshape(N, 4, 84, 84).double() / 255.0
z1 = F.conv2d(x, p["conv1.weight"], p["conv1.bias"], stride=4); a1 = F.relu(z1)
z2 = F.conv2d(a1, p["conv2.weight"], p["conv2.bias"], stride=2); a2 = F.relu(z2)
z3 = F.conv2d(a2, p["conv3.weight"], p["conv3.bias"], stride=1); a3 = F.relu(z3)
for name, mine, ref, z in (("act1", act1, a1, z1), ("act2", act2, a2, z2), ("act3", act3, a3, z3)):
    mine = mine.cpu().double().permute(0, 3, 1, 2)
    err = (mine - ref).abs().max().item()
    flips = ((mine > 0) != (ref > 0))
    print(name, "maxerr %.3e" % err, "sign flips", int(flips.sum()), "min |z| at flips", (z.abs()[flips].min().item() if flips.any() else None),
          "max |z| at flips", (z.abs()[flips].max().item() if flips.any() else None))
