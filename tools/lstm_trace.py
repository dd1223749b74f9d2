"""Per-phase timeline of the split LSTM wavefront kernels (TB_LSTM_TRACE): runs one learn step with tracing on and prints,
per kernel, the median cycles between consecutive phase stamps of thread 0 (steady-state steps, all CTAs)."""
import os, sys, struct
import numpy as np
os.environ["TB_LSTM_TRACE"] = "/tmp/lstm_trace"
sys.path.insert(0, ".")
import torch
from bench import synthetic_host_batch, flags_ns
from torchbeast_b200 import learner, monobeast, optim
T, B, A = 80, 32, 6
model = monobeast.AtariNet((4, 84, 84), A, True); actor = monobeast.AtariNet((4, 84, 84), A, True)
opt = optim.RMSprop(model, lr=0.00048, eps=0.01, alpha=0.99)
batch = {k: v.cuda() for k, v in synthetic_host_batch(T, B, A, 1, False).items()}
state = model.initial_state(B)
for _ in range(2):
    learner.learn_step(flags_ns(T, B), model, actor, batch, state, opt, None, stats_sync=False)
torch.cuda.synchronize()
NAMES = {"fwd": ["top", "polled", "h0 tiles landed", "X0 MMAs done (h1 landed)", "all MMAs + partials stored", "after sync C", "after gates+update+publish syncs", "after fence+flag"],
         "bwd": ["top", "pointwise+publish done (2 syncs)", "fence+flag", "polled", "gate0 MMAs", "all MMAs", "reduce done (2 syncs)", "-"]}
for tag in ("fwd", "bwd"):
    raw = open("/tmp/lstm_trace." + tag, "rb").read()
    ctas, steps, ph, _ = struct.unpack("4i", raw[:16])
    a = np.frombuffer(raw[16:], dtype=np.int64).reshape(ctas, steps, ph).astype(np.float64)
    print("==", tag, "ctas", ctas, "steps", steps)
    for lo, hi, label in ((0, ctas, "all") if tag == "fwd" else (0, ctas // 2, "upper role"), ) + (() if tag == "fwd" else ((ctas // 2, ctas, "lower role"),)):
        sl = a[lo:hi, 10:70]
        step_len = np.median(sl[:, 1:, 0] - sl[:, :-1, 0])
        print("  [%s] median cycles per wave step: %.0f (%.2f us at 1.9 GHz)" % (label, step_len, step_len / 1900))
        for p in range(1, ph):
            d = sl[:, :, p] - sl[:, :, p - 1]
            ok = (sl[:, :, p] > 0) & (sl[:, :, p - 1] > 0)
            if ok.any():
                print("     %-38s median %6.0f   p90 %6.0f   max %6.0f" % (NAMES[tag][p], np.median(d[ok]), np.percentile(d[ok], 90), d[ok].max()))
        d = sl[:, 1:, 0] - sl[:, :-1, ph - 1 if tag == "fwd" else 6]
        ok = (sl[:, :-1, ph - 1 if tag == "fwd" else 6] > 0)
        print("     %-38s median %6.0f" % ("tail -> next top", np.median(d[ok])))
