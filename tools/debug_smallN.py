"""Forward parity of AtariNet at tiny frame counts (inference sizes) per backend."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from oracle import learner_torch as LT
from torchbeast_b200 import monobeast

A = 6
for use_lstm in (False, True):
    for (T1, B) in ((1, 1), (1, 2), (2, 1), (1, 3), (1, 5), (1, 48), (3, 2)):
        batch = LT.synthetic_batch(T1 - 1, B, A, seed=31)
        params = LT.random_params(LT.atarinet_param_shapes(A, use_lstm), seed=32)
        state = ()
        if use_lstm:
            rs = np.random.RandomState(33)
            state = tuple(torch.from_numpy(rs.randn(2, B, 519).astype(np.float32) * 0.1) for _ in range(2))
        ol, ob, _ = LT.atarinet_forward(params, batch["frame"], batch["reward"], batch["done"], batch["last_action"], state)
        for prec in ("fp32", "bf16", "bf16x3"):
            m = monobeast.AtariNet((4, 84, 84), A, use_lstm, precision=prec)
            m.load_state_dict(params); m.eval()
            with torch.no_grad():
                o, _ = m({k: v.cuda() for k, v in batch.items()}, tuple(s.cuda() for s in state))
            e = float((o["policy_logits"].cpu() - ol).abs().max())
            print("lstm=%d T1=%d B=%d %-7s max|dlogits| %.2e %s" % (use_lstm, T1, B, prec, e, "BAD" if e > (1e-4 if prec != "bf16" else 5e-3) else ""), flush=True)
