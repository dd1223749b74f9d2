"""AtariNet: results must not depend on what a previous owner left in the workspace (torch.empty memory)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
@pytest.mark.parametrize("use_lstm", [False, True])
def test_atarinet_ignores_stale_workspace(precision, use_lstm):
    from oracle import learner_torch as LT
    from torchbeast_b200 import monobeast
    T, B, A = 5, 3, 6
    batch = LT.synthetic_batch(T, B, A, seed=3, with_last_action=True)
    model = monobeast.AtariNet((4, 84, 84), A, use_lstm, precision=precision)
    cb = {k: v.cuda() for k, v in batch.items()}
    st = model.initial_state(B)
    rs = np.random.RandomState(1)
    runs = []
    for fill in (None, 255, 0):
        if fill is not None:
            model._ws.fill_(fill)
        out = model.learner_forward(cb, st)
        w1 = torch.from_numpy(rs.randn(*out.policy_logits.shape)).float().cuda() if not runs else runs[0][2]
        w2 = torch.from_numpy(rs.randn(*out.baseline.shape)).float().cuda() if not runs else runs[0][3]
        model.learner_backward(w1.contiguous(), w2.contiguous())
        runs.append((out.policy_logits.clone(), model.flat_grad.clone(), w1, w2))
    for logits, grads, _, _ in runs[1:]:
        assert torch.isfinite(grads).all()
        assert torch.equal(logits, runs[0][0]) and torch.equal(grads, runs[0][1])
