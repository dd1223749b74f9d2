"""GPU: the inference path (reference polybeast_learner.py:269-285 + tests/polybeast_inference_test.py):
T = 1, B in {1, 48, 512} actors through polybeast_learner.inference with a mock DynamicBatcher batch; outputs on the CPU
with the reference's shapes, logits / baseline / carried LSTM state equal to the torch oracle's forward."""
import types
import unittest.mock as mock

import numpy as np
import pytest
import torch

from oracle import learner_torch as LT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("use_lstm", [False, True])
@pytest.mark.parametrize("B", [1, 48, 512])
def test_atarinet_inference_matches_oracle(B, use_lstm):
    from torchbeast_b200 import monobeast, polybeast_learner
    A = 6
    batch = LT.synthetic_batch(0, B, A, seed=31)  # T + 1 = 1 row
    params = LT.random_params(LT.atarinet_param_shapes(A, use_lstm), seed=32)
    model = monobeast.AtariNet((4, 84, 84), A, use_lstm)
    model.load_state_dict(params)
    model.eval()
    state = ()
    if use_lstm:
        rs = np.random.RandomState(33)
        state = tuple(torch.from_numpy(rs.randn(2, B, 512 + A + 1).astype(np.float32) * 0.1) for _ in range(2))
    env = (batch["frame"], batch["reward"], batch["done"], batch["episode_step"], batch["episode_return"], batch["last_action"])
    mb = mock.MagicMock()
    mb.get_inputs = mock.Mock(return_value=(env, state))
    mb.set_outputs = mock.Mock()
    batcher = mock.MagicMock()
    batcher.__iter__.return_value = iter([mb])
    flags = types.SimpleNamespace(actor_device="cuda:0", use_lstm=use_lstm)
    polybeast_learner.inference(flags, batcher, model)
    mb.get_inputs.assert_called_once()
    mb.set_outputs.assert_called_once()
    (outputs,), kw = mb.set_outputs.call_args
    assert kw == {}
    (action, logits, baseline), core_state = outputs
    assert tuple(action.shape) == (1, B) and tuple(logits.shape) == (1, B, A) and tuple(baseline.shape) == (1, B)
    for t in (action, logits, baseline) + tuple(core_state):
        assert t.device == torch.device("cpu")
    assert len(core_state) == (2 if use_lstm else 0)
    ol, ob, ostate = LT.atarinet_forward(params, batch["frame"], batch["reward"], batch["done"], batch["last_action"], state)
    np.testing.assert_allclose(logits.numpy(), ol.numpy(), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(baseline.numpy(), ob.numpy(), rtol=1e-4, atol=1e-4)
    assert torch.equal(action, logits.argmax(-1))  # eval mode: greedy (monobeast.py:621-623)
    for a, b in zip(core_state, ostate):
        assert tuple(a.shape) == (2, B, 512 + A + 1)
        np.testing.assert_allclose(a.numpy(), b.numpy(), rtol=1e-4, atol=1e-4)


def test_state_carry_over_two_inference_calls():
    """Two T=1 calls carrying the state == one T=2 forward (what an actor sees across DynamicBatcher calls)."""
    from torchbeast_b200 import monobeast
    A, B = 6, 48
    batch = LT.synthetic_batch(1, B, A, seed=41)
    params = LT.random_params(LT.atarinet_param_shapes(A, True), seed=42)
    model = monobeast.AtariNet((4, 84, 84), A, True)
    model.load_state_dict(params)
    model.eval()
    cb = {k: v.cuda() for k, v in batch.items()}
    with torch.no_grad():
        full, _ = model(cb, model.initial_state(B))
        st = model.initial_state(B)
        rows = []
        for t in range(2):
            o, st = model({k: v[t:t + 1] for k, v in cb.items()}, st)
            rows.append(o["policy_logits"])
    np.testing.assert_allclose(torch.cat(rows).cpu().numpy(), full["policy_logits"].cpu().numpy(), rtol=1e-5, atol=1e-5)
