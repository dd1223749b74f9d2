"""CPU: host-side logic of the network modules (no kernels run): state_dict contract of the
reference, flat-buffer aliasing, parameter layout == the C-ABI's."""
import numpy as np
import torch

from oracle import learner_torch as LT
from tests.common import golden


def test_state_dict_keys_and_shapes_match_reference():
    from torchbeast_b200.nets import AtariNet
    for use_lstm, fname in ((False, "learn_atari_T4_B2.npz"), (True, "learn_atari_lstm_T4_B2.npz")):
        m = AtariNet((4, 84, 84), 6, use_lstm, device="cpu")
        sd = m.state_dict()
        shapes = LT.atarinet_param_shapes(6, use_lstm)
        assert list(sd.keys()) == list(shapes.keys())
        for k, v in sd.items():
            assert tuple(v.shape) == tuple(shapes[k]), k
        # the golden fixture was recorded from the reference module's named_parameters()
        g = golden(fname)
        ref_names = [k.split("/", 1)[1] for k in g.files if k.startswith("grad_stats/")]
        assert sorted(ref_names) == sorted(sd.keys())
        assert sum(p.numel() for p in m.parameters()) == (6005848 if use_lstm else 1687768)


def test_flat_aliasing_survives_load_and_to():
    from torchbeast_b200.nets import AtariNet
    m = AtariNet((4, 84, 84), 6, False, device="cpu")
    p = LT.random_params(LT.atarinet_param_shapes(6, False), seed=3)
    m.load_state_dict(p)
    off = 0
    for name, t in m.state_dict().items():
        n = t.numel()
        assert torch.equal(m.flat_params[off:off + n].view(t.shape), p[name]), name
        assert t.data_ptr() == m.flat_params.data_ptr() + 4 * off
        off += n
    m2 = AtariNet((4, 84, 84), 6, False, device="cpu")
    m2.load_state_dict(m.state_dict())  # one-copy path
    assert torch.equal(m2.flat_params, m.flat_params)
    m.to("cpu")
    m.conv1.weight.data.zero_()
    assert float(m.flat_params[:8192].abs().sum()) == 0.0
    fg = m.attach_grads()
    assert m.fc.bias.grad.data_ptr() == fg.data_ptr() + 4 * (8192 + 32 + 32768 + 64 + 36864 + 64 + 512 * 3136)


def test_initial_state_and_init_distribution():
    from torchbeast_b200.nets import AtariNet
    m = AtariNet((4, 84, 84), 6, True, device="cpu")
    h, c = m.initial_state(5)
    assert tuple(h.shape) == (2, 5, 519) and float(h.abs().sum()) == 0
    assert AtariNet((4, 84, 84), 6, False, device="cpu").initial_state(3) == tuple()
    bound = 1 / np.sqrt(3136)
    w = m.fc.weight.detach()
    # |w| <= bound up to fp32 rounding of the bound itself: uniform() can return exactly -1 (probability 2^-24 per
    # element, ~10 % over the 1.6 M weights), and float32(bound) is slightly larger than the float64 value
    assert float(w.abs().max()) <= bound * (1 + 1e-6) and float(w.std()) > 0.5 * bound / np.sqrt(3)
