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


def test_state_dict_is_plain_tensors_and_round_trips(tmp_path):
    """state_dict() must carry nothing but tensors - torch.save / torch.load(weights_only=True) /
    deepcopy like any nn.Module - and an EDITED dict must be honoured, not silently replaced by the sibling's buffer."""
    import copy
    import io
    import pickle
    from torchbeast_b200.nets import AtariNet
    m = AtariNet((4, 84, 84), 6, False, device="cpu")
    sd = m.state_dict()
    assert not hasattr(sd, "_tb_flat_source") and not [k for k in vars(sd) if not k.startswith("_metadata")]
    buf = io.BytesIO()
    torch.save(sd, buf)  # (plain pickle would store the shared flat storage once per view; torch.save stores it once)
    assert buf.tell() < 4 * m.flat_params.numel() * 1.2 + 65536  # the weights, not the module / workspace behind them
    assert b"AtariNet" not in pickle.dumps({k: tuple(v.shape) for k, v in sd.items()}) and b"AtariNet" not in buf.getvalue()
    f = tmp_path / "model.tar"
    torch.save({"model_state_dict": sd, "flags": {"x": 1}}, f)  # reference checkpoint shape (polybeast_learner.py:539-547)
    ck = torch.load(f, weights_only=True)
    m2 = AtariNet((4, 84, 84), 6, False, device="cpu")
    res = m2.load_state_dict(ck["model_state_dict"])
    assert not res.missing_keys and not res.unexpected_keys and torch.equal(m2.flat_params, m.flat_params)
    sd2 = copy.deepcopy(sd)
    assert isinstance(sd2, dict) and torch.equal(sd2["fc.bias"], sd["fc.bias"])
    # edited copy of a sibling's state_dict: the edit wins (the aliasing fast path must not trigger)
    edited = dict(m.state_dict())
    edited["fc.bias"] = torch.full_like(edited["fc.bias"], 7.0)
    m3 = AtariNet((4, 84, 84), 6, False, device="cpu")
    m3.load_state_dict(edited)
    assert float(m3.fc.bias.min()) == 7.0 and torch.equal(m3.conv1.weight, m.conv1.weight)
    # filtered dict + strict=False: only the given keys change
    before = m3.fc.weight.clone()
    m3.load_state_dict({"conv1.bias": torch.ones(32)}, strict=False)
    assert float(m3.conv1.bias.sum()) == 32.0 and torch.equal(m3.fc.weight, before)
    # unedited sibling dict: one-copy path, same result
    m4 = AtariNet((4, 84, 84), 6, False, device="cpu")
    m4.load_state_dict(m.state_dict())
    assert torch.equal(m4.flat_params, m.flat_params)


def test_attach_grads_keeps_autograd_gradients():
    """gradients autograd left in .grad are copied into the flat buffer, not dropped."""
    from torchbeast_b200.nets import AtariNet
    m = AtariNet((4, 84, 84), 6, False, device="cpu")
    m.fc.bias.grad = torch.full_like(m.fc.bias, 3.0)
    m.policy.bias.grad = torch.full_like(m.policy.bias, 5.0)
    fg = m.attach_grads()
    assert float(m.fc.bias.grad.min()) == 3.0 and float(m.policy.bias.grad.max()) == 5.0
    off = dict((n, o) for (n, _), (_, o, _, _) in zip(m._spec, m._views))
    assert float(fg[off["fc.bias"]]) == 3.0 and float(fg[off["policy.bias"]]) == 5.0
    assert m.fc.bias.grad.data_ptr() == fg.data_ptr() + 4 * off["fc.bias"]


def test_rmsprop_load_state_dict_feeds_the_flat_buffers(tmp_path):
    """ADVICE r1 / SURVEY N4: optimizer.load_state_dict(checkpoint["optimizer_state_dict"]) (reference
    polybeast_learner.py:535-548) must land in the flat square_avg the fused kernel reads - from this class's own
    state_dict AND from a torch.optim.RMSprop state_dict of the reference model."""
    from torchbeast_b200 import optim
    from torchbeast_b200.nets import AtariNet
    m = AtariNet((4, 84, 84), 6, True, device="cpu")
    o = optim.RMSprop(m, lr=0.00048, momentum=0, eps=0.01, alpha=0.99)
    o.square_avg.copy_(torch.rand_like(o.square_avg))
    o._steps = 5
    for st in o.state.values():
        st["step"] = torch.tensor(5.0)
    f = tmp_path / "ck.tar"
    torch.save({"optimizer_state_dict": o.state_dict()}, f)
    m2 = AtariNet((4, 84, 84), 6, True, device="cpu")
    o2 = optim.RMSprop(m2, lr=0.1, momentum=0, eps=0.01, alpha=0.99)
    o2.load_state_dict(torch.load(f, weights_only=True)["optimizer_state_dict"])
    assert torch.equal(o2.square_avg, o.square_avg) and o2._steps == 5 and o2.param_groups[0]["lr"] == 0.00048
    for p, off, n, shape in m2._views:  # the per-parameter entries are views of the flat buffer again
        assert o2.state[p]["square_avg"].data_ptr() == o2.square_avg.data_ptr() + 4 * off
    # a stock torch.optim.RMSprop over reference-shaped parameters (what a reference checkpoint holds)
    ref_params = [torch.nn.Parameter(torch.zeros(s)) for _, s in m._spec]
    ro = torch.optim.RMSprop(ref_params, lr=0.00048, momentum=0, eps=0.01, alpha=0.99)
    for i, p in enumerate(ref_params):
        p.grad = torch.full_like(p, float(i + 1))
    ro.step()
    o3 = optim.RMSprop(AtariNet((4, 84, 84), 6, True, device="cpu"), lr=0.00048, momentum=0, eps=0.01, alpha=0.99)
    o3.load_state_dict(ro.state_dict())
    for i, (p, off, n, shape) in enumerate(o3.model._views):
        assert torch.allclose(o3.square_avg[off:off + n], ro.state[ref_params[i]]["square_avg"].reshape(-1)), i
    assert o3._steps == 1
