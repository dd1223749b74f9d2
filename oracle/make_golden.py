"""Generate tests/golden/*.npz by RUNNING THE REFERENCE ITSELF (build container only).

TEST INFRASTRUCTURE.  Needs /root/reference (read-only) plus
  oracle/_ref/nest*.so        built from /root/reference/nest/nest/nest_pybind.cc (oracle/Makefile)
  oracle/refshim/{gym,libtorchbeast.py}   empty stand-ins so the modules import
Nothing from /root/reference is copied; only its *outputs* on seeded inputs are stored.

    python oracle/make_golden.py          # rewrites tests/golden/

Fixtures (inputs are numpy-RandomState / arange formulas, so they are regenerated in
the tests rather than stored, except where noted):
  vtrace_fixture.npz   from_importance_weights at the reference test's formulas
                       (tests/vtrace_test.py:136-168) for (T,B) in (5,5),(5,1),(80,4):
                       reference output AND the reference test's O(T^2) ground truth.
  vtrace_random.npz    from_logits / from_importance_weights on seeded random inputs,
                       (T,B,A) in (80,32,6),(20,8,3),(7,2,18); clips (1,1),(None,None),(3.7,2.2)
  losses.npz           the literal-constant loss tests (polybeast_loss_functions_test.py)
                       values and autograd gradients.
  learn_*.npz          one reference learn() step (monobeast AtariNet / polybeast ResNet,
                       with and without LSTM): losses, learner outputs, clipped grads,
                       updated parameters (statistics + leading elements).
"""
import os
import sys
import types
import unittest.mock as mock

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path[:0] = ["/root/reference", os.path.join(HERE, "_ref"), os.path.join(HERE, "refshim"), ROOT]

import numpy as np  # noqa: E402
import torch  # noqa: E402

from torchbeast import monobeast, polybeast_learner  # noqa: E402  (the reference)
from torchbeast.core import vtrace as ref_vtrace  # noqa: E402

from oracle import learner_torch as LT  # noqa: E402  (only for the shared input generators)

OUT = os.path.join(ROOT, "tests", "golden")


def _load_ref_test_module():
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_vtrace_test", "/root/reference/tests/vtrace_test.py")
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def arange_inputs(T, B):
    """The reference test's input formulas (tests/vtrace_test.py:143-157)."""
    ar = np.arange(T * B, dtype=np.float32).reshape(T, B)
    return dict(
        log_rhos=(5 * (ar / (B * T) - 0.5)).astype(np.float32),
        discounts=np.array([[0.9 / (b + 1) for b in range(B)] for _ in range(T)], dtype=np.float32),
        rewards=ar.copy(),
        values=(ar / B).astype(np.float32),
        bootstrap_value=(np.arange(B, dtype=np.float32) + 1.0),
    )


def random_vtrace_inputs(T, B, A, seed):
    """SURVEY.md section 8(d) M2 V-trace-only inputs, numpy RandomState."""
    rs = np.random.RandomState(seed)
    return dict(
        behavior_policy_logits=rs.randn(T, B, A).astype(np.float32),
        target_policy_logits=rs.randn(T, B, A).astype(np.float32),
        actions=rs.randint(0, A, size=(T, B)).astype(np.int64),
        discounts=(0.99 * (rs.rand(T, B) > 0.05)).astype(np.float32),
        rewards=np.clip(rs.randn(T, B), -1, 1).astype(np.float32),
        values=rs.randn(T, B).astype(np.float32),
        bootstrap_value=rs.randn(B).astype(np.float32),
    )


def make_vtrace_fixture():
    rt = _load_ref_test_module()
    out = {}
    for T, B in ((5, 5), (5, 1), (80, 4)):
        v = arange_inputs(T, B)
        ref = ref_vtrace.from_importance_weights(
            **{k: torch.from_numpy(x) for k, x in v.items()}, clip_rho_threshold=3.7, clip_pg_rho_threshold=2.2)
        gt = rt._ground_truth_calculation(
            clip_rho_threshold=3.7, clip_pg_rho_threshold=2.2,
            **{k: x.astype(np.float64) for k, x in v.items()})
        tag = "T%d_B%d" % (T, B)
        out[tag + "_vs"] = ref.vs.numpy()
        out[tag + "_pg"] = ref.pg_advantages.numpy()
        out[tag + "_gt_vs"] = gt.vs
        out[tag + "_gt_pg"] = gt.pg_advantages
    np.savez_compressed(os.path.join(OUT, "vtrace_fixture.npz"), **out)


def make_vtrace_random():
    out = {}
    for (T, B, A), seed in (((80, 32, 6), 1), ((20, 8, 3), 2), ((7, 2, 18), 3), ((600, 16, 6), 4)):
        v = random_vtrace_inputs(T, B, A, seed)
        tv = {k: torch.from_numpy(x) for k, x in v.items()}
        for ci, (c1, c2) in enumerate(((1.0, 1.0), (None, None), (3.7, 2.2))):
            r = ref_vtrace.from_logits(clip_rho_threshold=c1, clip_pg_rho_threshold=c2, **tv)
            tag = "T%d_B%d_A%d_c%d" % (T, B, A, ci)
            for name in (r._fields if T * B <= 200 else ("vs", "pg_advantages")):
                out[tag + "_" + name] = getattr(r, name).numpy()
            if T * B > 200 and ci != 0:
                continue
            r64 = ref_vtrace.from_logits(
                clip_rho_threshold=c1, clip_pg_rho_threshold=c2,
                **{k: (x.double() if x.is_floating_point() else x) for k, x in tv.items()})
            out[tag + "_vs64"] = r64.vs.numpy()
            out[tag + "_pg64"] = r64.pg_advantages.numpy()
    np.savez_compressed(os.path.join(OUT, "vtrace_random.npz"), **out)


def make_losses():
    # Literal constants of the reference tests (polybeast_loss_functions_test.py:42,66,96-111).
    adv = np.array([1.4, 3.43, 5.2, 0.33])
    ent_logits = np.array([0.0012, 0.321, 0.523, 0.109, 0.416])
    pg_logits = np.array(
        [[[0.206, 0.738, 0.125, 0.484, 0.332], [0.168, 0.504, 0.523, 0.496, 0.626], [0.236, 0.186, 0.627, 0.441, 0.533]],
         [[0.015, 0.904, 0.583, 0.651, 0.855], [0.811, 0.292, 0.061, 0.597, 0.590], [0.999, 0.504, 0.464, 0.077, 0.143]]])
    pg_actions = np.array([[3, 0, 1], [4, 2, 2]])
    pg_adv = np.array([[1.4, 0.31, 0.75], [2.1, 1.5, 0.03]])
    out = dict(adv=adv, ent_logits=ent_logits, pg_logits=pg_logits, pg_actions=pg_actions, pg_adv=pg_adv)
    for mod, tag in ((polybeast_learner, "pl"), (monobeast, "mb")):
        a = torch.from_numpy(adv).requires_grad_()
        v = mod.compute_baseline_loss(a); v.backward()
        out[tag + "_baseline"], out[tag + "_baseline_grad"] = v.item(), a.grad.numpy()
        l = torch.from_numpy(ent_logits).requires_grad_()
        v = mod.compute_entropy_loss(l); v.backward()
        out[tag + "_entropy"], out[tag + "_entropy_grad"] = v.item(), l.grad.numpy()
        l = torch.from_numpy(pg_logits).requires_grad_()
        v = mod.compute_policy_gradient_loss(l, torch.from_numpy(pg_actions), torch.from_numpy(pg_adv)); v.backward()
        out[tag + "_pg"], out[tag + "_pg_grad"] = v.item(), l.grad.numpy()
    np.savez_compressed(os.path.join(OUT, "losses.npz"), **out)


def _stats(t):
    t = t.detach().double().flatten()
    return np.array([t.sum().item(), t.abs().sum().item(), t.norm().item()])


def _flags():
    f = types.SimpleNamespace()
    f.reward_clipping = "abs_one"; f.discounting = 0.99; f.baseline_cost = 0.5; f.entropy_cost = 0.0006
    f.grad_norm_clipping = 40.0; f.learner_device = "cpu"
    return f


GRAD_SAMPLES = 4096  # strided sample of every gradient / updated parameter tensor (large fixtures only)


def sample_index(numel, k=GRAD_SAMPLES):
    """Deterministic strided sample positions shared with the tests (tests/common.py has the same formula)."""
    if numel <= k:
        return np.arange(numel)
    return (np.arange(k, dtype=np.int64) * numel) // k


def _record(out, model, pre_grads, samples=False):
    for n, prm in model.named_parameters():
        out["grad_stats/" + n] = _stats(prm.grad)
        out["grad_head/" + n] = prm.grad.detach().flatten()[:16].numpy().copy()
        if samples:
            idx = torch.from_numpy(sample_index(prm.numel()))
            out["grad_sample/" + n] = prm.grad.detach().flatten()[idx].numpy().copy()
            out["param_sample/" + n] = prm.detach().flatten()[idx].numpy().copy()
        out["param_stats/" + n] = _stats(prm)
        out["param_head/" + n] = prm.detach().flatten()[:16].numpy().copy()
    out["clipped_grad_norm"] = np.sqrt(sum((p.grad.double() ** 2).sum().item() for p in model.parameters()))


def make_learn(net, use_lstm, T, B, seed, fname, clip=40.0, samples=False):
    A = 6
    torch.manual_seed(0)
    batch = LT.synthetic_batch(T, B, A, seed=seed, with_last_action=(net == "atari"))
    flags = _flags(); flags.unroll_length = T; flags.batch_size = B; flags.grad_norm_clipping = clip
    if net == "atari":
        shapes = LT.atarinet_param_shapes(A, use_lstm)
        model = monobeast.AtariNet((4, 84, 84), A, use_lstm)
        actor = monobeast.AtariNet((4, 84, 84), A, use_lstm)
    else:
        shapes = LT.resnet_param_shapes(A, use_lstm)
        model = polybeast_learner.Net(A, use_lstm)
        actor = polybeast_learner.Net(A, use_lstm)
    params = LT.random_params(shapes, seed=seed + 100)
    missing = model.load_state_dict(params, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    for n, prm in model.named_parameters():
        assert tuple(prm.shape) == tuple(shapes[n]), n
    opt = torch.optim.RMSprop(model.parameters(), lr=0.00048, momentum=0, eps=0.01, alpha=0.99)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda e: 1.0)
    state = model.initial_state(B)
    if use_lstm:  # non-zero initial state exercises the carry-in path
        rs = np.random.RandomState(seed + 7)
        state = tuple(torch.from_numpy(rs.randn(*s.shape).astype(np.float32) * 0.1) for s in state)
    out = {}
    # learner outputs (pre-update), via the reference model in eval mode for a clean forward
    with torch.no_grad():
        if net == "atari":
            o, _ = model(batch, state)
            out["policy_logits"], out["baseline"] = o["policy_logits"].numpy(), o["baseline"].numpy()
        else:
            (_, pl, bl), _ = model(dict(frame=batch["frame"], reward=batch["reward"], done=batch["done"]), state)
            out["policy_logits"], out["baseline"] = pl.numpy(), bl.numpy()
        # the V-trace targets learn() computes from these outputs (monobeast.py:245-262 / polybeast_learner.py:332-347),
        # through the reference's own core.vtrace
        rewards = torch.clamp(batch["reward"][1:], -1, 1)
        vt = ref_vtrace.from_logits(
            behavior_policy_logits=batch["policy_logits"][1:], target_policy_logits=torch.from_numpy(out["policy_logits"])[:-1],
            actions=batch["action"][1:], discounts=(~batch["done"][1:]).float() * flags.discounting, rewards=rewards,
            values=torch.from_numpy(out["baseline"])[:-1], bootstrap_value=torch.from_numpy(out["baseline"])[-1])
        out["vs"], out["pg_advantages"] = vt.vs.numpy(), vt.pg_advantages.numpy()
    if net == "atari":
        stats = monobeast.learn(flags, actor, model, batch, state, opt, sched)
    else:
        env = (batch["frame"], batch["reward"], batch["done"], batch["episode_step"], batch["episode_return"])
        agent = (batch["action"], batch["policy_logits"], batch["baseline"])
        q = mock.MagicMock(); q.__iter__.return_value = iter([((env, agent), state)]); q.size.return_value = 0
        stats = {}
        polybeast_learner.learn(flags, q, model, actor, opt, sched, stats, mock.Mock())
    for k in ("total_loss", "pg_loss", "baseline_loss", "entropy_loss"):
        out[k] = np.float64(stats[k])
    _record(out, model, None, samples=samples)
    out["meta"] = np.array([T, B, A, seed, int(use_lstm)])
    out["clip"] = np.float64(clip)
    np.savez_compressed(os.path.join(OUT, fname), **out)
    print(fname, {k: stats[k] for k in ("total_loss", "pg_loss", "baseline_loss", "entropy_loss")},
          "clipped_norm", out["clipped_grad_norm"])


def make_baseline_config():
    """BASELINE.json configs[1] (T=80, B=32, AtariNet +-LSTM) and one GPU's shard of configs[3] (ResNet, T=80, B=8):
    the sizes bench.py runs.  Inputs are regenerated from the seed; outputs carry strided samples of every gradient."""
    make_learn("atari", True, 80, 32, 21, "learn_atari_lstm_T80_B32.npz", samples=True)
    make_learn("atari", False, 80, 32, 22, "learn_atari_T80_B32.npz", samples=True)
    make_learn("resnet", True, 80, 8, 23, "learn_resnet_lstm_T80_B8.npz", samples=True)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    if "--baseline-config" in sys.argv:  # only the large cases (minutes of CPU time)
        make_baseline_config()
        sys.exit(0)
    make_vtrace_fixture()
    make_vtrace_random()
    make_losses()
    make_learn("atari", False, 4, 2, 11, "learn_atari_T4_B2.npz")
    make_learn("atari", True, 4, 2, 12, "learn_atari_lstm_T4_B2.npz")
    make_learn("atari", False, 20, 4, 13, "learn_atari_T20_B4.npz")
    make_learn("atari", True, 20, 4, 14, "learn_atari_lstm_T20_B4.npz")
    make_learn("atari", False, 40, 6, 17, "learn_atari_T40_B6_clip10.npz", clip=10.0)  # grad-norm clip active
    make_learn("resnet", False, 4, 2, 15, "learn_resnet_T4_B2.npz")
    make_learn("resnet", True, 4, 2, 16, "learn_resnet_lstm_T4_B2.npz")
    make_baseline_config()
    print("golden fixtures written to", OUT)
