"""CPU: pin the oracle (oracle/*.py) against outputs of the reference itself
(tests/golden/*.npz, made by oracle/make_golden.py) and against the reference
test-suite's own O(T^2) ground truth."""
import numpy as np
import pytest
import torch

from oracle import learner_torch as LT
from oracle import vtrace_np as VO
from tests.common import CLIPS, RANDOM_CASES, arange_inputs, golden, random_vtrace_inputs


@pytest.mark.parametrize("T,B", [(5, 5), (5, 1), (80, 4)])
def test_scan_matches_reference_and_ground_truth(T, B):
    g = golden("vtrace_fixture.npz")
    v = arange_inputs(T, B)
    tag = "T%d_B%d" % (T, B)
    r32 = VO.from_importance_weights(clip_rho_threshold=3.7, clip_pg_rho_threshold=2.2, **v)
    # reference tolerance: tests/vtrace_test.py:98-99
    np.testing.assert_allclose(r32.vs, g[tag + "_vs"], rtol=1e-6, atol=1e-5)
    np.testing.assert_allclose(r32.pg_advantages, g[tag + "_pg"], rtol=1e-6, atol=1e-5)
    r64 = VO.from_importance_weights(clip_rho_threshold=3.7, clip_pg_rho_threshold=2.2, dtype=np.float64, **v)
    np.testing.assert_allclose(r64.vs, g[tag + "_gt_vs"], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(r64.pg_advantages, g[tag + "_gt_pg"], rtol=1e-9, atol=1e-9)
    gt = VO.ground_truth_vtrace(v["discounts"], v["log_rhos"], v["rewards"], v["values"], v["bootstrap_value"], 3.7, 2.2)
    np.testing.assert_allclose(gt.vs, g[tag + "_gt_vs"], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(gt.pg_advantages, g[tag + "_gt_pg"], rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("case", RANDOM_CASES)
@pytest.mark.parametrize("ci", [0, 1, 2])
def test_from_logits_matches_reference(case, ci):
    (T, B, A), seed = case
    g = golden("vtrace_random.npz")
    v = random_vtrace_inputs(T, B, A, seed)
    c1, c2 = CLIPS[ci]
    tag = "T%d_B%d_A%d_c%d" % (T, B, A, ci)
    r = VO.from_logits(clip_rho_threshold=c1, clip_pg_rho_threshold=c2, **v)
    for name in r._fields:
        if tag + "_" + name in g:
            np.testing.assert_allclose(getattr(r, name), g[tag + "_" + name], rtol=1e-4, atol=1e-4, err_msg=name)  # north_star: 1e-4 fp32
    if tag + "_vs64" in g:
        r64 = VO.from_logits(clip_rho_threshold=c1, clip_pg_rho_threshold=c2, dtype=np.float64, **v)
        np.testing.assert_allclose(r64.vs, g[tag + "_vs64"], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(r64.pg_advantages, g[tag + "_pg64"], rtol=1e-12, atol=1e-12)


def test_loss_constants_match_reference():
    g = golden("losses.npz")
    for tag in ("pl", "mb"):
        np.testing.assert_allclose(VO.compute_baseline_loss(g["adv"]), g[tag + "_baseline"], rtol=1e-12)
        np.testing.assert_allclose(VO.compute_entropy_loss(g["ent_logits"]), g[tag + "_entropy"], rtol=1e-12)
        np.testing.assert_allclose(
            VO.compute_policy_gradient_loss(g["pg_logits"], g["pg_actions"], g["pg_adv"]), g[tag + "_pg"], rtol=1e-12)
        # closed-form gradients == reference autograd
        gl, _ = VO.loss_gradients(g["pg_logits"], g["pg_actions"], g["pg_adv"], 0 * g["pg_adv"], 0 * g["pg_adv"], 0.0, 0.0)
        np.testing.assert_allclose(gl, g[tag + "_pg_grad"], rtol=1e-10, atol=1e-12)
        lg = g["ent_logits"][None, None]
        ge, _ = VO.loss_gradients(lg, np.zeros((1, 1), np.int64), np.zeros((1, 1)), np.zeros((1, 1)), np.zeros((1, 1)), 0.0, 1.0)
        np.testing.assert_allclose(ge[0, 0], g[tag + "_entropy_grad"], rtol=1e-10, atol=1e-12)
        _, gv = VO.loss_gradients(lg, np.zeros((1, 1), np.int64), np.zeros((1, 1)), -g["adv"][None], 0 * g["adv"][None], 1.0, 0.0)
        np.testing.assert_allclose(-gv[0], g[tag + "_baseline_grad"], rtol=1e-12)  # dL/dV = -dL/dadv


def test_rank_mismatch_error_text():
    # reference: tests/vtrace_test.py:243-260
    z = np.zeros
    with pytest.raises(RuntimeError, match="same number of dimensions: got 3 and 2"):
        VO.from_importance_weights(z((3, 2, 1)), z((3, 2, 1)), z((3, 2, 42)), z((3, 2, 42)), z((2,)))
    out = VO.from_importance_weights(z((3, 2, 1)), z((3, 2, 1)), z((3, 2, 42)), z((3, 2, 42)), z((2, 42)))
    assert out.vs.shape == (3, 2, 42)


LEARN_CASES = [
    ("learn_atari_T4_B2.npz", "atari"), ("learn_atari_lstm_T4_B2.npz", "atari"),
    ("learn_atari_T20_B4.npz", "atari"), ("learn_atari_lstm_T20_B4.npz", "atari"),
    ("learn_atari_T40_B6_clip10.npz", "atari"),
    ("learn_resnet_T4_B2.npz", "resnet"), ("learn_resnet_lstm_T4_B2.npz", "resnet"),
]


def run_oracle_learn(fname, net, dtype=torch.float32):
    g = golden(fname)
    T, B, A, seed, use_lstm = [int(x) for x in g["meta"]]
    batch = LT.synthetic_batch(T, B, A, seed=seed, with_last_action=(net == "atari"))
    shapes = LT.atarinet_param_shapes(A, bool(use_lstm)) if net == "atari" else LT.resnet_param_shapes(A, bool(use_lstm))
    p = LT.random_params(shapes, seed=seed + 100, dtype=dtype)
    state = ()
    if use_lstm:
        L, H = (2, 512 + A + 1) if net == "atari" else (1, 256)
        rs = np.random.RandomState(seed + 7)
        state = tuple(torch.from_numpy(rs.randn(L, B, H).astype(np.float32) * 0.1).to(dtype) for _ in range(2))
    hyper = type("H", (LT.Hyper,), dict(grad_norm_clipping=float(g["clip"])))
    return g, LT.learner_step(p, batch, state, hyper=hyper, net=net, num_actions=A)


@pytest.mark.parametrize("fname,net", LEARN_CASES)
def test_learn_step_matches_reference(fname, net):
    g, o = run_oracle_learn(fname, net)
    np.testing.assert_allclose(o["policy_logits"].numpy(), g["policy_logits"], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(o["baseline"].numpy(), g["baseline"], rtol=1e-4, atol=2e-5)
    for k in ("total_loss", "pg_loss", "baseline_loss", "entropy_loss"):
        np.testing.assert_allclose(float(o[k]), float(g[k]), rtol=2e-5, atol=2e-5, err_msg=k)
    coef = min(1.0, float(g["clip"]) / (float(o["grad_norm"]) + 1e-6))
    np.testing.assert_allclose(float(o["grad_norm"]) * coef, float(g["clipped_grad_norm"]), rtol=1e-4)
    for n, gr in o["grads"].items():
        head = (gr * coef).flatten()[:16].numpy()
        ref = g["grad_head/" + n]
        scale = max(float(g["grad_stats/" + n][2]), 1e-6)
        np.testing.assert_allclose(head, ref, rtol=2e-3, atol=2e-4 * scale, err_msg=n)
        np.testing.assert_allclose(float((gr * coef).double().norm()), float(g["grad_stats/" + n][2]), rtol=1e-3, atol=1e-6, err_msg=n)
        np.testing.assert_allclose(o["params"][n].flatten()[:16].numpy(), g["param_head/" + n], rtol=1e-4, atol=1e-5, err_msg=n)
        np.testing.assert_allclose(float(o["params"][n].double().norm()), float(g["param_stats/" + n][2]), rtol=1e-5, err_msg=n)


@pytest.mark.parametrize("T,B", [(5, 5), (5, 1), (80, 4)])
def test_c_oracle_scan_matches_reference(T, B):
    from oracle import c_api
    g = golden("vtrace_fixture.npz")
    v = arange_inputs(T, B)
    vs, pg = c_api.vtrace_scan(v["log_rhos"], v["discounts"], v["rewards"], v["values"], v["bootstrap_value"], 3.7, 2.2)
    tag = "T%d_B%d" % (T, B)
    np.testing.assert_allclose(vs, g[tag + "_vs"], rtol=1e-6, atol=1e-5)
    np.testing.assert_allclose(pg, g[tag + "_pg"], rtol=1e-6, atol=1e-5)


def test_c_oracle_loss_matches_numpy_oracle():
    from oracle import c_api
    rs = np.random.RandomState(5)
    T, B, A = 33, 7, 6
    v = random_vtrace_inputs(T, B, A, 9)
    done = rs.rand(T, B) < 0.1
    rew = (rs.randn(T, B) * 2).astype(np.float32)
    c = c_api.impala_loss(v["behavior_policy_logits"], v["target_policy_logits"], v["actions"], rew, done,
                          v["values"], v["bootstrap_value"])
    o = VO.impala_loss(v["behavior_policy_logits"], v["target_policy_logits"], v["actions"], rew, done,
                       v["values"], v["bootstrap_value"], dtype=np.float64)
    np.testing.assert_allclose(c["vs"], o.vtrace.vs, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(c["pg_advantages"], o.vtrace.pg_advantages, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(c["losses"], [o.pg_loss, o.baseline_loss, o.entropy_loss], rtol=1e-5)
    np.testing.assert_allclose(c["grad_logits"], o.grad_logits, rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(c["grad_values"], o.grad_values, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("fname,net", [("learn_atari_lstm_T80_B32.npz", "atari"), ("learn_atari_T80_B32.npz", "atari"),
                                       ("learn_resnet_lstm_T80_B8.npz", "resnet")])
def test_learn_step_matches_reference_at_the_baseline_config(fname, net):
    """The oracle port against the reference's own output AT the sizes bench.py runs (BASELINE configs[1]: T=80, B=32; one
    GPU's shard of configs[3]: ResNet T=80, B=8): outputs, V-trace targets, losses, 4096 strided gradient samples per tensor.
    (the learn-step fixtures stopped at T=40, B=6.)"""
    from tests.common import sample_index
    g, o = run_oracle_learn(fname, net)
    np.testing.assert_allclose(o["policy_logits"].numpy(), g["policy_logits"], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(o["baseline"].numpy(), g["baseline"], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(o["vs"].numpy(), g["vs"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(o["pg_advantages"].numpy(), g["pg_advantages"], rtol=1e-4, atol=1e-4)
    for k in ("total_loss", "pg_loss", "baseline_loss", "entropy_loss"):
        np.testing.assert_allclose(float(o[k]), float(g[k]), rtol=2e-5, atol=2e-5, err_msg=k)
    coef = min(1.0, float(g["clip"]) / (float(o["grad_norm"]) + 1e-6))
    for n, gr in o["grads"].items():
        idx = torch.from_numpy(sample_index(gr.numel()))
        got = (gr * coef).flatten()[idx].double()
        ref = torch.from_numpy(g["grad_sample/" + n]).double()
        rel = float((got - ref).norm() / ref.norm().clamp_min(1e-30))
        assert rel < 6e-3, (n, rel)  # fp32 summation order moves ReLU ties at this size (see tests/test_learner_baseline_gpu.py)
