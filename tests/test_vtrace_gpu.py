"""GPU parity: torchbeast_b200.core.vtrace / losses (CUDA, via the C-ABI) against
 (a) the golden fixtures produced by the reference itself (tests/golden),
 (b) the oracle (oracle/vtrace_np.py, oracle/vtrace_c.c) on seeded inputs,
 (c) size-independent properties at BASELINE.json's full sizes.
The cases mirror /root/reference/tests/vtrace_test.py and polybeast_loss_functions_test.py.
Tolerances: element-wise atol 1e-4 fp32 (north_star); the reference's own rtol 1e-6/atol 1e-5
where the fixture is small; scalar losses rtol 1e-6 against the fp64 oracle."""
import numpy as np
import pytest
import torch

from oracle import vtrace_np as VO
from tests.common import CLIPS, RANDOM_CASES, arange_inputs, golden, random_vtrace_inputs

pytestmark = pytest.mark.gpu


def cu(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


@pytest.fixture(scope="module")
def vtrace():
    from torchbeast_b200.core import vtrace as v
    return v


# ---- reference tests/vtrace_test.py:103-132 ---------------------------------------------
@pytest.mark.parametrize("batch_size", [2, 1])
def test_action_log_probs(vtrace, batch_size):
    seq_len, num_actions = 7, 3
    logits = np.arange(seq_len * batch_size * num_actions, dtype=np.float32).reshape(seq_len, batch_size, num_actions) + 10
    actions = np.random.RandomState(0).randint(0, num_actions, size=(seq_len, batch_size)).astype(np.int64)
    out = vtrace.action_log_probs(cu(logits), cu(actions))
    sm = np.exp(logits) / np.sum(np.exp(logits), axis=-1, keepdims=True)
    gt = np.take_along_axis(np.log(sm), actions[..., None], -1)[..., 0]
    np.testing.assert_allclose(out.cpu().numpy(), gt, rtol=1e-6, atol=1e-5)
    out64 = vtrace.action_log_probs(cu(logits.astype(np.float64)), cu(actions))
    assert out64.dtype == torch.float64
    np.testing.assert_allclose(out64.cpu().numpy(), VO.action_log_probs(logits.astype(np.float64), actions), rtol=1e-12)


def test_action_log_probs_grad(vtrace):
    rs = np.random.RandomState(1)
    logits = cu(rs.randn(5, 3, 6)).double().requires_grad_()
    actions = cu(rs.randint(0, 6, size=(5, 3)).astype(np.int64))
    w = cu(rs.randn(5, 3))
    (vtrace.action_log_probs(logits, actions) * w).sum().backward()
    ref = logits.detach().clone().requires_grad_()
    (torch.log_softmax(ref, -1).gather(-1, actions.unsqueeze(-1)).squeeze(-1) * w).sum().backward()
    np.testing.assert_allclose(logits.grad.cpu().numpy(), ref.grad.cpu().numpy(), rtol=1e-10, atol=1e-12)


# ---- reference tests/vtrace_test.py:136-168 + BASELINE.json configs[0] (T=80,B=4) --------
@pytest.mark.parametrize("T,B", [(5, 5), (5, 1), (80, 4)])
def test_vtrace_fixture(vtrace, T, B):
    g = golden("vtrace_fixture.npz")
    v = arange_inputs(T, B)
    out = vtrace.from_importance_weights(**{k: cu(x) for k, x in v.items()}, clip_rho_threshold=3.7, clip_pg_rho_threshold=2.2)
    tag = "T%d_B%d" % (T, B)
    for got, name in ((out.vs, "vs"), (out.pg_advantages, "pg")):
        # reference's own output and the reference test's O(T^2) ground truth
        np.testing.assert_allclose(got.cpu().numpy(), g[tag + "_" + name], rtol=1e-6, atol=1e-5)
        np.testing.assert_allclose(got.cpu().numpy(), g[tag + "_gt_" + name], rtol=2e-6, atol=1e-4)
    out64 = vtrace.from_importance_weights(**{k: cu(x.astype(np.float64)) for k, x in v.items()},
                                           clip_rho_threshold=3.7, clip_pg_rho_threshold=2.2)
    np.testing.assert_allclose(out64.vs.cpu().numpy(), g[tag + "_gt_vs"], rtol=1e-10, atol=1e-10)
    np.testing.assert_allclose(out64.pg_advantages.cpu().numpy(), g[tag + "_gt_pg"], rtol=1e-10, atol=1e-10)


# ---- reference tests/vtrace_test.py:170-227, plus golden outputs of the reference ------
@pytest.mark.parametrize("case", RANDOM_CASES)
@pytest.mark.parametrize("ci", [0, 1, 2])
def test_from_logits_golden(vtrace, case, ci):
    (T, B, A), seed = case
    g = golden("vtrace_random.npz")
    v = random_vtrace_inputs(T, B, A, seed)
    c1, c2 = CLIPS[ci]
    tv = {k: cu(x) for k, x in v.items()}
    out = vtrace.from_logits(clip_rho_threshold=c1, clip_pg_rho_threshold=c2, **tv)
    tag = "T%d_B%d_A%d_c%d" % (T, B, A, ci)
    for name in out._fields:
        if tag + "_" + name in g:
            np.testing.assert_allclose(getattr(out, name).cpu().numpy(), g[tag + "_" + name], rtol=1e-4, atol=1e-4, err_msg=name)
    # from_logits == from_importance_weights o action_log_probs (the reference's consistency test)
    tlp = vtrace.action_log_probs(tv["target_policy_logits"], tv["actions"])
    blp = vtrace.action_log_probs(tv["behavior_policy_logits"], tv["actions"])
    iw = vtrace.from_importance_weights(tlp - blp, tv["discounts"], tv["rewards"], tv["values"], tv["bootstrap_value"], c1, c2)
    np.testing.assert_allclose(iw.vs.cpu().numpy(), out.vs.cpu().numpy(), rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(iw.pg_advantages.cpu().numpy(), out.pg_advantages.cpu().numpy(), rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(tlp.cpu().numpy(), out.target_action_log_probs.cpu().numpy(), rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(blp.cpu().numpy(), out.behavior_action_log_probs.cpu().numpy(), rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose((tlp - blp).cpu().numpy(), out.log_rhos.cpu().numpy(), rtol=1e-6, atol=1e-6)
    # float64 composed path against the fp64 golden
    if tag + "_vs64" in g:
        o64 = vtrace.from_logits(clip_rho_threshold=c1, clip_pg_rho_threshold=c2,
                                 **{k: (x.double() if x.is_floating_point() else x) for k, x in tv.items()})
        np.testing.assert_allclose(o64.vs.cpu().numpy(), g[tag + "_vs64"], rtol=1e-11, atol=1e-11)
        np.testing.assert_allclose(o64.pg_advantages.cpu().numpy(), g[tag + "_pg64"], rtol=1e-11, atol=1e-11)


# ---- reference tests/vtrace_test.py:229-260 ---------------------------------------------
def test_higher_rank_inputs(vtrace):
    T, B = 3, 2
    z = lambda *s: torch.zeros(*s, device="cuda")
    out = vtrace.from_importance_weights(z(T, B, 1), z(T, B, 1), z(T, B, 42), z(T, B, 42), z(B, 42))
    assert tuple(out.vs.shape) == (T, B, 42)
    rs = np.random.RandomState(3)
    lr, dc = rs.randn(T, B, 1) * 0.3, rs.rand(T, B, 1)
    rw, va, bs = rs.randn(T, B, 5), rs.randn(T, B, 5), rs.randn(B, 5)
    got = vtrace.from_importance_weights(cu(lr), cu(dc), cu(rw), cu(va), cu(bs))
    ref = VO.from_importance_weights(lr, dc, rw, va, bs, dtype=np.float64)
    np.testing.assert_allclose(got.vs.cpu().numpy(), ref.vs, rtol=1e-11, atol=1e-11)
    np.testing.assert_allclose(got.pg_advantages.cpu().numpy(), ref.pg_advantages, rtol=1e-11, atol=1e-11)


def test_inconsistent_rank_inputs(vtrace):
    T, B = 3, 2
    z = lambda *s: torch.zeros(*s, device="cuda")
    with pytest.raises(RuntimeError, match="same number of dimensions: got 3 and 2"):
        vtrace.from_importance_weights(z(T, B, 1), z(T, B, 1), z(T, B, 42), z(T, B, 42), z(B))


# ---- oracle sweep: ragged / edge shapes (every T-split W, partial tiles, empty) ----------
@pytest.mark.parametrize("T,B", [(1, 1), (1, 33), (2, 31), (3, 32), (4, 65), (7, 1), (17, 5), (64, 100), (81, 32),
                                 (160, 3), (600, 128), (33, 4800), (9, 40000)])
@pytest.mark.parametrize("clips", [(1.0, 1.0), (None, None), (0.5, 7.0)])
def test_scan_vs_oracle(vtrace, T, B, clips):
    rs = np.random.RandomState(T * 1000 + B)
    lr = (rs.randn(T, B) * 0.5).astype(np.float32)
    dc = (0.99 * (rs.rand(T, B) > 0.05)).astype(np.float32)
    rw = np.clip(rs.randn(T, B), -1, 1).astype(np.float32)
    va = rs.randn(T, B).astype(np.float32)
    bs = rs.randn(B).astype(np.float32)
    got = vtrace.from_importance_weights(cu(lr), cu(dc), cu(rw), cu(va), cu(bs), *clips)
    ref = VO.from_importance_weights(lr, dc, rw, va, bs, *clips, dtype=np.float64)
    np.testing.assert_allclose(got.vs.cpu().numpy(), ref.vs, rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(got.pg_advantages.cpu().numpy(), ref.pg_advantages, rtol=1e-5, atol=1e-4)
    from oracle import c_api
    cvs, cpg = c_api.vtrace_scan(lr, dc, rw, va, bs, *clips)
    np.testing.assert_allclose(got.vs.cpu().numpy(), cvs, rtol=1e-5, atol=2e-5)
    np.testing.assert_allclose(got.pg_advantages.cpu().numpy(), cpg, rtol=1e-5, atol=2e-5)


def test_empty_inputs(vtrace):
    z = lambda *s: torch.zeros(*s, device="cuda")
    out = vtrace.from_importance_weights(z(0, 4), z(0, 4), z(0, 4), z(0, 4), z(4))
    assert tuple(out.vs.shape) == (0, 4)
    out = vtrace.from_importance_weights(z(5, 0), z(5, 0), z(5, 0), z(5, 0), z(0))
    assert tuple(out.pg_advantages.shape) == (5, 0)


# ---- size-independent properties at the full BASELINE sizes -----------------------------
@pytest.mark.parametrize("T,B", [(80, 32), (600, 128), (80, 1 << 16)])
def test_scan_properties_full_size(vtrace, T, B):
    g = torch.Generator(device="cuda").manual_seed(0)
    lr = 0.5 * torch.randn(T, B, device="cuda", generator=g)
    dc = 0.99 * (torch.rand(T, B, device="cuda", generator=g) > 0.01).float()
    rw = torch.randn(T, B, device="cuda", generator=g).clamp(-1, 1)
    va = torch.randn(T, B, device="cuda", generator=g)
    bs = torch.randn(B, device="cuda", generator=g)
    out = vtrace.from_importance_weights(lr, dc, rw, va, bs)
    # (1) the outputs satisfy the defining recurrence: vs_t - V_t = delta_t + g_t c_t (vs_{t+1} - V_{t+1})
    rho = lr.exp()
    v_next = torch.cat([va[1:], bs[None]], 0)
    vs_next = torch.cat([out.vs[1:], bs[None]], 0)
    delta = rho.clamp(max=1.0) * (rw + dc * v_next - va)
    rhs = delta + dc * rho.clamp(max=1.0) * (vs_next - v_next) + va
    assert torch.allclose(out.vs, rhs, rtol=1e-5, atol=1e-4)
    assert torch.allclose(out.pg_advantages, rho.clamp(max=1.0) * (rw + dc * vs_next - va), rtol=1e-5, atol=1e-4)
    # (2) on-policy, no clipping-active case: vs equals the discounted n-step return (rho = 1)
    on = vtrace.from_importance_weights(torch.zeros_like(lr), dc, rw, va, bs)
    ret = bs.clone()
    rets = []
    for t in range(T - 1, -1, -1):
        ret = rw[t] + dc[t] * ret
        rets.append(ret)
    rets = torch.stack(rets[::-1], 0)
    assert torch.allclose(on.vs, rets, rtol=1e-4, atol=1e-4)
    # (3) column independence: permuting batch columns permutes the outputs bit-for-bit
    perm = torch.randperm(B, device="cuda", generator=g)
    outp = vtrace.from_importance_weights(lr[:, perm].contiguous(), dc[:, perm].contiguous(), rw[:, perm].contiguous(),
                                          va[:, perm].contiguous(), bs[perm].contiguous())
    assert torch.equal(outp.vs, out.vs[:, perm])
    assert torch.equal(outp.pg_advantages, out.pg_advantages[:, perm])
    # (4) run-to-run determinism
    again = vtrace.from_importance_weights(lr, dc, rw, va, bs)
    assert torch.equal(again.vs, out.vs) and torch.equal(again.pg_advantages, out.pg_advantages)


# ---- the three loss functions: reference tests/polybeast_loss_functions_test.py ----------
@pytest.mark.parametrize("mod", ["monobeast", "polybeast_learner"])
def test_loss_functions_values_and_grads(mod):
    import importlib
    m = importlib.import_module("torchbeast_b200." + mod)
    g = golden("losses.npz")
    tag = "mb" if mod == "monobeast" else "pl"
    adv = cu(g["adv"]).requires_grad_()
    v = m.compute_baseline_loss(adv); v.backward()
    assert v.shape == ()
    np.testing.assert_allclose(v.item(), g[tag + "_baseline"], rtol=1e-6, atol=1e-5)
    np.testing.assert_allclose(adv.grad.cpu().numpy(), g[tag + "_baseline_grad"], rtol=1e-6, atol=1e-5)
    lg = cu(g["ent_logits"]).requires_grad_()
    v = m.compute_entropy_loss(lg); v.backward()
    np.testing.assert_allclose(v.item(), g[tag + "_entropy"], rtol=1e-6, atol=1e-5)
    np.testing.assert_allclose(lg.grad.cpu().numpy(), g[tag + "_entropy_grad"], rtol=1e-6, atol=1e-5)
    lg = cu(g["pg_logits"]).requires_grad_()
    advt = cu(g["pg_adv"]).requires_grad_()
    v = m.compute_policy_gradient_loss(lg, cu(g["pg_actions"]), advt); v.backward()
    assert v.shape == ()
    np.testing.assert_allclose(v.item(), g[tag + "_pg"], rtol=1e-6, atol=1e-5)
    np.testing.assert_allclose(lg.grad.cpu().numpy(), g[tag + "_pg_grad"], rtol=1e-6, atol=1e-5)
    assert advt.grad is None  # advantages get no gradient (reference test :165-177)
    # float32 inputs too
    v32 = m.compute_policy_gradient_loss(cu(g["pg_logits"].astype(np.float32)), cu(g["pg_actions"]), cu(g["pg_adv"].astype(np.float32)))
    np.testing.assert_allclose(v32.item(), g[tag + "_pg"], rtol=1e-5)


# ---- fused loss kernel (what learn() launches) vs the fp64 oracle --------------------------
@pytest.mark.parametrize("T,B,A", [(80, 32, 6), (5, 3, 3), (20, 70, 18), (600, 128, 6), (33, 5, 11), (80, 4096, 6), (1, 1, 2)])
def test_fused_loss_vs_oracle(T, B, A):
    from torchbeast_b200 import learner
    rs = np.random.RandomState(T + B + A)
    v = random_vtrace_inputs(T, B, A, T * 7 + B)
    done = rs.rand(T, B) < 0.05
    rew = (rs.randn(T, B) * 2).astype(np.float32)
    o = VO.impala_loss(v["behavior_policy_logits"], v["target_policy_logits"], v["actions"], rew, done,
                       v["values"], v["bootstrap_value"], dtype=np.float64)
    r = learner.impala_loss_fwd_bwd(cu(v["behavior_policy_logits"]), cu(v["target_policy_logits"]), cu(v["actions"]),
                                    cu(rew), cu(done), cu(v["values"]), cu(v["bootstrap_value"]))
    np.testing.assert_allclose(r.vs.cpu().numpy(), o.vtrace.vs, rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(r.pg_advantages.cpu().numpy(), o.vtrace.pg_advantages, rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(r.log_rhos.cpu().numpy(), o.vtrace.log_rhos, rtol=1e-5, atol=1e-5)
    losses = r.losses.cpu().numpy().astype(np.float64)
    # scalars: rtol 1e-6 vs fp64 (SURVEY.md section 7 "hard parts"), abs floor for cancellation in pg_loss
    scale = np.abs(o.vtrace.pg_advantages).sum() + 1.0
    np.testing.assert_allclose(losses[0], o.pg_loss, rtol=2e-6, atol=2e-7 * scale)
    np.testing.assert_allclose(losses[1], o.baseline_loss, rtol=2e-6)
    np.testing.assert_allclose(losses[2], o.entropy_loss, rtol=2e-6)
    np.testing.assert_allclose(losses[3], o.total_loss, rtol=2e-6, atol=2e-7 * scale)
    gl = r.grad_logits.cpu().numpy(); gv = r.grad_values.cpu().numpy()
    assert gl.shape == (T + 1, B, A) and gv.shape == (T + 1, B)
    np.testing.assert_allclose(gl[:-1], o.grad_logits, rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(gv[:-1], o.grad_values, rtol=1e-4, atol=1e-5)
    assert not gl[-1].any() and not gv[-1].any()
    # C restatement agrees too
    from oracle import c_api
    c = c_api.impala_loss(v["behavior_policy_logits"], v["target_policy_logits"], v["actions"], rew, done,
                          v["values"], v["bootstrap_value"])
    np.testing.assert_allclose(gl[:-1], c["grad_logits"], rtol=1e-4, atol=2e-5)
    # determinism
    r2 = learner.impala_loss_fwd_bwd(cu(v["behavior_policy_logits"]), cu(v["target_policy_logits"]), cu(v["actions"]),
                                     cu(rew), cu(done), cu(v["values"]), cu(v["bootstrap_value"]))
    assert torch.equal(r2.losses, r.losses) and torch.equal(r2.grad_logits, r.grad_logits)


def test_out_of_range_action_is_loud_not_out_of_bounds(vtrace):
    """an action index outside [0, A) (the reference raises 'Target out of bounds') must neither read out of
    bounds nor silently pick a logit: the affected elements are NaN, everything else is untouched."""
    logits = torch.randn(4, 3, 6, device="cuda")
    actions = torch.randint(0, 6, (4, 3), device="cuda")
    good = vtrace.action_log_probs(logits, actions)
    bad_actions = actions.clone()
    bad_actions[1, 2] = 6
    bad_actions[3, 0] = -1
    bad = vtrace.action_log_probs(logits, bad_actions)
    assert torch.isnan(bad[1, 2]) and torch.isnan(bad[3, 0])
    mask = torch.ones_like(good, dtype=torch.bool)
    mask[1, 2] = False; mask[3, 0] = False
    assert torch.equal(bad[mask], good[mask])
    big = torch.randn(4, 3, 40, device="cuda")  # runtime-A path
    ba = torch.randint(0, 40, (4, 3), device="cuda"); ba[0, 0] = 40
    assert torch.isnan(vtrace.action_log_probs(big, ba)[0, 0])
