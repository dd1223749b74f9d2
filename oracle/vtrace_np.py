"""ORACLE (test infrastructure, never shipped, never the thing measured).

NumPy restatement of the reference's V-trace + IMPALA loss arithmetic.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module.

Parity pin: tests/test_oracle_golden.py checks every function here against
(i) the reference's own O(T^2) ground truth (reference tests/vtrace_test.py:46-95)
and (ii) fixtures produced by running the reference itself in the build
container (oracle/make_golden.py -> tests/golden/*.npz).

Reference lines followed (all under /root/reference/torchbeast/):
  action_log_probs        core/vtrace.py:50-55
  from_importance_weights core/vtrace.py:91-139
  from_logits             core/vtrace.py:58-88
  baseline / entropy / pg loss   monobeast.py:107-125 (== polybeast_learner.py:113-131)
  total loss assembly     monobeast.py:245-277 (== polybeast_learner.py:332-361)

`dtype` selects the arithmetic type; float32 mirrors the reference's op order
(so it can be compared bit-for-bit-ish with torch CPU fp32), float64 is the
tolerance-budgeting restatement used for scalar losses.
"""
import collections

import numpy as np

VTraceReturns = collections.namedtuple("VTraceReturns", "vs pg_advantages")
VTraceFromLogitsReturns = collections.namedtuple(
    "VTraceFromLogitsReturns",
    "vs pg_advantages log_rhos behavior_action_log_probs target_action_log_probs",
)
LossTerms = collections.namedtuple(
    "LossTerms",
    "pg_loss baseline_loss entropy_loss total_loss grad_logits grad_values vtrace",
)


def log_softmax(logits):
    """Row-wise log-softmax over the last axis (max-shifted, as ATen does)."""
    z = logits - np.max(logits, axis=-1, keepdims=True)
    return z - np.log(np.sum(np.exp(z), axis=-1, keepdims=True))


def action_log_probs(policy_logits, actions):
    """log pi(a_t | x_t) for the taken action. vtrace.py:50-55."""
    lsm = log_softmax(policy_logits)
    picked = np.take_along_axis(lsm, actions[..., None].astype(np.int64), axis=-1)
    return picked[..., 0]


def from_importance_weights(
    log_rhos,
    discounts,
    rewards,
    values,
    bootstrap_value,
    clip_rho_threshold=1.0,
    clip_pg_rho_threshold=1.0,
    dtype=np.float32,
):
    """Reverse-time V-trace recurrence. vtrace.py:91-139.

    Shapes: log_rhos/discounts/rewards/values [T, B, ...], bootstrap [B, ...];
    trailing dims broadcast like the reference (vtrace_test.py:229-241).
    """
    log_rhos = np.asarray(log_rhos, dtype=dtype)
    discounts = np.asarray(discounts, dtype=dtype)
    rewards = np.asarray(rewards, dtype=dtype)
    values = np.asarray(values, dtype=dtype)
    bootstrap_value = np.asarray(bootstrap_value, dtype=dtype)
    if bootstrap_value.ndim + 1 != values.ndim:
        # The reference surfaces torch.cat's message here (vtrace_test.py:257-260).
        raise RuntimeError(
            "Tensors must have same number of dimensions: got %d and %d"
            % (values.ndim, bootstrap_value.ndim + 1)
        )
    T = discounts.shape[0]
    rhos = np.exp(log_rhos)
    rho_bar = rhos if clip_rho_threshold is None else np.minimum(rhos, dtype(clip_rho_threshold))
    cs = np.minimum(rhos, dtype(1.0))
    v_next = np.concatenate([values[1:], bootstrap_value[None]], axis=0)
    deltas = rho_bar * (rewards + discounts * v_next - values)
    acc = np.zeros_like(bootstrap_value)
    out = [None] * T
    for t in range(T - 1, -1, -1):
        acc = deltas[t] + discounts[t] * cs[t] * acc
        out[t] = acc
    vs = np.stack(out, axis=0) + values
    vs_next = np.concatenate([vs[1:], (np.ones_like(vs[0]) * bootstrap_value)[None]], axis=0)
    rho_pg = rhos if clip_pg_rho_threshold is None else np.minimum(rhos, dtype(clip_pg_rho_threshold))
    pg_adv = rho_pg * (rewards + discounts * vs_next - values)
    return VTraceReturns(vs=vs.astype(dtype), pg_advantages=pg_adv.astype(dtype))


def from_logits(
    behavior_policy_logits,
    target_policy_logits,
    actions,
    discounts,
    rewards,
    values,
    bootstrap_value,
    clip_rho_threshold=1.0,
    clip_pg_rho_threshold=1.0,
    dtype=np.float32,
):
    """vtrace.py:58-88."""
    tlp = action_log_probs(np.asarray(target_policy_logits, dtype=dtype), actions)
    blp = action_log_probs(np.asarray(behavior_policy_logits, dtype=dtype), actions)
    log_rhos = tlp - blp
    r = from_importance_weights(
        log_rhos, discounts, rewards, values, bootstrap_value,
        clip_rho_threshold, clip_pg_rho_threshold, dtype=dtype,
    )
    return VTraceFromLogitsReturns(
        vs=r.vs, pg_advantages=r.pg_advantages, log_rhos=log_rhos,
        behavior_action_log_probs=blp, target_action_log_probs=tlp,
    )


def compute_baseline_loss(advantages):
    """0.5 * sum(adv^2). monobeast.py:107-108."""
    return 0.5 * np.sum(np.square(advantages))


def compute_entropy_loss(logits):
    """sum(p * log p) (negative entropy). monobeast.py:111-115."""
    lsm = log_softmax(logits)
    return np.sum(np.exp(lsm) * lsm)


def compute_policy_gradient_loss(logits, actions, advantages):
    """sum(-log pi(a) * adv). monobeast.py:118-125."""
    return np.sum(-action_log_probs(logits, actions) * advantages)


def loss_gradients(logits, actions, advantages, values, vs, baseline_cost, entropy_cost):
    """Closed-form d(total)/d(logits), d(total)/d(values).

    d/dlogit_j = adv*(p_j - 1[j=a]) + entropy_cost * p_j*(log p_j - sum_k p_k log p_k)
    d/dV       = -baseline_cost * (vs - V)          (vs is detached, vtrace.py:91)
    """
    lsm = log_softmax(logits)
    p = np.exp(lsm)
    onehot = np.zeros_like(p)
    np.put_along_axis(onehot, actions[..., None].astype(np.int64), 1.0, axis=-1)
    ent_row = np.sum(p * lsm, axis=-1, keepdims=True)
    g_logits = advantages[..., None] * (p - onehot) + entropy_cost * p * (lsm - ent_row)
    g_values = -baseline_cost * (vs - values)
    return g_logits, g_values


def impala_loss(
    behavior_policy_logits,
    target_policy_logits,
    actions,
    rewards,
    done,
    values,
    bootstrap_value,
    discounting=0.99,
    baseline_cost=0.5,
    entropy_cost=0.0006,
    reward_clipping="abs_one",
    clip_rho_threshold=1.0,
    clip_pg_rho_threshold=1.0,
    dtype=np.float64,
):
    """The loss block of learn(): monobeast.py:245-277 / polybeast_learner.py:332-361.

    Inputs are already shifted (batch[1:], learner_outputs[:-1]). `done` is bool.
    Returns the three weighted losses, their sum and the closed-form gradients.
    """
    rewards = np.asarray(rewards, dtype=dtype)
    values = np.asarray(values, dtype=dtype)
    tl = np.asarray(target_policy_logits, dtype=dtype)
    if reward_clipping == "abs_one":
        rewards = np.clip(rewards, -1, 1)
    discounts = (~np.asarray(done, dtype=bool)).astype(dtype) * dtype(discounting)
    vt = from_logits(
        behavior_policy_logits, tl, actions, discounts, rewards, values,
        np.asarray(bootstrap_value, dtype=dtype), clip_rho_threshold, clip_pg_rho_threshold, dtype=dtype,
    )
    pg = compute_policy_gradient_loss(tl, actions, vt.pg_advantages)
    bl = baseline_cost * compute_baseline_loss(vt.vs - values)
    en = entropy_cost * compute_entropy_loss(tl)
    g_logits, g_values = loss_gradients(tl, actions, vt.pg_advantages, values, vt.vs, baseline_cost, entropy_cost)
    return LossTerms(pg, bl, en, pg + bl + en, g_logits, g_values, vt)


def ground_truth_vtrace(discounts, log_rhos, rewards, values, bootstrap_value,
                        clip_rho_threshold, clip_pg_rho_threshold):
    """Independent O(T^2) statement of the V-trace definition (IMPALA paper eq. 1):

        v_s = V(x_s) + sum_{t>=s} (prod_{i=s}^{t-1} gamma_i c_i) rho_t (r_t + gamma_t V(x_{t+1}) - V(x_t))

    Same maths as the reference's test helper (tests/vtrace_test.py:46-95), written
    with a running product instead of np.prod over slices. float64 throughout.
    """
    discounts = np.asarray(discounts, np.float64)
    log_rhos = np.asarray(log_rhos, np.float64)
    rewards = np.asarray(rewards, np.float64)
    values = np.asarray(values, np.float64)
    bootstrap_value = np.asarray(bootstrap_value, np.float64)
    T = discounts.shape[0]
    rhos = np.exp(log_rhos)
    cs = np.minimum(rhos, 1.0)
    rho_bar = np.minimum(rhos, clip_rho_threshold) if clip_rho_threshold else rhos
    rho_pg = np.minimum(rhos, clip_pg_rho_threshold) if clip_pg_rho_threshold else rhos
    v_ext = np.concatenate([values, bootstrap_value[None]], axis=0)
    td = rho_bar * (rewards + discounts * v_ext[1:] - values)
    vs = np.empty_like(values)
    for s in range(T):
        w = np.ones_like(values[0])
        tot = values[s].copy()
        for t in range(s, T):
            tot += w * td[t]
            w = w * discounts[t] * cs[t]
        vs[s] = tot
    vs_next = np.concatenate([vs[1:], bootstrap_value[None]], axis=0)
    pg = rho_pg * (rewards + discounts * vs_next - values)
    return VTraceReturns(vs=vs, pg_advantages=pg)
