"""Fused pieces of the learner step shared by monobeast.learn and polybeast_learner.learn."""
import collections

import torch

from torchbeast_b200 import _lib

ImpalaLoss = collections.namedtuple(
    "ImpalaLoss",
    "vs pg_advantages log_rhos behavior_action_log_probs target_action_log_probs losses grad_logits grad_values",
)


@torch.no_grad()
def impala_loss_fwd_bwd(
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
    with_grads=True,
):
    """The whole loss block of learn() in ONE kernel launch (tb_impala_loss_fwd_bwd_f32).

    Inputs are the shifted [T,B,...] slices (batch[1:], learner_outputs[:-1]); `done` is a
    bool/uint8 tensor.  Replaces monobeast.py:245-277 / polybeast_learner.py:332-361 and the
    autograd graph behind them: returns vs / pg_advantages / log-rhos, losses =
    [pg, baseline_cost*baseline, entropy_cost*entropy, total] and d total / d(target logits,
    values) laid out as [T+1,B,...] with a zero bootstrap row, ready for the network backward.
    """
    tensors = (behavior_policy_logits, target_policy_logits, actions, rewards, done, values, bootstrap_value)
    _lib.require_cuda(*tensors)
    if reward_clipping not in ("abs_one", "none"):
        raise ValueError("reward_clipping must be 'abs_one' or 'none'")
    T, B, A = target_policy_logits.shape
    dev = values.device
    f32 = torch.float32
    bl = behavior_policy_logits.to(f32).contiguous()
    tl = target_policy_logits.detach().to(f32).contiguous()
    ac = actions.to(torch.int64).contiguous()
    rw = rewards.to(f32).contiguous()
    if done.dtype == torch.bool:
        dn = done.contiguous().view(torch.uint8)
        disc = None
    elif done.dtype == torch.uint8:
        # `~done` on uint8 is a bitwise NOT in the reference: discount = (255-done)*gamma
        dn, disc = None, ((~done).to(f32) * discounting).contiguous()
    else:
        raise _lib.TorchBeastB200Error("done must be bool or uint8")
    va = values.detach().to(f32).contiguous()
    bs = bootstrap_value.detach().to(f32).contiguous()
    outs = [torch.empty((T, B), dtype=f32, device=dev) for _ in range(5)]
    losses = torch.empty(4, dtype=f32, device=dev)
    gl = torch.empty((T + 1, B, A), dtype=f32, device=dev) if with_grads else None
    gv = torch.empty((T + 1, B), dtype=f32, device=dev) if with_grads else None
    _lib.check(
        _lib.lib().tb_impala_loss_fwd_bwd_f32(
            _lib.ptr(bl), _lib.ptr(tl), _lib.ptr(ac), _lib.ptr(rw), _lib.ptr(dn), _lib.ptr(disc),
            _lib.ptr(va), _lib.ptr(bs), T, B, A, float(discounting), float(baseline_cost), float(entropy_cost),
            int(reward_clipping == "abs_one"), _lib.clip_arg(clip_rho_threshold), _lib.clip_arg(clip_pg_rho_threshold),
            *[_lib.ptr(o) for o in outs], _lib.ptr(losses), _lib.ptr(gl), _lib.ptr(gv), 1,
            _lib.ptr(_lib.workspace()), _lib.stream_ptr()),
        "tb_impala_loss_fwd_bwd_f32")
    vs, pg, lr, blp, tlp = outs
    return ImpalaLoss(vs, pg, lr, blp, tlp, losses, gl, gv)
