"""CUDA drop-in for the learner-side symbols of torchbeast/polybeast_learner.py.

Exports (same names/signatures as /root/reference/torchbeast/polybeast_learner.py):
  compute_baseline_loss / compute_entropy_loss / compute_policy_gradient_loss   :113-131
  learn(flags, learner_queue, model, actor_model, optimizer, scheduler, stats, plogger, lock) :295-389
ActorPool / DynamicBatcher / gRPC env server / train() orchestration are out of scope
(SURVEY.md 2); learn() consumes the same ((env_outputs, actor_outputs), initial_agent_state)
nest that learner_queue yields.
"""
from torchbeast_b200.losses import (  # noqa: F401
    compute_baseline_loss,
    compute_entropy_loss,
    compute_policy_gradient_loss,
)

import collections  # noqa: E402
import threading  # noqa: E402

import torch  # noqa: E402

from torchbeast_b200 import learner as _learner  # noqa: E402
from torchbeast_b200.nets import ResNet as Net  # noqa: E402,F401

EnvOutput = collections.namedtuple("EnvOutput", "frame rewards done episode_step episode_return")
AgentOutput = collections.namedtuple("AgentOutput", "action policy_logits baseline")
Batch = collections.namedtuple("Batch", "env agent")


def _map_nest(fn, n):
    """nest.map for the tuple / list / dict nests the reference passes around (nest_pybind.h:61-67: vectors are tuples)."""
    if isinstance(n, torch.Tensor):
        return fn(n)
    if isinstance(n, dict):
        return {k: _map_nest(fn, v) for k, v in n.items()}
    return tuple(_map_nest(fn, v) for v in n)


def inference(flags, inference_batcher, model, lock=threading.Lock()):  # noqa: B008
    """The inference thread body - reference polybeast_learner.py:269-285.

    `inference_batcher` yields DynamicBatcher batches (actorpool.cc:224-340): `batch.get_inputs()` returns
    `(batched_env_outputs, agent_state)` with [T=1, B, ...] leaves, B = 1 .. 512 actors; the forward of the SAME CUDA
    kernels the learner uses runs under `lock` on flags.actor_device, and `batch.set_outputs(((action, policy_logits,
    baseline), core_state))` receives CPU tensors, like the reference.  Works for polybeast's Net (ResNet) and for AtariNet
    (whose `last_action` input is batched_env_outputs[5] if the nest carries one, else zeros)."""
    device = torch.device(getattr(flags, "actor_device", None) or model.flat_params.device)
    with torch.no_grad():
        for batch in inference_batcher:
            batched_env_outputs, agent_state = batch.get_inputs()
            frame, reward, done, *rest = batched_env_outputs
            inputs = dict(frame=frame.to(device, non_blocking=True), reward=reward.to(device, non_blocking=True),
                          done=done.to(device, non_blocking=True))
            if getattr(model, "needs_last_action", False):
                la = rest[2] if len(rest) > 2 else torch.zeros(reward.shape, dtype=torch.int64)
                inputs["last_action"] = la.to(device, non_blocking=True)
            agent_state = _map_nest(lambda t: t.to(device, non_blocking=True), agent_state)
            with lock:
                outputs = model(inputs, agent_state)
            if isinstance(outputs[0], dict):  # AtariNet returns a dict (monobeast.py:626-632): same tuple order as Net
                o = outputs[0]
                outputs = ((o["action"], o["policy_logits"], o["baseline"]), outputs[1])
            batch.set_outputs(_map_nest(lambda t: t.cpu(), outputs))


def _to_device(t, device):
    return t.to(device, non_blocking=True) if isinstance(t, torch.Tensor) else t


def learn(
    flags,
    learner_queue,
    model,
    actor_model,
    optimizer,
    scheduler,
    stats,
    plogger,
    lock=threading.Lock(),  # noqa: B008
):
    """The learner thread body - reference polybeast_learner.py:295-389.

    Consumes the same nest the reference's BatchingQueue yields:
    ((env_outputs, actor_outputs), initial_agent_state) with [T+1, B, ...] CPU (or CUDA) leaves,
    env_outputs = (frame u8, reward, done, episode_step, episode_return), actor_outputs =
    (action, policy_logits, baseline); fills `stats` with the reference's keys and calls plogger.log."""
    for tensors in learner_queue:
        batch, initial_agent_state = tensors
        env_outputs, actor_outputs = batch
        env = EnvOutput._make(list(env_outputs)[:5])
        agent = AgentOutput._make(list(actor_outputs)[:3])
        rollout = dict(frame=env.frame, reward=env.rewards, done=env.done, episode_return=env.episode_return,
                       episode_step=env.episode_step, policy_logits=agent.policy_logits, action=agent.action)
        # host nest (what the reference's BatchingQueue yields) -> pinned slot -> async H2D, OUTSIDE the lock: with
        # num_learner_threads = 2 (polybeast_learner.py:62,505-521) this thread's copy overlaps the other thread's step.
        # _learner.learn takes the lock for the step itself (only one thread learning at a time, pl:313).
        mean_step = torch.mean(env.episode_step[1:].float()).item()
        out = _learner.learn(flags, model, actor_model, rollout, tuple(initial_agent_state), optimizer, scheduler, lock)
        with lock:
            stats["step"] = stats.get("step", 0) + flags.unroll_length * flags.batch_size
            stats["episode_returns"] = out["episode_returns"]
            stats["mean_episode_return"] = out["mean_episode_return"]
            stats["mean_episode_step"] = mean_step
            for k in ("total_loss", "pg_loss", "baseline_loss", "entropy_loss"):
                stats[k] = out[k]
            stats["learner_queue_size"] = learner_queue.size()
            plogger.log(stats)
            if not len(out["episode_returns"]):
                stats["mean_episode_return"] = None  # hide the mean-of-empty NaN, like the reference
