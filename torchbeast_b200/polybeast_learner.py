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
