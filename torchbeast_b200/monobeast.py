"""CUDA drop-in for the learner-side symbols of torchbeast/monobeast.py.

Exports (same names/signatures as /root/reference/torchbeast/monobeast.py):
  compute_baseline_loss / compute_entropy_loss / compute_policy_gradient_loss   :107-125
  AtariNet (alias Net)                                                          :545-635
  learn(flags, actor_model, model, batch, initial_agent_state, optimizer, scheduler, lock) :226-296
Actors, shared-memory buffers, env wrappers and the CLI are out of scope.
"""
from torchbeast_b200.losses import (  # noqa: F401
    compute_baseline_loss,
    compute_entropy_loss,
    compute_policy_gradient_loss,
)
