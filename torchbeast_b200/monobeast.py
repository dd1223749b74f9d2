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

import threading  # noqa: E402

import torch  # noqa: E402

from torchbeast_b200 import learner as _learner  # noqa: E402
from torchbeast_b200.nets import AtariNet  # noqa: E402,F401

Net = AtariNet


def learn(
    flags,
    actor_model,
    model,
    batch,
    initial_agent_state,
    optimizer,
    scheduler,
    lock=threading.Lock(),  # noqa: B008
):
    """Performs a learning (optimization) step - reference monobeast.py:226-296.

    Same arguments and returned stats dict.  `model` is a torchbeast_b200 network, `batch` the dict of
    [T+1, B, ...] tensors get_batch() produces (monobeast.py:194-223): CUDA tensors as in the reference, or HOST
    tensors (the buffers themselves) - those are staged through the pinned RolloutStager outside the lock.
    flags.cuda_graph / TB_CUDA_GRAPH=1 replays the whole device side as one CUDA graph (learner.learn)."""
    return _learner.learn(flags, model, actor_model, batch, initial_agent_state, optimizer, scheduler, lock)
