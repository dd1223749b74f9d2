"""GPU: N1 (RolloutStager), the host-batch path of monobeast.learn / polybeast_learner.learn, the graphed step, two learner
threads sharing one model (polybeast_learner.py:505-521) and the synthetic actor pipeline (configs[2])."""
import threading
import types

import numpy as np
import pytest
import torch

from oracle import learner_torch as LT
from tests.test_learner_gpu import build_case, flags_for, to_cuda

pytestmark = pytest.mark.gpu


def test_stager_round_trip_and_in_place_submit():
    from torchbeast_b200 import staging
    dev = torch.device("cuda", 0)
    spec = staging.spec_for(5, 3, 6)
    st = staging.RolloutStager(spec, dev, depth=2)
    batch = LT.synthetic_batch(5, 3, 6, seed=1)
    # (a) drop-in path: a pageable host batch is copied into a pinned slot
    i = st.put(batch)
    d, j = st.get()
    assert i == j
    for k, v in batch.items():
        assert torch.equal(d[k].cpu(), v), k
    st.release(j)
    # (b) actors write columns in place; put() recognises the slot's own tensors (no memcpy) and submits it
    i = st.acquire_host()
    for b in range(3):
        col = st.column(i, b)
        for k, v in batch.items():
            col[k].copy_(v[:, b])
    assert st.put(st.host[i]) == i
    d, j = st.get()
    for k, v in batch.items():
        assert torch.equal(d[k].cpu(), v), k
    st.release(j)
    assert st.h2d_bytes == sum(v.numel() * v.element_size() for v in batch.values())


@pytest.mark.parametrize("graph", [False, True])
def test_learn_from_host_batch_matches_device_batch(graph):
    """monobeast.learn(host rollout) == monobeast.learn(device rollout): staging (and the CUDA-graph replay) change nothing."""
    from torchbeast_b200 import monobeast
    res = []
    for host in (False, True):
        g, model, actor, batch, params, state, opt, sched = build_case("learn_atari_lstm_T20_B4.npz")
        flags = flags_for(g)
        flags.cuda_graph = graph and host
        b = batch if host else to_cuda(batch)
        st = tuple(s if host else s.cuda() for s in state)
        s1 = monobeast.learn(flags, actor, model, b, st, opt, sched)
        s2 = monobeast.learn(flags, actor, model, b, st, opt, sched)
        res.append((s1["total_loss"], s2["total_loss"], model.flat_params.clone(), s1["episode_returns"]))
        np.testing.assert_allclose(s1["total_loss"], float(g["total_loss"]), rtol=2e-5)
        assert torch.equal(actor.flat_params, model.flat_params)
    assert res[0][0] == res[1][0] and res[0][1] == res[1][1] and torch.equal(res[0][2], res[1][2])
    assert res[0][3] == res[1][3]


def test_two_learner_threads_share_one_model():
    """B5: two threads run polybeast_learner.learn on one model/optimizer/lock; every rollout is consumed exactly once and
    the result equals the same rollouts learned sequentially (the lock serialises the steps; order is immaterial here
    because every queue item is the same rollout)."""
    from torchbeast_b200 import actors, monobeast, optim, polybeast_learner
    T, B, A = 6, 4, 6
    batch = LT.synthetic_batch(T, B, A, seed=5, with_last_action=False)
    env = (batch["frame"], batch["reward"], batch["done"], batch["episode_step"], batch["episode_return"])
    agent = (batch["action"], batch["policy_logits"], batch["baseline"])
    flags = types.SimpleNamespace(reward_clipping="abs_one", discounting=0.99, baseline_cost=0.5, entropy_cost=0.0006,
                                  grad_norm_clipping=40.0, unroll_length=T, batch_size=B)
    outs = []
    for nthreads in (1, 2):
        model = monobeast.AtariNet((4, 84, 84), A, False)
        model.reset_parameters_like_torch(seed=3)
        actor = monobeast.AtariNet((4, 84, 84), A, False)
        opt = optim.RMSprop(model, lr=0.00048, momentum=0, eps=0.01, alpha=0.99)
        q = actors.LearnerQueue()
        for _ in range(6):
            q.put(((env, agent), ()))
        q.close()
        stats, logged, lock = {}, [], threading.Lock()

        class Log:
            def log(self, s):
                logged.append(s["total_loss"])
        ths = [threading.Thread(target=polybeast_learner.learn, args=(flags, q, model, actor, opt, None, stats, Log(), lock))
               for _ in range(nthreads)]
        for t in ths:
            t.start()
        for t in ths:
            t.join(timeout=120)
            assert not t.is_alive()
        assert len(logged) == 6 and stats["step"] == 6 * T * B
        assert torch.equal(actor.flat_params, model.flat_params)
        outs.append((sorted(logged), model.flat_params.clone()))
    assert outs[0][0] == outs[1][0] and torch.equal(outs[0][1], outs[1][1])


def test_synthetic_actor_pipeline_runs():
    """configs[2] in miniature: 6 actor threads -> pinned slots -> queue -> 2 learner threads."""
    from torchbeast_b200 import actors, monobeast, optim, polybeast_learner, staging
    T, B, A = 5, 4, 6
    dev = torch.device("cuda", 0)
    model = monobeast.AtariNet((4, 84, 84), A, True)
    actor = monobeast.AtariNet((4, 84, 84), A, True)
    opt = optim.RMSprop(model, lr=0.00048, momentum=0, eps=0.01, alpha=0.99)
    stager = staging.RolloutStager(staging.spec_for(T, B, A, use_last_action=False), dev, depth=3)
    model._tb_stager = stager
    q = actors.LearnerQueue()
    pool = actors.SyntheticActors(stager, 6, T, B, A, q, state_shape=(2, 512 + A + 1))
    flags = types.SimpleNamespace(reward_clipping="abs_one", discounting=0.99, baseline_cost=0.5, entropy_cost=0.0006,
                                  grad_norm_clipping=40.0, unroll_length=T, batch_size=B)
    stats, lock, n = {}, threading.Lock(), [0]

    class Log:
        def log(self, s):
            n[0] += 1
            assert np.isfinite(s["total_loss"])
            if n[0] == 8:
                pool.stop()
    pool.start()
    ths = [threading.Thread(target=polybeast_learner.learn, args=(flags, q, model, actor, opt, None, stats, Log(), lock)) for _ in range(2)]
    for t in ths:
        t.start()
    for t in ths:
        t.join(timeout=120)
        assert not t.is_alive()
    assert n[0] >= 8 and stats["step"] == n[0] * T * B and pool.rollouts >= 8 * B
