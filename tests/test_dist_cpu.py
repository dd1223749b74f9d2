"""CPU, world_size 2, gloo: the data-parallel host logic (column sharding + one SUM all-reduce of
the flat gradient).  Each rank computes the oracle's gradient of its batch-column
shard; after learner._all_reduce_grads the flat gradient must equal the full-batch gradient."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _flat_grads(o, names):
    return torch.cat([o["grads"][n].flatten() for n in names])


def _worker(rank, world, port, use_lstm, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        from oracle import learner_torch as LT
        from torchbeast_b200 import learner
        T, B, A = 5, 4, 6
        batch = LT.synthetic_batch(T, B, A, seed=21)
        shapes = LT.atarinet_param_shapes(A, use_lstm)
        p = LT.random_params(shapes, seed=5, dtype=torch.float64)
        state = ()
        if use_lstm:
            rs = np.random.RandomState(3)
            state = tuple(torch.from_numpy(rs.randn(2, B, 512 + A + 1) * 0.1) for _ in range(2))
        names = list(shapes)
        shard, sstate = learner.shard_rollout(batch, state, rank, world)
        assert shard["frame"].shape[1] == B // world and shard["frame"].is_contiguous()
        o = LT.learner_step(p, shard, sstate, net="atari", update=False)
        flat = _flat_grads(o, names)
        losses = torch.stack([o["pg_loss"], o["baseline_loss"], o["entropy_loss"]])
        assert learner._all_reduce_grads(flat) is True
        learner._all_reduce_grads(losses)
        # the two-bucket overlapped path (LSTM + heads slice reduced between the backward phases, trunk slice after):
        # same result as the single all-reduce, every element reduced exactly once
        local = _flat_grads(o, names)
        split = local.numel() // 3

        class TwoPhase:
            flat_params = local
            def grad_split(self):
                return split
            def learner_backward(self, gl, gv, between=None):
                fg = torch.zeros_like(local)
                fg[split:] = local[split:]          # phase 1: heads + LSTM slice final
                between(fg, split)
                fg[:split] = local[:split]          # phase 2: trunk slice
                return fg
        two = learner._backward_with_overlapped_all_reduce(TwoPhase(), None, None)
        assert torch.equal(two, flat), float((two - flat).abs().max())
        if rank == 0:
            full = LT.learner_step(p, batch, state, net="atari", update=False)
            ref = _flat_grads(full, names)
            ret["grad_err"] = float((flat - ref).abs().max() / ref.abs().max())
            ret["loss_err"] = float((losses.sum() - full["total_loss"]).abs() / full["total_loss"].abs())
    finally:
        dist.destroy_process_group()


def _run(use_lstm):
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, use_lstm, ret), nprocs=2, join=True)
    assert ret["grad_err"] < 1e-10, ret["grad_err"]
    assert ret["loss_err"] < 1e-10, ret["loss_err"]


def test_sharded_gradient_sum_equals_full_batch():
    _run(False)


def test_sharded_gradient_sum_equals_full_batch_lstm():
    _run(True)


def test_shard_rollout_rejects_indivisible_batch():
    import pytest
    from torchbeast_b200 import learner
    with pytest.raises(ValueError):
        learner.shard_rollout(dict(x=torch.zeros(3, 5)), (), 0, 2)
