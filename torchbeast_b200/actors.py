"""Synthetic actor pool + learner queue: BASELINE.json configs[2] ("polybeast_learner 48 synthetic actors").

Stands in for the reference's ActorPool loops + BatchingQueue (/root/reference/src/cc/actorpool.cc:342-470 actor loop,
:147-186 dequeue_many, :493-506 the rollout nest) WITHOUT environments, gRPC or inference: `num_actors` host threads each
produce `[T+1, ...]` rollouts of synthetic 84x84x4 uint8 frames and write them straight into one batch column of a
PINNED `[T+1, B, ...]` slot of a RolloutStager (N1: no per-rollout tensors, no torch.cat, no pageable copy); a slot
whose B columns are complete is handed to the learner threads through `LearnerQueue` as the reference's nest
`((env_outputs, actor_outputs), initial_agent_state)` whose leaves ARE the slot's tensors, so
polybeast_learner.learn() submits it with one async H2D copy.  Row 0 of every rollout repeats the last row of that
actor's previous rollout (actorpool.cc:443).  Out of scope: real environments / inference."""
import collections
import queue
import threading

import numpy as np
import torch


class LearnerQueue:
    """Iterator protocol of the reference's BatchingQueue as polybeast_learner.learn uses it (pl:306, 380)."""

    def __init__(self, maxsize=0):
        self._q = queue.Queue(maxsize)
        self._closed = False

    def put(self, item):
        self._q.put(item)

    def close(self):
        self._closed = True
        self._q.put(None)

    def size(self):
        return self._q.qsize()

    def __iter__(self):
        return self

    def __next__(self):
        item = self._q.get()
        if item is None:
            self._q.put(None)  # wake the other learner threads too
            raise StopIteration
        return item


class SyntheticActors:
    def __init__(self, stager, num_actors, T, B, num_actions, learner_queue, state_shape=None, seed=0, frame_pool=8):
        self.stager, self.n, self.T, self.B, self.A = stager, num_actors, T, B, num_actions
        self.queue = learner_queue
        self.state_shape = state_shape  # (layers, H) or None
        self._lock = threading.Lock()        # column counters / statistics
        self._claim_lock = threading.Lock()  # column hand-out (may block waiting for a free slot: never held with _lock)
        self._slot = None
        self._next_col = 0
        self._remaining = {}
        self._stop = False
        self._threads = []
        self.rollouts = 0
        rs = np.random.RandomState(seed)
        # a small pool of pre-rendered frame stacks per actor: the host work per rollout is the copy into the pinned column,
        # as it is for a real actor handing over frames it received from its environment
        self._frames = torch.from_numpy(rs.randint(0, 256, size=(frame_pool, T + 1, 4, 84, 84), dtype=np.uint8))
        self._seed = seed

    def _claim(self):
        """(slot index, column) for the next rollout; blocks while every slot is in flight."""
        with self._claim_lock:
            if self._slot is None:
                i = self.stager.acquire_host(timeout=0.5)
                with self._lock:
                    self._remaining[i] = self.B
                self._slot = i
                self._next_col = 0
            i, b = self._slot, self._next_col
            self._next_col += 1
            if self._next_col == self.B:
                self._slot = None
            return i, b

    def _done(self, i):
        with self._lock:
            self._remaining[i] -= 1
            full = self._remaining[i] == 0
            if full:
                del self._remaining[i]
        if full:
            h = self.stager.host[i]
            env = (h["frame"], h["reward"], h["done"], h["episode_step"], h["episode_return"])
            agent = (h["action"], h["policy_logits"], h["baseline"])
            state = ()
            if self.state_shape is not None:
                layers, H = self.state_shape
                state = tuple(torch.zeros(layers, self.B, H) for _ in range(2))
            self.queue.put(((env, agent), state))

    def _make_pool(self, k, pool=4):
        """A few pre-generated rollouts per actor (what its environment steps would have produced): the per-rollout host work
        that remains is the hand-over into the pinned column, as for a real actor."""
        rs = np.random.RandomState(self._seed * 1000 + k)
        T1, A = self.T + 1, self.A
        out = []
        for j in range(pool):
            out.append(dict(
                frame=self._frames[(k + j) % self._frames.shape[0]],
                reward=torch.from_numpy(rs.randn(T1).astype(np.float32)),
                done=torch.from_numpy(rs.rand(T1) < 0.01),
                episode_return=torch.from_numpy(rs.randn(T1).astype(np.float32)),
                episode_step=torch.from_numpy(rs.randint(0, 1000, size=T1).astype(np.int32)),
                policy_logits=torch.from_numpy(rs.randn(T1, A).astype(np.float32)),
                baseline=torch.from_numpy(rs.randn(T1).astype(np.float32)),
                action=torch.from_numpy(rs.randint(0, A, size=T1).astype(np.int64))))
        # row 0 of a rollout = last row of this actor's previous rollout (actorpool.cc:443): the pool is handed over cyclically,
        # so that overlap is baked in once (own copies of the shared frame stacks: row 0 differs per actor)
        out = [dict(r, frame=r["frame"].clone()) for r in out]
        for j in range(pool):
            prev = out[(j - 1) % pool]
            for key, v in out[j].items():
                v[0].copy_(prev[key][-1])
        return [self.stager.prepare_rollout(r) for r in out]

    def _actor(self, k):
        pool = self._make_pool(k)
        n = 0
        while not self._stop:
            try:
                i, b = self._claim()
            except TimeoutError:
                continue
            self.stager.write_prepared(i, b, pool[n % len(pool)])   # ONE native, GIL-free call copies all leaves into the column
            n += 1
            with self._lock:
                self.rollouts += 1
            self._done(i)

    def start(self):
        for k in range(self.n):
            t = threading.Thread(target=self._actor, args=(k,), daemon=True)
            t.start()
            self._threads.append(t)

    def stop(self):
        self._stop = True
        self.queue.close()
