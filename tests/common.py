"""Shared seeded input generators for the parity tests (same formulas as
oracle/make_golden.py, which produced tests/golden/*.npz from the reference)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def arange_inputs(T, B):
    """Reference test formulas, /root/reference/tests/vtrace_test.py:143-157."""
    ar = np.arange(T * B, dtype=np.float32).reshape(T, B)
    return dict(
        log_rhos=(5 * (ar / (B * T) - 0.5)).astype(np.float32),
        discounts=np.array([[0.9 / (b + 1) for b in range(B)] for _ in range(T)], dtype=np.float32),
        rewards=ar.copy(),
        values=(ar / B).astype(np.float32),
        bootstrap_value=(np.arange(B, dtype=np.float32) + 1.0),
    )


def random_vtrace_inputs(T, B, A, seed):
    rs = np.random.RandomState(seed)
    return dict(
        behavior_policy_logits=rs.randn(T, B, A).astype(np.float32),
        target_policy_logits=rs.randn(T, B, A).astype(np.float32),
        actions=rs.randint(0, A, size=(T, B)).astype(np.int64),
        discounts=(0.99 * (rs.rand(T, B) > 0.05)).astype(np.float32),
        rewards=np.clip(rs.randn(T, B), -1, 1).astype(np.float32),
        values=rs.randn(T, B).astype(np.float32),
        bootstrap_value=rs.randn(B).astype(np.float32),
    )


RANDOM_CASES = (((80, 32, 6), 1), ((20, 8, 3), 2), ((7, 2, 18), 3), ((600, 16, 6), 4))
CLIPS = ((1.0, 1.0), (None, None), (3.7, 2.2))


GRAD_SAMPLES = 4096


def sample_index(numel, k=GRAD_SAMPLES):
    """Strided sample positions of the large fixtures' grad_sample/ and param_sample/ arrays (same formula as
    oracle/make_golden.py:sample_index)."""
    if numel <= k:
        return np.arange(numel)
    return (np.arange(k, dtype=np.int64) * numel) // k
