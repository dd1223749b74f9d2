"""ORACLE (test infrastructure). ctypes loader for oracle/_build/liboracle.so (vtrace_c.c)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "liboracle.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-C", _HERE, "_build/liboracle.so"])
        _LIB = ctypes.CDLL(path)
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _clip(c):
    return ctypes.c_float(-1.0 if c is None else float(c))


def vtrace_scan(log_rhos, discounts, rewards, values, bootstrap, clip_rho=1.0, clip_pg_rho=1.0):
    T, B = values.shape
    arrs = [np.ascontiguousarray(x, np.float32) for x in (log_rhos, discounts, rewards, values, bootstrap)]
    vs = np.empty((T, B), np.float32)
    pg = np.empty((T, B), np.float32)
    rc = lib().oracle_vtrace_scan_f32(*[_p(a) for a in arrs], ctypes.c_int64(T), ctypes.c_int64(B),
                                      _clip(clip_rho), _clip(clip_pg_rho), _p(vs), _p(pg))
    assert rc == 0
    return vs, pg


def impala_loss(behavior_logits, target_logits, actions, rewards, done, values, bootstrap,
                discounting=0.99, baseline_cost=0.5, entropy_cost=0.0006, clip_rewards=True,
                clip_rho=1.0, clip_pg_rho=1.0):
    T, B, A = target_logits.shape
    bl = np.ascontiguousarray(behavior_logits, np.float32)
    tl = np.ascontiguousarray(target_logits, np.float32)
    ac = np.ascontiguousarray(actions, np.int64)
    rw = np.ascontiguousarray(rewards, np.float32)
    dn = np.ascontiguousarray(done, np.uint8)
    va = np.ascontiguousarray(values, np.float32)
    bs = np.ascontiguousarray(bootstrap, np.float32)
    vs = np.empty((T, B), np.float32); pg = np.empty((T, B), np.float32)
    losses = np.zeros(3, np.float64)
    gl = np.empty((T, B, A), np.float32); gv = np.empty((T, B), np.float32)
    rc = lib().oracle_impala_loss_f32(
        _p(bl), _p(tl), _p(ac), _p(rw), _p(dn), _p(va), _p(bs),
        ctypes.c_int64(T), ctypes.c_int64(B), ctypes.c_int64(A),
        ctypes.c_float(discounting), ctypes.c_float(baseline_cost), ctypes.c_float(entropy_cost),
        ctypes.c_int(int(clip_rewards)), _clip(clip_rho), _clip(clip_pg_rho),
        _p(vs), _p(pg), _p(losses), _p(gl), _p(gv))
    assert rc == 0
    return dict(vs=vs, pg_advantages=pg, losses=losses, grad_logits=gl, grad_values=gv)
