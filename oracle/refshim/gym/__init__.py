"""Minimal stand-in for `gym` so the reference's monobeast module imports.

TEST INFRASTRUCTURE ONLY (used by oracle/make_golden.py in the build container).
Only the class names that atari_wrappers.py subclasses at import time exist.
"""
from . import spaces  # noqa: F401


class Env:
    pass


class Wrapper(Env):
    def __init__(self, env=None):
        self.env = env


class RewardWrapper(Wrapper):
    pass


class ObservationWrapper(Wrapper):
    pass


class ActionWrapper(Wrapper):
    pass


def make(*a, **k):
    raise RuntimeError("gym stub: no environments available")
