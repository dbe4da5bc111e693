"""Environment interface consumed by the collectors (mirror of ``rlpyt/envs/base.py:5-60`` and
the two space types the path needs from ``rlpyt/spaces``).  Environments stay on the CPU
(north_star); nothing here is accelerated."""
from collections import namedtuple

import numpy as np

EnvStep = namedtuple("EnvStep", ["observation", "reward", "done", "env_info"])
EnvInfo = namedtuple("EnvInfo", [])
EnvSpaces = namedtuple("EnvSpaces", ["observation", "action"])


class IntBox:
    """Integer box space; ``n`` exists for scalar (discrete-action) spaces (rlpyt/spaces/int_box.py)."""

    def __init__(self, low, high, shape=None, dtype="int64", null_value=None):
        self.low, self.high = low, high
        self.shape = () if shape is None else tuple(shape)
        self.dtype = np.dtype(dtype)
        self._null_value = low if null_value is None else null_value

    def sample(self):
        return np.random.randint(low=self.low, high=self.high, size=self.shape, dtype=self.dtype)

    def null_value(self):
        null = np.zeros(self.shape, dtype=self.dtype)
        if self._null_value:
            null[...] = self._null_value
        return null

    @property
    def n(self):
        return self.high - self.low

    def __repr__(self):
        return f"IntBox({self.low}-{self.high - 1} shape={self.shape})"


class FloatBox:
    """Float box space (rlpyt/spaces/float_box.py)."""

    def __init__(self, low, high, shape=None, dtype="float32"):
        self.dtype = np.dtype(dtype)
        self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), shape or np.shape(low)).copy()
        self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), shape or np.shape(high)).copy()
        self.shape = self.low.shape

    def sample(self):
        return (np.random.rand(*self.shape) * (self.high - self.low) + self.low).astype(self.dtype)

    def null_value(self):
        return np.zeros(self.shape, dtype=self.dtype)


class Env:
    """``step(action) -> EnvStep`` and ``reset() -> observation`` (rlpyt/envs/base.py:11-60)."""

    def step(self, action):
        raise NotImplementedError

    def reset(self):
        raise NotImplementedError

    @property
    def action_space(self):
        return self._action_space

    @property
    def observation_space(self):
        return self._observation_space

    @property
    def spaces(self):
        return EnvSpaces(observation=self.observation_space, action=self.action_space)

    @property
    def horizon(self):
        raise NotImplementedError

    def seed(self, seed):
        pass

    def close(self):
        pass
