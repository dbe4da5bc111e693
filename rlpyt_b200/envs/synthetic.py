"""Synthetic Atari-shaped environment for benchmarks and collector tests:
uint8 observations drawn from a pre-generated frame pool, sparse +-1 rewards, geometric episode
length.  ``atari_py`` / ``gym`` are not installed and there is no network, so this stands in for
``rlpyt/envs/atari/atari_env.py`` with the same interface (EnvInfo(game_score, traj_done)) at a
near-zero and controllable per-step cost (``step_cost_us`` adds a busy-wait to emulate ALE)."""
import time
from collections import namedtuple

import numpy as np

from rlpyt_b200.envs.base import Env, EnvStep, IntBox

EnvInfo = namedtuple("EnvInfo", ["game_score", "traj_done"])


class SyntheticAtariEnv(Env):

    _POOLS = {}

    def __init__(self, image_shape=(4, 84, 84), n_actions=6, p_done=1 / 500., p_reward=0.04,
                 pool_frames=32, seed=0, step_cost_us=0.0):
        self.image_shape = tuple(image_shape)
        self._action_space = IntBox(low=0, high=n_actions)
        self._observation_space = IntBox(low=0, high=256, shape=self.image_shape, dtype="uint8")
        key = (self.image_shape, pool_frames)
        if key not in SyntheticAtariEnv._POOLS:  # one shared pool per process (forked workers share it)
            rng = np.random.default_rng(1234)
            SyntheticAtariEnv._POOLS[key] = rng.integers(0, 256, size=(pool_frames,) + self.image_shape,
                                                         dtype=np.uint8)
        self._pool = SyntheticAtariEnv._POOLS[key]
        self.p_done, self.p_reward = p_done, p_reward
        self.step_cost_us = step_cost_us
        self.seed(seed)

    def seed(self, seed):
        self._rng = np.random.default_rng(seed)
        self._cursor = int(self._rng.integers(0, len(self._pool)))

    def reset(self):
        self._cursor = int(self._rng.integers(0, len(self._pool)))
        return self._pool[self._cursor]

    def step(self, action):
        if self.step_cost_us:
            end = time.perf_counter() + self.step_cost_us * 1e-6
            while time.perf_counter() < end:
                pass
        u = self._rng.random(2)
        self._cursor = (self._cursor + 1 + int(action)) % len(self._pool)
        reward = 0.0
        if u[0] < self.p_reward:
            reward = 1.0 if u[0] < self.p_reward / 2 else -1.0
        done = bool(u[1] < self.p_done)
        return EnvStep(self._pool[self._cursor], np.float32(reward), done, EnvInfo(int(reward), done))

    @property
    def horizon(self):
        return 27000
