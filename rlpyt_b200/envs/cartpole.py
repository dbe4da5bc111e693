"""numpy CartPole-v1 (``gym`` is not installed): the classic cart-pole dynamics (Barto, Sutton &
Anderson 1983; same constants as gym's CartPole-v1) behind the rlpyt Env interface, for
BASELINE.json config 1 (SerialSampler + A2C, T=5, B=8)."""
import math
from collections import namedtuple

import numpy as np

from rlpyt_b200.envs.base import Env, EnvStep, IntBox, FloatBox

EnvInfo = namedtuple("EnvInfo", ["traj_done"])


class CartPoleEnv(Env):
    GRAVITY, M_CART, M_POLE, HALF_LEN, FORCE, TAU = 9.8, 1.0, 0.1, 0.5, 10.0, 0.02
    X_LIMIT, THETA_LIMIT, MAX_STEPS = 2.4, 12 * 2 * math.pi / 360, 500

    def __init__(self, seed=0):
        high = np.array([self.X_LIMIT * 2, np.finfo(np.float32).max, self.THETA_LIMIT * 2,
                         np.finfo(np.float32).max], dtype=np.float32)
        self._observation_space = FloatBox(low=-high, high=high)
        self._action_space = IntBox(low=0, high=2)
        self.seed(seed)
        self._state = np.zeros(4)
        self._t = 0

    def seed(self, seed):
        self._rng = np.random.default_rng(seed)

    def reset(self):
        self._state = self._rng.uniform(-0.05, 0.05, size=4)
        self._t = 0
        return self._state.astype(np.float32)

    def step(self, action):
        x, x_dot, th, th_dot = self._state
        force = self.FORCE if int(action) == 1 else -self.FORCE
        total_m = self.M_CART + self.M_POLE
        pm_l = self.M_POLE * self.HALF_LEN
        cos, sin = math.cos(th), math.sin(th)
        temp = (force + pm_l * th_dot ** 2 * sin) / total_m
        th_acc = (self.GRAVITY * sin - cos * temp) / (self.HALF_LEN * (4.0 / 3.0 - self.M_POLE * cos ** 2 / total_m))
        x_acc = temp - pm_l * th_acc * cos / total_m
        self._state = np.array([x + self.TAU * x_dot, x_dot + self.TAU * x_acc,
                                th + self.TAU * th_dot, th_dot + self.TAU * th_acc])
        self._t += 1
        done = bool(abs(self._state[0]) > self.X_LIMIT or abs(self._state[2]) > self.THETA_LIMIT
                    or self._t >= self.MAX_STEPS)
        return EnvStep(self._state.astype(np.float32), np.float32(1.0), done, EnvInfo(done))

    @property
    def horizon(self):
        return self.MAX_STEPS
