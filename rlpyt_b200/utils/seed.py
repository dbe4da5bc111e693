"""Seeding helpers (mirror of ``rlpyt/utils/seed.py:10-65``)."""
import time

import numpy as np


def set_seed(seed):
    """Seed python, numpy, torch and (if initialised) cuda (seed.py:10-22)."""
    import random
    import torch
    seed %= 4294967294
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available() and torch.cuda.is_initialized():
        torch.cuda.manual_seed(seed)


def make_seed():
    """Semi-random seed from the clock (seed.py:29-36)."""
    d = 10000
    t = time.time()
    sub1 = int(t * d) % d
    sub2 = int(t * d ** 2) % d
    s = 1e-3
    s_inv = 1. / s
    time.sleep(s * sub2 / d)
    t2 = time.time()
    t2 = t2 - int(t2)
    t2 = int(t2 * d * s_inv) % d
    time.sleep(s * sub1 / d)
    t3 = time.time()
    t3 = t3 - int(t3)
    t3 = int(t3 * d * s_inv * 10) % 10
    return (sub1 * d + t2) * 10 + t3


def set_envs_seeds(envs, seed):
    """Environment ``i`` gets ``seed + i`` (seed.py:54-65)."""
    if seed is not None:
        for i, env in enumerate(envs):
            env.seed(seed + i)
