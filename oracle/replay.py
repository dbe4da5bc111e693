"""Oracle: n-step-return frame replay buffer with prioritized / uniform sampling (numpy).

Test infrastructure only (see oracle/__init__.py).  One class restates the behaviour of the
reference's ``PrioritizedReplayFrameBuffer`` / ``UniformReplayFrameBuffer`` stack:
  rlpyt/replays/n_step.py:41-108          ring buffer, cursor, append + compute_returns
  rlpyt/replays/frame.py:27-59            frame-wise storage (only the newest frame per step)
  rlpyt/replays/non_sequence/n_step.py:16-43   extract_batch
  rlpyt/replays/non_sequence/frame.py:14-30    extract_observation (+ blanking after done)
  rlpyt/replays/non_sequence/prioritized.py:24-79   tree coupling, is_weights, priority update
  rlpyt/replays/non_sequence/uniform.py:17-28  sample_idxs
Samples are plain dicts of numpy arrays: observation [T,B,C,H,W] u8 (C = n_frames, oldest->newest),
action [T,B] i64, reward [T,B] f32, done [T,B] bool.
"""
import math

import numpy as np

from oracle.returns import discount_return_n_step
from oracle.sum_tree import SumTree

EPS = 1e-6  # rlpyt/replays/non_sequence/prioritized.py:8


class FrameReplay:

    def __init__(self, obs_shape, size, B, discount=1, n_step_return=1, prioritized=True, alpha=0.6,
                 beta=0.4, default_priority=1, unique=False, action_dtype=np.int64):
        self.n_frames = nf = obs_shape[0]
        self.frame_shape = tuple(obs_shape[1:])
        self.T = T = math.ceil(size / B)                                      # n_step.py:42
        self.B, self.size = B, T * B
        self.discount, self.n_step_return = discount, n_step_return
        self.t = 0
        self.action = np.zeros((T, B), dtype=action_dtype)
        self.reward = np.zeros((T, B), dtype=np.float32)
        self.done = np.zeros((T, B), dtype=bool)
        if n_step_return > 1:                                                 # n_step.py:50-57
            self.return_ = np.zeros((T, B), dtype=np.float32)
            self.done_n = np.zeros((T, B), dtype=bool)
        else:
            self.return_, self.done_n = self.reward, self.done
        self._buffer_full = False
        self.off_backward = n_step_return                                     # n_step.py:59
        self.off_forward = max(1, nf - 1)                                     # n_step.py:60, frame.py:44
        self.frames = np.zeros((T + nf - 1, B) + self.frame_shape, dtype=np.uint8)   # frame.py:39-41
        self.new_frames = self.frames[nf - 1:]                                # frame.py:43
        self.prioritized = prioritized
        self.alpha, self.beta, self.unique = alpha, beta, unique
        if prioritized:                                                       # prioritized.py:30-41
            self.tree = SumTree(T, B, self.off_backward, self.off_forward,
                                default_value=default_priority ** alpha)

    # ---- append ---------------------------------------------------------------------------------
    def append_samples(self, samples):
        obs, action, reward, done = (samples[k] for k in ("observation", "action", "reward", "done"))
        T, B = reward.shape[:2]
        assert B == self.B
        t, fm1 = self.t, self.n_frames - 1
        idxs = np.arange(t, t + T) % self.T if t + T > self.T else slice(t, t + T)   # n_step.py:70-73
        self.action[idxs] = action
        self.reward[idxs] = reward
        self.done[idxs] = done
        self._compute_returns(T)                                              # n_step.py:75
        if not self._buffer_full and t + T >= self.T:
            self._buffer_full = True
        self.t = (t + T) % self.T
        self.new_frames[idxs] = obs[:, :, -1]                                 # frame.py:53
        if t == 0:                                                            # frame.py:54-56
            for f in range(fm1):
                self.frames[f] = obs[0, :, f]
        elif self.t < t and fm1 > 0:                                          # frame.py:57-58
            self.frames[:fm1] = self.frames[-fm1:]
        if self.prioritized:
            self.tree.advance(T)                                              # prioritized.py:57
        return T, idxs

    def _compute_returns(self, T):
        """n_step.py:81-108 (in place without wrap; copies - and a deliberate wrong wrap at the very
        first call - otherwise)."""
        if self.n_step_return == 1:
            return
        t, nm1 = self.t, self.n_step_return - 1
        if t - nm1 >= 0 and t + T <= self.T:
            ret, dn = discount_return_n_step(self.reward[t - nm1:t + T], self.done[t - nm1:t + T],
                                             self.n_step_return, self.discount)
            self.return_[t - nm1:t - nm1 + T] = ret
            self.done_n[t - nm1:t - nm1 + T] = dn
        else:
            idxs = np.arange(t - nm1, t + T) % self.T
            ret, dn = discount_return_n_step(self.reward[idxs], self.done[idxs], self.n_step_return,
                                             self.discount)
            self.return_[idxs[:-nm1]] = ret
            self.done_n[idxs[:-nm1]] = dn

    # ---- extraction -----------------------------------------------------------------------------
    def extract_observation(self, T_idxs, B_idxs):
        """non_sequence/frame.py:14-30."""
        nf = self.n_frames
        obs = np.stack([self.frames[t:t + nf, b] for t, b in zip(T_idxs, B_idxs)], axis=0)
        for f in range(1, nf):
            blank = np.where(self.done[T_idxs - f, B_idxs])[0]                # negative index wraps
            obs[blank, :nf - f] = 0
        return obs

    def extract_batch(self, T_idxs, B_idxs):
        """non_sequence/n_step.py:16-43."""
        T_idxs, B_idxs = np.asarray(T_idxs), np.asarray(B_idxs)
        tgt = (T_idxs + self.n_step_return) % self.T
        batch = dict(
            observation=self.extract_observation(T_idxs, B_idxs),
            prev_action=self.action[T_idxs - 1, B_idxs],
            prev_reward=self.reward[T_idxs - 1, B_idxs],
            action=self.action[T_idxs, B_idxs],
            return_=self.return_[T_idxs, B_idxs],
            done=self.done[T_idxs, B_idxs],
            done_n=self.done_n[T_idxs, B_idxs],
            target_observation=self.extract_observation(tgt, B_idxs),
            target_prev_action=self.action[tgt - 1, B_idxs],
            target_prev_reward=self.reward[tgt - 1, B_idxs],
        )
        t_news = np.where(self.done[T_idxs - 1, B_idxs])[0]                   # n_step.py:40-42
        batch["prev_action"][t_news] = 0
        batch["prev_reward"][t_news] = 0
        return batch

    # ---- sampling -------------------------------------------------------------------------------
    def sample_idxs(self, batch_B):
        """uniform.py:17-28 (``np.random.randint`` twice)."""
        t, b, f = self.t, self.off_backward, self.off_forward
        high = self.T - b - f if self._buffer_full else t - b
        low = 0 if self._buffer_full else f
        T_idxs = np.random.randint(low=low, high=high, size=(batch_B,))
        T_idxs[T_idxs >= t - b] += min(t, b) + f
        B_idxs = np.random.randint(low=0, high=self.B, size=(batch_B,))
        return T_idxs, B_idxs

    def sample_batch(self, batch_B, random_values=None):
        if not self.prioritized:
            T_idxs, B_idxs = self.sample_idxs(batch_B)
            batch = self.extract_batch(T_idxs, B_idxs)
            batch["T_idxs"], batch["B_idxs"] = T_idxs, B_idxs
            return batch
        (T_idxs, B_idxs), pri = self.tree.sample(batch_B, unique=self.unique, random_values=random_values)
        batch = self.extract_batch(T_idxs, B_idxs)                            # prioritized.py:60-71
        is_w = (1. / (pri + EPS)) ** self.beta
        is_w /= max(is_w)
        batch["is_weights"] = is_w.astype(np.float32)
        batch["T_idxs"], batch["B_idxs"], batch["priorities"] = T_idxs, B_idxs, pri
        return batch

    def update_batch_priorities(self, priorities):
        """prioritized.py:73-79: f32 priorities ** alpha (numpy f32 pow), upcast inside the tree."""
        self.tree.update_batch_priorities(np.asarray(priorities) ** self.alpha)
