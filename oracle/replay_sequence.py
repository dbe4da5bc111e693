"""Oracle: sequence replay (R2D1) - n-step-return frame buffer that returns [T,B] sequences, with optional
periodic RNN-state storage and uniform / prioritized sampling (numpy).

Test infrastructure only (see oracle/__init__.py).  Restates, on top of oracle/replay.py's ring buffer:
  rlpyt/utils/misc.py:38-56                    extract_sequences (wrap at the end; and the placement it gives a
                                               NEGATIVE start index: the wrapped rows land at the END of the sequence)
  rlpyt/replays/sequence/n_step.py:28-48       storage of every rsi-th prev_rnn_state, T rounded up to a multiple of rsi
  rlpyt/replays/sequence/n_step.py:50-66       append_samples: which rows of the incoming batch carry a stored state
  rlpyt/replays/sequence/n_step.py:68-101      extract_batch (all_observation / all_action / all_reward cover T + n_step
                                               steps, actions and rewards start one step earlier)
  rlpyt/replays/sequence/frame.py:18-50        observation sequences from single frames, blanked after a done
  rlpyt/replays/sequence/uniform.py:22-40      sample_idxs
  rlpyt/replays/sequence/prioritized.py:53-108 tree over stored-state steps, input priorities, is_weights WITHOUT epsilon
Samples are dicts of numpy arrays like oracle/replay.py, plus ``prev_rnn_state``: dict name -> [T,B,...] array.
"""
import math

import numpy as np

from oracle.replay import FrameReplay
from oracle.sum_tree import SumTree


def sequence_rows(t, T, L):
    """Row of a length-L ring that position j = 0..T-1 of a sequence started at ``t`` reads (misc.py:45-54)."""
    j = np.arange(T)
    if t + T > L:                       # wrap at the end
        return (t + j) % L
    if t < 0:                           # "wrap beginning": first T+t rows from the head, last -t rows from the tail
        return np.where(j < T + t, j, L - T + j)
    return t + j


def extract_sequences(arr, T_idxs, B_idxs, T):
    out = np.empty((T, len(B_idxs)) + arr.shape[2:], dtype=arr.dtype)
    for i, (t, b) in enumerate(zip(T_idxs, B_idxs)):
        out[:, i] = arr[sequence_rows(int(t), T, len(arr)), b]
    return out


class SequenceFrameReplay(FrameReplay):

    def __init__(self, obs_shape, size, B, rnn_state_interval, batch_T, rnn_state_shapes=None, discount=1,
                 n_step_return=1, prioritized=True, alpha=0.6, beta=0.4, default_priority=1, unique=False,
                 input_priorities=False, input_priority_shift=0):
        self.rsi = rsi = rnn_state_interval
        self.batch_T = batch_T
        if rsi > 1:                                                            # sequence/n_step.py:37-44
            size = B * rsi * math.ceil(math.ceil(size / B) / rsi)
        super().__init__(obs_shape, size, B, discount=discount, n_step_return=n_step_return, prioritized=False)
        self.rnn_state = None
        if rsi >= 1:
            rows = self.T // rsi if rsi > 1 else self.T
            self.rnn_state = {k: np.zeros((rows, B) + tuple(shp), dtype=np.float32) for k, shp in rnn_state_shapes.items()}
        if rsi > 1:
            assert self.T % rsi == 0
            self.rnn_T = self.T // rsi
        self.seq_prioritized = prioritized         # (the parent's own ``prioritized`` stays False: it must not advance a tree)
        self.alpha, self.beta, self.unique = alpha, beta, unique
        if prioritized:                                                        # sequence/prioritized.py:60-73
            r = max(1, rsi)
            self.tree = SumTree(T=self.T // r, B=B,
                                off_backward=math.ceil((1 + self.off_backward + batch_T) / r),
                                off_forward=math.ceil(self.off_forward / r),
                                default_value=default_priority ** alpha,
                                enable_input_priorities=input_priorities, input_priority_shift=input_priority_shift)

    def append_samples(self, samples, priorities=None):
        t, rsi = self.t, self.rsi
        T, idxs = super().append_samples(samples)        # ring buffer + returns + frames (prioritized=False there)
        if rsi == 1:
            for k, v in samples["prev_rnn_state"].items():
                self.rnn_state[k][idxs] = v
        elif rsi > 1:                                                          # sequence/n_step.py:58-65
            start, stop = math.ceil(t / rsi), ((t + T - 1) // rsi) + 1
            offset = (rsi - t) % rsi
            rows = np.arange(start, stop) % self.rnn_T
            for k, v in samples["prev_rnn_state"].items():
                self.rnn_state[k][rows] = v[offset::rsi]
        if self.seq_prioritized:                                               # sequence/prioritized.py:79-100
            if rsi <= 1:
                self.tree.advance(T, priorities=priorities)
            else:
                if priorities is not None and np.ndim(priorities) == 2:
                    priorities = priorities[(rsi - t) % rsi::rsi]
                n = self.t // rsi - t // rsi
                if self.t < t:
                    n += self.T // rsi
                self.tree.advance(n, priorities=priorities)
        return T, idxs

    # ---- extraction -----------------------------------------------------------------------------
    def extract_observation_sequences(self, T_idxs, B_idxs, T):
        """sequence/frame.py:18-50: position j of sample i is the frame stack of ring time (t+j) % T_buf; channel c
        (0 = oldest) is blanked iff a done lies 1..nf-1-c steps before that time."""
        nf = self.n_frames
        obs = np.empty((T, len(B_idxs), nf) + self.frame_shape, dtype=np.uint8)
        for i, (t, b) in enumerate(zip(T_idxs, B_idxs)):
            times = (int(t) + np.arange(T)) % self.T
            for c in range(nf):
                obs[:, i, c] = self.frames[times + c, b]
            for k in range(1, nf):
                hit = self.done[(times - k) % self.T, b]
                obs[hit, i, :nf - k] = 0
        return obs

    def extract_batch(self, T_idxs, B_idxs, T):
        T_idxs, B_idxs = np.asarray(T_idxs), np.asarray(B_idxs)
        n = self.n_step_return
        batch = dict(
            all_observation=self.extract_observation_sequences(T_idxs, B_idxs, T + n),
            all_action=extract_sequences(self.action, T_idxs - 1, B_idxs, T + n),
            all_reward=extract_sequences(self.reward, T_idxs - 1, B_idxs, T + n),
            return_=extract_sequences(self.return_, T_idxs, B_idxs, T),
            done=extract_sequences(self.done, T_idxs, B_idxs, T),
            done_n=extract_sequences(self.done_n, T_idxs, B_idxs, T),
        )
        if self.rsi > 1:
            assert np.all(T_idxs % self.rsi == 0)
            batch["init_rnn_state"] = {k: v[T_idxs // self.rsi, B_idxs] for k, v in self.rnn_state.items()}
        elif self.rsi == 1:
            batch["init_rnn_state"] = {k: v[T_idxs, B_idxs] for k, v in self.rnn_state.items()}
        return batch

    # ---- sampling -------------------------------------------------------------------------------
    def sample_idxs(self, batch_B, batch_T):
        """sequence/uniform.py:22-40."""
        t, b, f = self.t, self.off_backward + batch_T, self.off_forward
        high = self.T - b - f if self._buffer_full else t - b - f
        T_idxs = np.random.randint(low=0, high=high, size=(batch_B,))
        T_idxs[T_idxs >= t - b] += min(t, b) + f
        if self.rsi > 0:
            T_idxs = (T_idxs // self.rsi) * self.rsi
        B_idxs = np.random.randint(low=0, high=self.B, size=(batch_B,))
        return T_idxs, B_idxs

    def sample_batch(self, batch_B, random_values=None):
        if not self.seq_prioritized:
            T_idxs, B_idxs = self.sample_idxs(batch_B, self.batch_T)
            batch = self.extract_batch(T_idxs, B_idxs, self.batch_T)
            batch["T_idxs"], batch["B_idxs"] = T_idxs, B_idxs
            return batch
        (T_idxs, B_idxs), pri = self.tree.sample(batch_B, unique=self.unique, random_values=random_values)
        if self.rsi > 1:
            T_idxs = T_idxs * self.rsi
        batch = self.extract_batch(T_idxs, B_idxs, self.batch_T)
        is_w = (1. / pri) ** self.beta                                         # sequence/prioritized.py:112 (no epsilon)
        is_w /= max(is_w)
        batch["is_weights"] = is_w.astype(np.float32)
        batch["T_idxs"], batch["B_idxs"], batch["priorities"] = T_idxs, B_idxs, pri
        return batch
