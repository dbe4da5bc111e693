"""Uniformly sampled sequence replay (mirror of ``rlpyt/replays/sequence/uniform.py:7-45``)."""
import numpy as np

from rlpyt_b200.replays.sequence.n_step import SequenceNStepReturnBuffer


class UniformSequenceReplay:

    def set_batch_T(self, batch_T):
        self.batch_T = batch_T

    def sample_batch(self, batch_B, batch_T=None):
        batch_T = self.batch_T if batch_T is None else batch_T
        T_idxs, B_idxs = self.sample_idxs(batch_B, batch_T)
        return self.extract_batch(T_idxs, B_idxs, batch_T)

    def sample_idxs(self, batch_B, batch_T):
        """uniform.py:22-40: the same two ``np.random.randint`` draws as the reference, so a seeded host generator
        gives the reference's index stream; the invalid band around the cursor widens by the sequence length, and
        starts are rounded down to steps with a stored RNN state."""
        t, b, f = self.t, self.off_backward + batch_T, self.off_forward
        high = self.T - b - f if self._buffer_full else t - b - f
        T_idxs = np.random.randint(low=0, high=high, size=(batch_B,))
        T_idxs[T_idxs >= t - b] += min(t, b) + f
        if self.rnn_state_interval > 0:
            T_idxs = (T_idxs // self.rnn_state_interval) * self.rnn_state_interval
        B_idxs = np.random.randint(low=0, high=self.B, size=(batch_B,))
        return T_idxs, B_idxs


class UniformSequenceReplayBuffer(UniformSequenceReplay, SequenceNStepReturnBuffer):
    pass
