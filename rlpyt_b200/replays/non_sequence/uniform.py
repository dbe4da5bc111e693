"""Uniform replay sampling (mirror of ``rlpyt/replays/non_sequence/uniform.py:7-36``)."""
import numpy as np

from rlpyt_b200.replays.non_sequence.n_step import NStepReturnBuffer


class UniformReplay:

    def sample_batch(self, batch_B):
        T_idxs, B_idxs = self.sample_idxs(batch_B)
        return self.extract_batch(T_idxs, B_idxs)

    def sample_idxs(self, batch_B):
        """uniform.py:17-28: two ``np.random.randint`` draws on the host (same global stream as the
        reference), skipping the invalid rows around the cursor."""
        t, b, f = self.t, self.off_backward, self.off_forward
        high = self.T - b - f if self._buffer_full else t - b
        low = 0 if self._buffer_full else f
        T_idxs = np.random.randint(low=low, high=high, size=(batch_B,))
        T_idxs[T_idxs >= t - b] += min(t, b) + f
        B_idxs = np.random.randint(low=0, high=self.B, size=(batch_B,))
        return T_idxs, B_idxs


class UniformReplayBuffer(UniformReplay, NStepReturnBuffer):
    pass
