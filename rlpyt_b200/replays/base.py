"""Replay-buffer interface (mirror of ``rlpyt/replays/base.py:3-13``)."""


class BaseReplayBuffer:

    async_ = False

    def append_samples(self, samples):
        raise NotImplementedError

    def sample_batch(self, batch_B, batch_T=None):
        raise NotImplementedError
