"""Frame-wise observation storage (mirror of ``rlpyt/replays/frame.py:10-59`` ``FrameBufferMixin``):
only the newest frame of each multi-frame observation is stored - ``[T+n_frames-1, B, H, W]`` u8 in
HBM (7.06 GB for 1 M 84x84 frames, a fraction of the B200's 180 GB)."""
import torch

from rlpyt_b200.replays.n_step import _dev
from rlpyt_b200.utils.buffer import get_leading_dims
from rlpyt_b200.utils.collections import namedarraytuple

BufferSamples = None


class FrameBufferMixin:

    def __init__(self, example, **kwargs):
        field_names = [f for f in example._fields if f != "observation"]
        global BufferSamples
        BufferSamples = namedarraytuple("BufferSamples", field_names)
        buffer_example = BufferSamples(*(v for k, v in example.items() if k != "observation"))
        super().__init__(example=buffer_example, **kwargs)
        self.n_frames = n_frames = get_leading_dims(example.observation, n_dim=1)[0]
        frame = torch.as_tensor(example.observation[0])
        self.samples_frames = torch.zeros((self.T + n_frames - 1, self.B) + tuple(frame.shape),
                                          dtype=frame.dtype, device=self.device)         # frame.py:39-41
        self.samples_new_frames = self.samples_frames[n_frames - 1:]                        # frame.py:43
        self.off_forward = max(self.off_forward, n_frames - 1)

    def append_samples(self, samples):
        """frame.py:46-59: store the newest frame of every observation; on the first append also the
        history frames; on wrap duplicate the tail frames to the head."""
        t, fm1 = self.t, self.n_frames - 1
        buffer_samples = BufferSamples(*(v for k, v in samples.items() if k != "observation"))
        T, idxs = super().append_samples(buffer_samples)
        obs = _dev(samples.observation, self.device)
        self.samples_new_frames[idxs] = obs[:, :, -1]
        if t == 0:
            for f in range(fm1):
                self.samples_frames[f] = obs[0, :, f]
        elif self.t < t and fm1 > 0:
            self.samples_frames[:fm1] = self.samples_frames[-fm1:]
        return T, idxs
