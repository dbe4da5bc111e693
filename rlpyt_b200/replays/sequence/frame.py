"""Frame-wise sequence replay classes (mirror of ``rlpyt/replays/sequence/frame.py:10-70``): observations are
re-assembled from single frames as [T, B, C, H, W], oldest to newest along C, frames of a previous episode zeroed -
inside the extraction kernel (csrc/replay.cu ``replay_extract_seq_frames_kernel``)."""
from rlpyt_b200.replays.frame import FrameBufferMixin
from rlpyt_b200.replays.sequence.n_step import SequenceNStepReturnBuffer
from rlpyt_b200.replays.sequence.prioritized import PrioritizedSequenceReplay
from rlpyt_b200.replays.sequence.uniform import UniformSequenceReplay


class SequenceNStepFrameBuffer(FrameBufferMixin, SequenceNStepReturnBuffer):

    def _observation_store(self):
        return self.samples_frames, self.n_frames

    def _obs_out_shape(self, L, n, store, nf):
        return (L, n, nf) + tuple(store.shape[2:])


class UniformSequenceReplayFrameBuffer(UniformSequenceReplay, SequenceNStepFrameBuffer):
    pass


class PrioritizedSequenceReplayFrameBuffer(PrioritizedSequenceReplay, SequenceNStepFrameBuffer):
    pass


# Asynchronous-runner variants (rlpyt/replays/sequence/frame.py:63-70)
from rlpyt_b200.replays.async_ import AsyncReplayBufferMixin  # noqa: E402


class AsyncUniformSequenceReplayFrameBuffer(AsyncReplayBufferMixin, UniformSequenceReplayFrameBuffer):
    pass


class AsyncPrioritizedSequenceReplayFrameBuffer(AsyncReplayBufferMixin, PrioritizedSequenceReplayFrameBuffer):
    pass
