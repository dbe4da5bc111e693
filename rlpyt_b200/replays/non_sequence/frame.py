"""Frame-buffer replay classes (mirror of ``rlpyt/replays/non_sequence/frame.py:11-42``)."""
from rlpyt_b200.replays.frame import FrameBufferMixin
from rlpyt_b200.replays.non_sequence.n_step import NStepReturnBuffer
from rlpyt_b200.replays.non_sequence.prioritized import PrioritizedReplay
from rlpyt_b200.replays.non_sequence.uniform import UniformReplay


class NStepFrameBuffer(FrameBufferMixin, NStepReturnBuffer):
    """Observations are re-assembled from single frames, oldest to newest; frames that belong to
    the previous episode (``done`` within the last n_frames-1 steps) are zeroed
    (non_sequence/frame.py:14-30) - inside the fused extraction kernel."""

    def _observation_store(self):
        return self.samples_frames, self.n_frames

    def _stacked(self):
        return True


class UniformReplayFrameBuffer(UniformReplay, NStepFrameBuffer):
    pass


class PrioritizedReplayFrameBuffer(PrioritizedReplay, NStepFrameBuffer):
    pass


# Asynchronous-runner variants (rlpyt/replays/non_sequence/frame.py:33-42)
from rlpyt_b200.replays.async_ import AsyncReplayBufferMixin  # noqa: E402


class AsyncUniformReplayFrameBuffer(AsyncReplayBufferMixin, UniformReplayFrameBuffer):
    pass


class AsyncPrioritizedReplayFrameBuffer(AsyncReplayBufferMixin, PrioritizedReplayFrameBuffer):
    pass
