"""Batch extraction for non-sequence replay (mirror of ``rlpyt/replays/non_sequence/n_step.py:9-48``).
``extract_batch`` is ONE fused kernel launch (csrc/replay.cu): both frame-stack gathers with
done-blanking and all scalar fields."""
import numpy as np
import torch

from rlpyt_b200 import _lib
from rlpyt_b200.agents.base import AgentInputs
from rlpyt_b200.replays.n_step import BaseNStepReturnBuffer
from rlpyt_b200.utils.collections import namedarraytuple

SamplesFromReplay = namedarraytuple("SamplesFromReplay",
                                    ["agent_inputs", "action", "return_", "done", "done_n", "target_inputs"])


def _idx(x, device):
    if isinstance(x, torch.Tensor) and x.dtype == torch.int64 and x.device == device and x.is_contiguous():
        return x                                             # what the samplers of this package hand over
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x))
    return torch.as_tensor(x).to(device=device, dtype=torch.int64).contiguous()


class NStepReturnBuffer(BaseNStepReturnBuffer):

    def _observation_store(self):
        """(storage [rows, B, ...], n_frames): plain buffers keep whole observations = 1 'frame'."""
        return self.samples.observation, 1

    def extract_batch(self, T_idxs, B_idxs):
        """non_sequence/n_step.py:16-43 -> SamplesFromReplay of CUDA tensors, leading dim
        ``len(T_idxs)``."""
        s, dev = self.samples, self.device
        T_idxs, B_idxs = _idx(T_idxs, dev), _idx(B_idxs, dev)
        n = T_idxs.numel()
        store, nf = self._observation_store()
        if s.action.dtype != torch.int64 or s.action.dim() != 2 or s.reward.dtype != torch.float32:
            raise NotImplementedError("fused extraction supports int64 scalar actions and float32 rewards")
        item_shape = tuple(store.shape[2:])
        frame_bytes = int(np.prod(item_shape, dtype=np.int64)) * store.element_size() if item_shape else store.element_size()
        out_shape = (n, nf) + item_shape if nf > 1 or self._stacked() else (n,) + item_shape
        obs = torch.empty(out_shape, dtype=store.dtype, device=dev)
        tgt = torch.empty(out_shape, dtype=store.dtype, device=dev)
        # the eight scalar fields share ONE allocation (the batch is built ~10 000 times per second: allocator calls count)
        n8 = (n + 7) // 8 * 8
        arena = torch.empty(n8 * (3 * 8 + 3 * 4 + 2), dtype=torch.uint8, device=dev)
        pa, act, tpa = (arena[k * 8 * n8:(k + 1) * 8 * n8].view(torch.int64)[:n] for k in range(3))
        pr, ret, tpr = (arena[24 * n8 + k * 4 * n8:24 * n8 + (k + 1) * 4 * n8].view(torch.float32)[:n] for k in range(3))
        done, done_n = (arena[36 * n8 + k * n8:36 * n8 + k * n8 + n] for k in range(2))
        with torch.cuda.device(dev):
            _lib.call("rl_replay_extract", _lib.ptr(store), _lib.ptr(s.action), _lib.ptr(s.reward),
                      _lib.ptr(s.done.view(torch.uint8)), _lib.ptr(self.samples_return_),
                      _lib.ptr(self.samples_done_n.view(torch.uint8)), self.T, self.B, frame_bytes, nf,
                      self.n_step_return, _lib.ptr(T_idxs), _lib.ptr(B_idxs), n, _lib.ptr(obs), _lib.ptr(tgt),
                      _lib.ptr(pa), _lib.ptr(pr), _lib.ptr(act), _lib.ptr(ret), _lib.ptr(done), _lib.ptr(done_n),
                      _lib.ptr(tpa), _lib.ptr(tpr), _lib.stream())
        return SamplesFromReplay(
            agent_inputs=AgentInputs(observation=obs, prev_action=pa, prev_reward=pr),
            action=act, return_=ret, done=done.view(torch.bool), done_n=done_n.view(torch.bool),
            target_inputs=AgentInputs(observation=tgt, prev_action=tpa, prev_reward=tpr))

    def _stacked(self):
        return False

    def extract_observation(self, T_idxs, B_idxs):
        """Plain ``observation[T_idxs, B_idxs]`` (non_sequence/n_step.py:45-48)."""
        return self.extract_batch(T_idxs, B_idxs).agent_inputs.observation
