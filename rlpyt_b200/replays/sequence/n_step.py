"""Sequence replay with n-step returns and stored recurrent state, in HBM (mirror of
``rlpyt/replays/sequence/n_step.py:19-105`` ``SequenceNStepReturnBuffer``).  ``extract_batch`` is two kernel
launches (csrc/replay.cu: frame-stack sequences with done-blanking; every scalar sequence) plus a small gather
of the stored RNN state."""
import math

import numpy as np
import torch

from rlpyt_b200 import _lib
from rlpyt_b200.replays.n_step import BaseNStepReturnBuffer, _dev
from rlpyt_b200.replays.non_sequence.n_step import _idx
from rlpyt_b200.utils.buffer import buffer_from_example
from rlpyt_b200.utils.collections import namedarraytuple

SamplesFromReplay = namedarraytuple("SamplesFromReplay", ["all_observation", "all_action", "all_reward", "return_", "done",
                                                           "done_n", "init_rnn_state"])
SamplesToBuffer = None


def _map_leaves(buf, fn):
    if buf is None:
        return None
    if isinstance(buf, torch.Tensor):
        return fn(buf)
    return type(buf)(*(_map_leaves(b, fn) for b in buf))


class SequenceNStepReturnBuffer(BaseNStepReturnBuffer):

    def __init__(self, example, size, B, rnn_state_interval, batch_T=None, **kwargs):
        """sequence/n_step.py:28-48.  ``rnn_state_interval``: 0 = no RNN state stored, 1 = every step, k > 1 = every
        k-th step (T is rounded up to a multiple of k; sequences may only start there)."""
        self.rnn_state_interval = rsi = rnn_state_interval
        self.batch_T = batch_T
        if rsi <= 1:
            buffer_example = example
        else:
            global SamplesToBuffer
            names = [f for f in example._fields if f != "prev_rnn_state"]
            SamplesToBuffer = namedarraytuple("SamplesToBuffer", names)
            buffer_example = SamplesToBuffer(*(v for k, v in example.items() if k != "prev_rnn_state"))
            size = B * rsi * math.ceil(math.ceil(size / B) / rsi)
        super().__init__(example=buffer_example, size=size, B=B, **kwargs)
        if rsi > 1:
            assert self.T % rsi == 0
            self.rnn_T = self.T // rsi
            self.samples_prev_rnn_state = buffer_from_example(example.prev_rnn_state, (self.rnn_T, B), where="cuda",
                                                              device=self.device)

    def append_samples(self, samples):
        """sequence/n_step.py:50-66: rows ``offset::rsi`` of the incoming batch are the steps whose state is kept."""
        t, rsi = self.t, self.rnn_state_interval
        if rsi <= 1:
            return super().append_samples(samples)
        buffer_samples = SamplesToBuffer(*(v for k, v in samples.items() if k != "prev_rnn_state"))
        T, idxs = super().append_samples(buffer_samples)
        start, stop = math.ceil(t / rsi), ((t + T - 1) // rsi) + 1
        offset = (rsi - t) % rsi
        rows = torch.as_tensor(np.arange(start, stop) % self.rnn_T, device=self.device) if stop > self.rnn_T else slice(start, stop)
        for dst, src in zip(_leaves(self.samples_prev_rnn_state), _leaves(samples.prev_rnn_state)):
            dst[rows] = _dev(src, self.device)[offset::rsi].to(dst.dtype)
        return T, idxs

    def _observation_store(self):
        """(storage [rows, B, ...], n_frames): whole observations count as 1 'frame' per step."""
        return self.samples.observation, 1

    def _obs_out_shape(self, L, n, store, nf):
        return (L, n) + tuple(store.shape[2:])

    def extract_batch(self, T_idxs, B_idxs, T):
        """sequence/n_step.py:68-101 -> SamplesFromReplay of CUDA tensors with leading dims [T (+ n_step), len(B_idxs)]."""
        s, dev, rsi = self.samples, self.device, self.rnn_state_interval
        T_host = np.asarray(T_idxs.cpu() if isinstance(T_idxs, torch.Tensor) else T_idxs) if rsi > 1 else None
        T_idxs, B_idxs = _idx(T_idxs, dev), _idx(B_idxs, dev)
        n = T_idxs.numel()
        if rsi > 1:
            assert np.all(T_host % rsi == 0)
            init = _map_leaves(self.samples_prev_rnn_state, lambda x: x[T_idxs // rsi, B_idxs])
        elif rsi == 1:
            init = _map_leaves(s.prev_rnn_state, lambda x: x[T_idxs, B_idxs])
        else:
            init = None
        store, nf = self._observation_store()
        if s.action.dtype != torch.int64 or s.action.dim() != 2 or s.reward.dtype != torch.float32:
            raise NotImplementedError("fused extraction supports int64 scalar actions and float32 rewards")
        item_shape = tuple(store.shape[2:])
        frame_bytes = max(1, int(np.prod(item_shape, dtype=np.int64))) * store.element_size()
        L = T + self.n_step_return
        obs = torch.empty(self._obs_out_shape(L, n, store, nf), dtype=store.dtype, device=dev)
        act = torch.empty((L, n), dtype=torch.int64, device=dev)
        rew = torch.empty((L, n), dtype=torch.float32, device=dev)
        ret = torch.empty((T, n), dtype=torch.float32, device=dev)
        done, done_n = (torch.empty((T, n), dtype=torch.uint8, device=dev) for _ in range(2))
        with torch.cuda.device(dev):
            _lib.call("rl_replay_extract_sequences", _lib.ptr(store), _lib.ptr(s.action), _lib.ptr(s.reward),
                      _lib.ptr(s.done.view(torch.uint8)), _lib.ptr(self.samples_return_),
                      _lib.ptr(self.samples_done_n.view(torch.uint8)), self.T, self.B, frame_bytes, nf, self.n_step_return,
                      _lib.ptr(T_idxs), _lib.ptr(B_idxs), n, T, _lib.ptr(obs), _lib.ptr(act), _lib.ptr(rew), _lib.ptr(ret),
                      _lib.ptr(done), _lib.ptr(done_n), _lib.stream(), n_launch=2)
        return SamplesFromReplay(all_observation=obs, all_action=act, all_reward=rew, return_=ret,
                                 done=done.view(torch.bool), done_n=done_n.view(torch.bool), init_rnn_state=init)


def _leaves(buf):
    if buf is None:
        return []
    if isinstance(buf, (torch.Tensor, np.ndarray)):
        return [buf]
    out = []
    for b in buf:
        out.extend(_leaves(b))
    return out
