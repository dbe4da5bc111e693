"""Ring buffer with n-step returns in HBM (mirror of ``rlpyt/replays/n_step.py:11-108``
``BaseNStepReturnBuffer``).  Storage is ``[T,B]`` time-major like the reference; the n-step return
of the newly completed rows is computed by the ``rl_nstep_return_f32`` kernel on append."""
import math

import numpy as np
import torch

from rlpyt_b200 import _lib
from rlpyt_b200.algos.utils import _discount_pow
from rlpyt_b200.replays.base import BaseReplayBuffer
from rlpyt_b200.utils.buffer import buffer_from_example, get_leading_dims


def _dev(x, device):
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(x)
    return x if (isinstance(x, torch.Tensor) and x.device == device) else torch.as_tensor(x).to(device)


def _assign_rows(dst, idxs, src, device):
    """``dst[idxs] = src`` leaf by leaf (nested namedarraytuples: e.g. ``prev_rnn_state`` = (h, c))."""
    if dst is None:
        return
    if isinstance(dst, torch.Tensor):
        dst[idxs] = _dev(src, device).to(dst.dtype)
        return
    for name, d in dst.items():
        _assign_rows(d, idxs, getattr(src, name), device)


class BaseNStepReturnBuffer(BaseReplayBuffer):

    def __init__(self, example, size, B, discount=1, n_step_return=1, device=None):
        self.T = T = math.ceil(size / B)
        self.B = B
        self.size = T * B
        self.discount = discount
        self.n_step_return = n_step_return
        self.t = 0
        if device is None:
            if not torch.cuda.is_available():
                raise _lib.B200LibraryError("rlpyt_b200 replay buffers live in HBM: a CUDA device is required")
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)
        self.samples = buffer_from_example(example, (T, B), where="cuda", device=self.device)
        if n_step_return > 1:                                                   # n_step.py:50-57
            self.samples_return_ = torch.zeros((T, B), dtype=torch.float32, device=self.device)
            self.samples_done_n = torch.zeros((T, B), dtype=torch.bool, device=self.device)
        else:
            self.samples_return_ = self.samples.reward
            self.samples_done_n = self.samples.done
        self._buffer_full = False
        self.off_backward = n_step_return
        self.off_forward = 1

    def append_samples(self, samples):
        """n_step.py:62-79.  ``samples`` leaves may be CUDA / CPU tensors or numpy arrays."""
        T, B = get_leading_dims(samples, n_dim=2)
        assert B == self.B
        t = self.t
        if t + T > self.T:
            idxs = torch.as_tensor(np.arange(t, t + T) % self.T, device=self.device)
        else:
            idxs = slice(t, t + T)
        _assign_rows(self.samples, idxs, samples, self.device)
        self.compute_returns(T)
        if not self._buffer_full and t + T >= self.T:
            self._buffer_full = True
        self.t = (t + T) % self.T
        return T, idxs

    def compute_returns(self, T):
        """n_step.py:81-108 on the device (in place when the rows do not wrap)."""
        if self.n_step_return == 1:
            return
        t, s, n = self.t, self.samples, self.n_step_return
        nm1 = n - 1
        gpow = _discount_pow(self.discount, n, self.device)
        done_u8 = s.done.view(torch.uint8)
        if t - nm1 >= 0 and t + T <= self.T:
            rew, dn = s.reward[t - nm1:t + T], done_u8[t - nm1:t + T]
            ret_dst = self.samples_return_[t - nm1:t - nm1 + T]
            dn_dst = self.samples_done_n.view(torch.uint8)[t - nm1:t - nm1 + T]
            with torch.cuda.device(self.device):
                _lib.call("rl_nstep_return_f32", _lib.ptr(rew), _lib.ptr(dn), _lib.ptr(gpow), _lib.ptr(ret_dst),
                          _lib.ptr(dn_dst), T + nm1, self.B, n, 0, _lib.stream())
        else:  # wrap: gather rows, compute, scatter (the "wrong wrap at first call" is kept, :100)
            rows = torch.as_tensor(np.arange(t - nm1, t + T) % self.T, device=self.device)
            rew, dn = s.reward[rows].contiguous(), done_u8[rows].contiguous()
            ret = torch.empty((T, self.B), dtype=torch.float32, device=self.device)
            dno = torch.empty((T, self.B), dtype=torch.uint8, device=self.device)
            with torch.cuda.device(self.device):
                _lib.call("rl_nstep_return_f32", _lib.ptr(rew), _lib.ptr(dn), _lib.ptr(gpow), _lib.ptr(ret),
                          _lib.ptr(dno), T + nm1, self.B, n, 0, _lib.stream())
            self.samples_return_[rows[:-nm1]] = ret
            self.samples_done_n[rows[:-nm1]] = dno.view(torch.bool)
