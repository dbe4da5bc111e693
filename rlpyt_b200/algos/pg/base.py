"""Policy-gradient base: optimizer construction and ``process_returns`` on the GPU
(mirror of ``rlpyt/algos/pg/base.py:14-75``)."""
from collections import namedtuple

import numpy as np
import torch

from rlpyt_b200 import _lib
from rlpyt_b200.algos.base import RlAlgorithm
from rlpyt_b200.algos.optim import FlatAdam
from rlpyt_b200.algos.utils import normalize_advantage_

# Convention: traj_info fields CamelCase, opt_info fields lowerCamelCase (pg/base.py:9-10)
OptInfo = namedtuple("OptInfo", ["loss", "gradNorm", "entropy", "perplexity"])
AgentTrain = namedtuple("AgentTrain", ["dist_info", "value"])


class PolicyGradientAlgo(RlAlgorithm):

    bootstrap_value = True  # tells the sampler to record Value(State') (pg/base.py:21)
    opt_info_fields = tuple(OptInfo._fields)

    def initialize(self, agent, n_itr, batch_spec, mid_batch_reset=False, examples=None,
                   world_size=1, rank=0):
        """Build the optimizer (pg/base.py:24-39).  With the default ``OptimCls=FlatAdam`` the
        parameters are re-pointed into one flat buffer and the update is the fused
        all-reduce/clip/Adam step; any ``torch.optim`` class still works (unfused)."""
        self.optimizer = self.OptimCls(agent.parameters(), lr=self.learning_rate, **self.optim_kwargs)
        if isinstance(self.optimizer, FlatAdam):
            self.optimizer.set_world_size(world_size)
        if self.initial_optim_state_dict is not None:
            self.optimizer.load_state_dict(self.initial_optim_state_dict)
        self.agent = agent
        self.n_itr = n_itr
        self.batch_spec = batch_spec
        self.mid_batch_reset = mid_batch_reset
        self.rank = rank
        self.world_size = world_size
        self._ret_bufs = {}

    # ------------------------------------------------------------------------------------
    def _device(self):
        dev = getattr(self.agent, "device", None)
        if dev is None or dev.type != "cuda":
            if not torch.cuda.is_available():
                raise _lib.B200LibraryError("rlpyt_b200 algorithms need a CUDA device (no CPU fallback)")
            dev = torch.device("cuda", torch.cuda.current_device())
        return dev

    def _on_device(self, x, dtype=None):
        if not x.is_cuda:
            x = x.to(self._device(), non_blocking=True)
        if dtype is not None and x.dtype != dtype:
            x = x.to(dtype)
        return x.contiguous()

    def process_returns(self, samples):
        """Returns ``(return_, advantage, valid)`` as CUDA tensors shaped like ``reward``.

        Same decisions as pg/base.py:41-75: lambda==1 -> discounted return and
        ``advantage = return_ - value``; else GAE; ``valid`` only when ``not mid_batch_reset`` (or
        recurrent); optional normalisation with LOCAL statistics (no cross-rank reduce).  Two or
        three kernel launches instead of a T-step Python loop over torch-CPU tensors.
        """
        reward = self._on_device(samples.env.reward, torch.float32)
        done = samples.env.done
        done = self._on_device(done)
        done_u8 = done.view(torch.uint8) if done.dtype == torch.bool else (done != 0).view(torch.uint8)
        value = self._on_device(samples.agent.agent_info.value, torch.float32)
        bv = self._on_device(samples.agent.bootstrap_value, torch.float32).reshape(-1)
        T = reward.shape[0]
        B = reward.numel() // T
        key = (tuple(reward.shape), str(reward.device))
        bufs = self._ret_bufs.get(key)
        if bufs is None:
            bufs = (torch.empty_like(reward), torch.empty_like(reward), torch.empty_like(reward))
            self._ret_bufs[key] = bufs
        return_, advantage, valid_buf = bufs
        st = _lib.stream
        with torch.cuda.device(reward.device):
            if self.gae_lambda == 1:  # pg/base.py:53-55
                _lib.call("rl_discount_return_f32", _lib.ptr(reward), _lib.ptr(done_u8), _lib.ptr(bv),
                          _lib.ptr(value), _lib.ptr(return_), _lib.ptr(advantage), T, B,
                          float(self.discount), 0, st())
            else:  # pg/base.py:56-58
                gl = float(np.float32(float(self.discount) * float(self.gae_lambda)))
                _lib.call("rl_gae_f32", _lib.ptr(reward), _lib.ptr(value), _lib.ptr(done_u8), _lib.ptr(bv),
                          _lib.ptr(advantage), _lib.ptr(return_), T, B, float(self.discount), gl, 0, st())
            if not self.mid_batch_reset or self.agent.recurrent:  # pg/base.py:60-63
                valid = valid_buf
                _lib.call("rl_valid_from_done_f32", _lib.ptr(done_u8), _lib.ptr(valid), T, B, st())
            else:
                valid = None
        if self.normalize_advantage:  # pg/base.py:65-73
            normalize_advantage_(advantage, valid)
        return return_, advantage, valid
