"""Prioritized replay (mirror of ``rlpyt/replays/non_sequence/prioritized.py:15-84``): device sum-tree,
fused extraction, importance weights and priority updates without leaving the GPU."""
import numpy as np
import torch

from rlpyt_b200 import _lib
from rlpyt_b200.replays.non_sequence.n_step import NStepReturnBuffer, SamplesFromReplay
from rlpyt_b200.replays.sum_tree import SumTree
from rlpyt_b200.utils.collections import namedarraytuple

EPS = 1e-6

SamplesFromReplayPri = namedarraytuple("SamplesFromReplayPri", SamplesFromReplay._fields + ("is_weights",))


class PrioritizedReplay:

    def __init__(self, alpha=0.6, beta=0.4, default_priority=1, unique=False, input_priorities=False,
                 input_priority_shift=0, pow_on_host=False, **kwargs):
        """``pow_on_host``: evaluate ``priorities ** alpha`` with numpy on the host like the reference
        (one small D2H/H2D round trip + sync per update) instead of the device kernel.  numpy's
        float32 ``power`` is SIMD/SVML-dispatched and differs from the correctly rounded result by
        1 ulp on ~20 % of inputs depending on the host CPU, so only this mode reproduces a recorded
        reference stream bit-for-bit; the default device kernel returns the correctly rounded value."""
        super().__init__(**kwargs)
        self.pow_on_host = pow_on_host
        self.alpha, self.beta = alpha, beta
        self.default_priority = default_priority
        self.unique = unique
        self.input_priorities = input_priorities
        self.input_priority_shift = input_priority_shift
        self.init_priority_tree()

    def init_priority_tree(self):
        self.priority_tree = SumTree(
            T=self.T, B=self.B, off_backward=self.off_backward, off_forward=self.off_forward,
            default_value=self.default_priority ** self.alpha,
            enable_input_priorities=self.input_priorities,
            input_priority_shift=self.input_priority_shift, device=self.device)

    def set_beta(self, beta):
        self.beta = beta

    def append_samples(self, samples):
        """prioritized.py:46-58."""
        if hasattr(samples, "priorities"):
            priorities = self._pow_alpha(samples.priorities)
            samples = samples.samples
        else:
            priorities = None
        T, idxs = super().append_samples(samples)
        self.priority_tree.advance(T, priorities=priorities)
        return T, idxs

    def _pow_alpha(self, priorities):
        """``priorities ** alpha`` as numpy float32 pow, widened to fp64 (prioritized.py:49,79)."""
        if getattr(self, "pow_on_host", False):
            host = priorities.detach().cpu().numpy() if isinstance(priorities, torch.Tensor) else np.asarray(priorities)
            return torch.from_numpy((host ** self.alpha).astype(np.float64)).to(self.device)
        p = torch.as_tensor(priorities).to(self.device, dtype=torch.float32).contiguous()
        out = torch.empty(p.shape, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            _lib.call("rl_pow_f32_to_f64", _lib.ptr(p), float(np.float32(self.alpha)), _lib.ptr(out), p.numel(),
                      _lib.stream())
        return out

    def sample_batch(self, batch_B, random_values=None):
        """prioritized.py:60-71.  ``random_values`` optionally injects the uniforms (tests)."""
        (T_idxs, B_idxs), priorities = self.priority_tree.sample(batch_B, unique=self.unique,
                                                                 random_values=random_values)
        batch = self.extract_batch(T_idxs, B_idxs)
        is_weights = torch.empty(batch_B, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.call("rl_is_weights_f32", _lib.ptr(priorities.contiguous()), float(self.beta), _lib.ptr(is_weights),
                      int(batch_B), _lib.stream())
        self._last_idxs = (T_idxs, B_idxs)
        return SamplesFromReplayPri(*batch, is_weights=is_weights)

    def update_batch_priorities(self, priorities):
        """prioritized.py:73-79."""
        if getattr(self, "pow_on_host", False):                # numpy's own float32 pow on the host (bit-for-bit replay of a recorded stream)
            self.priority_tree.update_batch_priorities(self._pow_alpha(priorities).reshape(-1))
        else:                                                  # pow folded into the tree update (csrc/sumtree.cu)
            self.priority_tree.update_batch_priorities(priorities, alpha=self.alpha)


class PrioritizedReplayBuffer(PrioritizedReplay, NStepReturnBuffer):
    pass
