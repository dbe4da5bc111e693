"""Prioritized sequence replay (mirror of ``rlpyt/replays/sequence/prioritized.py:15-120``): the device sum-tree
holds one leaf per stored RNN state (``T // rnn_state_interval`` rows), priorities refer to the start of a whole
sequence, and new samples may bring their own priorities (shifted back ``input_priority_shift`` rows)."""
import math

import numpy as np
import torch

from rlpyt_b200 import _lib
from rlpyt_b200.replays.non_sequence.prioritized import PrioritizedReplay
from rlpyt_b200.replays.sequence.n_step import SamplesFromReplay, SequenceNStepReturnBuffer
from rlpyt_b200.replays.sum_tree import SumTree
from rlpyt_b200.utils.collections import namedarraytuple

SamplesFromReplayPri = namedarraytuple("SamplesFromReplayPri", SamplesFromReplay._fields + ("is_weights",))


class PrioritizedSequenceReplay:

    def __init__(self, alpha=0.6, beta=0.4, default_priority=1, unique=False, input_priorities=False,
                 input_priority_shift=0, pow_on_host=False, **kwargs):
        super().__init__(**kwargs)
        self.alpha, self.beta, self.default_priority, self.unique = alpha, beta, default_priority, unique
        self.input_priorities, self.input_priority_shift = input_priorities, input_priority_shift
        self.pow_on_host = pow_on_host          # see PrioritizedReplay.__init__: numpy's SIMD float32 pow vs the device kernel
        assert self.batch_T is not None, "Must assign fixed batch_T for prioritized."
        self.init_priority_tree()

    def init_priority_tree(self):
        """sequence/prioritized.py:60-73."""
        rsi = max(1, self.rnn_state_interval)
        self.priority_tree = SumTree(
            T=self.T // rsi, B=self.B,
            off_backward=math.ceil((1 + self.off_backward + self.batch_T) / rsi),
            off_forward=math.ceil(self.off_forward / rsi),
            default_value=self.default_priority ** self.alpha,
            enable_input_priorities=self.input_priorities, input_priority_shift=self.input_priority_shift,
            device=self.device)

    def set_beta(self, beta):
        self.beta = beta

    _pow_alpha = PrioritizedReplay._pow_alpha

    def append_samples(self, samples):
        """sequence/prioritized.py:78-100 (NB: unlike the non-sequence buffer, input priorities enter the tree as
        given - no ``** alpha``)."""
        if hasattr(samples, "priorities"):
            priorities = samples.priorities
            samples = samples.samples
        else:
            priorities = None
        t, rsi = self.t, self.rnn_state_interval
        T, idxs = super().append_samples(samples)
        if priorities is not None:
            priorities = torch.as_tensor(priorities).to(self.device, dtype=torch.float64)
        if rsi <= 1:
            self.priority_tree.advance(T, priorities=priorities)
        else:
            if priorities is not None and priorities.dim() == 2:
                priorities = priorities[(rsi - t) % rsi::rsi]
            n = self.t // rsi - t // rsi
            if self.t < t:
                n += self.T // rsi
            self.priority_tree.advance(n, priorities=priorities)
        return T, idxs

    def sample_batch(self, batch_B, random_values=None):
        """sequence/prioritized.py:102-115.  ``random_values`` optionally injects the uniforms (tests)."""
        (T_idxs, B_idxs), priorities = self.priority_tree.sample(batch_B, unique=self.unique, random_values=random_values)
        if self.rnn_state_interval > 1:
            T_idxs = T_idxs * self.rnn_state_interval
        batch = self.extract_batch(T_idxs, B_idxs, self.batch_T)
        is_weights = torch.empty(batch_B, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.call("rl_is_weights_eps_f32", _lib.ptr(priorities.contiguous()), float(self.beta), 0.0, _lib.ptr(is_weights),
                      int(batch_B), _lib.stream())
        return SamplesFromReplayPri(*batch, is_weights=is_weights)

    def update_batch_priorities(self, priorities):
        """sequence/prioritized.py:117-119."""
        if getattr(self, "pow_on_host", False):                # numpy's own float32 pow on the host (bit-for-bit replay of a recorded stream)
            self.priority_tree.update_batch_priorities(self._pow_alpha(priorities).reshape(-1))
        else:                                                  # pow folded into the tree update (csrc/sumtree.cu)
            self.priority_tree.update_batch_priorities(priorities, alpha=self.alpha)


class PrioritizedSequenceReplayBuffer(PrioritizedSequenceReplay, SequenceNStepReturnBuffer):
    pass
