"""Epsilon-greedy action selection from Q-values (mirror of ``rlpyt/distributions/epsilon_greedy.py:7-33``).

Same draw structure as the reference (``rand`` mask over the leading dims, ``randint`` for the masked
entries), evaluated on the device the Q-values live on so ``agent.step`` never leaves the GPU; the
CPU and CUDA generators differ, so action draws are statistically - not bitwise - those of the
reference (as for ``Categorical.sample``)."""
import torch

from rlpyt_b200.distributions.base import Distribution
from rlpyt_b200.distributions.categorical import DiscreteMixin


class EpsilonGreedy(DiscreteMixin, Distribution):

    def __init__(self, epsilon=1, **kwargs):
        super().__init__(**kwargs)
        self._epsilon = epsilon

    def sample(self, q):
        """q [T,B,A] or [B,A]; a vector epsilon of length B applies across the batch dim."""
        arg_select = torch.argmax(q, dim=-1)
        eps = self._epsilon
        if isinstance(eps, torch.Tensor):
            eps = eps.to(q.device)
        mask = torch.rand(arg_select.shape, device=q.device) < eps
        arg_rand = torch.randint(low=0, high=q.shape[-1], size=arg_select.shape, device=q.device)
        return torch.where(mask, arg_rand, arg_select)   # no data-dependent shape: graph-capturable

    @property
    def epsilon(self):
        return self._epsilon

    def set_epsilon(self, epsilon):
        self._epsilon = epsilon
