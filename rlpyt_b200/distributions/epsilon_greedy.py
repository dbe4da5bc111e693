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
        self._eps_dev = None          # persistent device copy of epsilon: what ``sample`` actually reads

    def _device_epsilon(self, device):
        """Epsilon as a tensor that lives on ``device`` for the lifetime of the distribution.  ``sample`` may be
        captured in a CUDA graph (samplers/rollout.py captures ``agent.step``): a Python float would be frozen into
        the captured kernel arguments and a host tensor would record a pageable H2D copy, so the annealing of
        ``EpsilonGreedyAgentMixin.sample_mode`` would silently stop working under replay.  ``set_epsilon``
        refreshes this buffer IN PLACE (outside any capture), replays read the current value."""
        eps = self._epsilon
        shape = tuple(eps.shape) if isinstance(eps, torch.Tensor) else ()
        if self._eps_dev is None or self._eps_dev.device != device or tuple(self._eps_dev.shape) != shape:
            self._eps_dev = torch.empty(shape, dtype=torch.float32, device=device)
            self._write_device_epsilon()
        return self._eps_dev

    def _write_device_epsilon(self):
        eps = self._epsilon
        if isinstance(eps, torch.Tensor):
            self._eps_dev.copy_(eps.to(torch.float32))
        else:
            self._eps_dev.fill_(float(eps))

    def sample(self, q):
        """q [T,B,A] or [B,A]; a vector epsilon of length B applies across the batch dim."""
        arg_select = torch.argmax(q, dim=-1)
        eps = self._device_epsilon(q.device)
        mask = torch.rand(arg_select.shape, device=q.device) < eps
        arg_rand = torch.randint(low=0, high=q.shape[-1], size=arg_select.shape, device=q.device)
        return torch.where(mask, arg_rand, arg_select)   # no data-dependent shape: graph-capturable

    @property
    def epsilon(self):
        return self._epsilon

    def set_epsilon(self, epsilon):
        self._epsilon = epsilon
        if self._eps_dev is not None:
            shape = tuple(epsilon.shape) if isinstance(epsilon, torch.Tensor) else ()
            if tuple(self._eps_dev.shape) == shape:
                self._write_device_epsilon()             # in place: CUDA graphs that captured ``sample`` see it
            else:
                self._eps_dev = None                     # scalar <-> per-env vector: rebuilt at the next sample
