"""Categorical action distribution (mirror of ``rlpyt/distributions/categorical.py:14-43`` and
``rlpyt/distributions/discrete.py``).  Plain torch ops here are the generic, differentiable
API (any device); the PPO/A2C hot path does not go through them - it uses the fused kernel in
``rlpyt_b200.algos.pg.loss_ops`` - and ``agent.step`` samples inside the fused forward kernel.
"""
import torch

from rlpyt_b200.distributions.base import Distribution
from rlpyt_b200.utils.collections import namedarraytuple
from rlpyt_b200.utils.tensor import valid_mean, select_at_indexes, to_onehot, from_onehot

EPS = 1e-8  # rlpyt/distributions/categorical.py:9

DistInfo = namedarraytuple("DistInfo", ["prob"])


class DiscreteMixin:
    """One-hot conversions (rlpyt/distributions/discrete.py:7-27)."""

    def __init__(self, dim, dtype=torch.long, onehot_dtype=torch.float):
        self._dim = dim
        self.dtype = dtype
        self.onehot_dtype = onehot_dtype

    @property
    def dim(self):
        return self._dim

    def to_onehot(self, indexes, dtype=None):
        return to_onehot(indexes, self._dim, dtype=dtype or self.onehot_dtype)

    def from_onehot(self, onehot, dtype=None):
        return from_onehot(onehot, dtype=dtype or self.dtype)


class Categorical(DiscreteMixin, Distribution):

    def kl(self, old_dist_info, new_dist_info):
        p, q = old_dist_info.prob, new_dist_info.prob
        return torch.sum(p * (torch.log(p + EPS) - torch.log(q + EPS)), dim=-1)

    def mean_kl(self, old_dist_info, new_dist_info, valid=None):
        return valid_mean(self.kl(old_dist_info, new_dist_info), valid)

    def sample(self, dist_info, uniform=None):
        """A draw per row over the trailing dim (categorical.py:25-30 uses ``torch.multinomial``).  CUDA
        probabilities take the inverse-CDF kernel of csrc/categorical.cu - ``uniform`` injects the draws (parity
        tests against oracle/pg_loss.py:sample_categorical), otherwise they come from a Philox stream whose
        (seed, counter) state lives on the device (``manual_seed``; graph-capturable: replays advance it).  CPU
        tensors keep ``torch.multinomial`` (only the start-up example step runs there)."""
        p = dist_info.prob
        if not p.is_cuda:
            draw = torch.multinomial(p.reshape(-1, self.dim), num_samples=1)
            return draw.view(p.shape[:-1]).type(self.dtype)
        from rlpyt_b200 import _lib
        flat = p.reshape(-1, self.dim).contiguous()
        if flat.dtype != torch.float32:
            flat = flat.float()
        out = torch.empty(flat.shape[0], dtype=torch.int64, device=p.device)
        state = None
        if uniform is None:
            state = self._rng_state(p.device)
        else:
            uniform = uniform.reshape(-1).to(device=p.device, dtype=torch.float32).contiguous()
        with torch.cuda.device(p.device):
            _lib.call("rl_categorical_sample_f32", _lib.ptr(flat), _lib.ptr(uniform), _lib.ptr(state), _lib.ptr(out), None,
                      flat.shape[0], self.dim, _lib.stream())
        return out.view(p.shape[:-1]).type(self.dtype)

    def _rng_state(self, device):
        st = getattr(self, "_rng", None)
        if st is None or st.device != device:
            seed = int(torch.initial_seed()) & 0x7FFFFFFFFFFFFFFF       # follows torch.manual_seed / set_seed
            st = self._rng = torch.tensor([seed, 0, 0], dtype=torch.int64, device=device)     # seed, call counter, block ticket
        return st

    def manual_seed(self, seed, counter=0, device=None):
        """Re-key the device Philox stream; an existing state tensor is updated in place (captured graphs keep
        reading it)."""
        new = torch.tensor([int(seed) & 0x7FFFFFFFFFFFFFFF, int(counter), 0], dtype=torch.int64)
        st = getattr(self, "_rng", None)
        if st is None or (device is not None and st.device != torch.device(device)):
            dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
            self._rng = new.to(dev)
        else:
            st.copy_(new)

    def entropy(self, dist_info):
        p = dist_info.prob
        return -torch.sum(p * torch.log(p + EPS), dim=-1)

    def log_likelihood(self, indexes, dist_info):
        return torch.log(select_at_indexes(indexes, dist_info.prob) + EPS)

    def likelihood_ratio(self, indexes, old_dist_info, new_dist_info):
        num = select_at_indexes(indexes, new_dist_info.prob)
        den = select_at_indexes(indexes, old_dist_info.prob)
        return (num + EPS) / (den + EPS)
