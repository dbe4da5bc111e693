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

    def sample(self, dist_info):
        """``torch.multinomial`` over the trailing dim (categorical.py:25-30)."""
        p = dist_info.prob
        draw = torch.multinomial(p.reshape(-1, self.dim), num_samples=1)
        return draw.view(p.shape[:-1]).type(self.dtype)

    def entropy(self, dist_info):
        p = dist_info.prob
        return -torch.sum(p * torch.log(p + EPS), dim=-1)

    def log_likelihood(self, indexes, dist_info):
        return torch.log(select_at_indexes(indexes, dist_info.prob) + EPS)

    def likelihood_ratio(self, indexes, old_dist_info, new_dist_info):
        num = select_at_indexes(indexes, new_dist_info.prob)
        den = select_at_indexes(indexes, old_dist_info.prob)
        return (num + EPS) / (den + EPS)
