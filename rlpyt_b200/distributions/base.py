"""Distribution interface (mirror of ``rlpyt/distributions/base.py:9-68``)."""
import torch

from rlpyt_b200.utils.tensor import valid_mean


class Distribution:
    """Methods take ``dist_info`` namedarraytuples of tensors with arbitrary leading dims."""

    @property
    def dim(self):
        raise NotImplementedError

    def sample(self, dist_info):
        raise NotImplementedError

    def kl(self, old_dist_info, new_dist_info):
        raise NotImplementedError

    def mean_kl(self, old_dist_info, new_dist_info, valid=None):
        raise NotImplementedError

    def log_likelihood(self, x, dist_info):
        raise NotImplementedError

    def likelihood_ratio(self, x, old_dist_info, new_dist_info):
        raise NotImplementedError

    def entropy(self, dist_info):
        raise NotImplementedError

    def perplexity(self, dist_info):
        return torch.exp(self.entropy(dist_info))

    def mean_entropy(self, dist_info, valid=None):
        return valid_mean(self.entropy(dist_info), valid)

    def mean_perplexity(self, dist_info, valid=None):
        return valid_mean(self.perplexity(dist_info), valid)
