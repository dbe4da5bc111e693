"""Oracle: categorical distribution pieces and the PPO / A2C loss (forward + gradients).

Test infrastructure only (see oracle/__init__.py).  torch-CPU fp32 ops in the reference's
operation order (the reference itself evaluates this arithmetic with torch on the CPU:
rlpyt/agents/pg/categorical.py:25 moves the network outputs to "cpu" first), with autograd
providing dLoss/dprob and dLoss/dvalue.

Reference: rlpyt/algos/pg/ppo.py:117-154 (PPO.loss), rlpyt/algos/pg/a2c.py:63-103 (A2C.loss),
rlpyt/distributions/categorical.py:32-43, rlpyt/distributions/base.py:57-68,
rlpyt/utils/tensor.py:5-15 (select_at_indexes), :39-46 (valid_mean).
"""
import numpy as np
import torch

EPS = 1e-8  # rlpyt/distributions/categorical.py:9


def select_at_indexes(indexes, tensor):
    """tensor[..., indexes] along the last dim (rlpyt/utils/tensor.py:5-15)."""
    flat = tensor.reshape(-1, tensor.shape[-1])
    return flat[torch.arange(flat.shape[0]), indexes.reshape(-1).long()].reshape(indexes.shape)


def valid_mean(x, valid=None):
    """rlpyt/utils/tensor.py:39-46: plain mean, or sum(x*valid)/sum(valid)."""
    if valid is None:
        return x.mean()
    valid = valid.type(x.dtype)
    return (x * valid).sum() / valid.sum()


def entropy(prob):
    """rlpyt/distributions/categorical.py:32-34."""
    return -torch.sum(prob * torch.log(prob + EPS), dim=-1)


def likelihood_ratio(action, old_prob, new_prob):
    """rlpyt/distributions/categorical.py:40-43."""
    return (select_at_indexes(action, new_prob) + EPS) / (select_at_indexes(action, old_prob) + EPS)


def log_likelihood(action, prob):
    """rlpyt/distributions/categorical.py:36-38."""
    return torch.log(select_at_indexes(action, prob) + EPS)


def _t(x, dtype=torch.float32):
    if x is None:
        return None
    return torch.as_tensor(np.asarray(x)).to(dtype)


def ppo_loss(prob_new, value, prob_old, action, return_, advantage, valid, ratio_clip,
             value_loss_coeff, entropy_loss_coeff):
    """PPO.loss arithmetic after the network forward (rlpyt/algos/pg/ppo.py:136-153).

    Returns dict(loss, entropy, perplexity, pi_loss, value_loss, grad_prob, grad_value) as numpy.
    """
    p = _t(prob_new).clone().requires_grad_(True)
    v = _t(value).clone().requires_grad_(True)
    po, a = _t(prob_old), _t(action, torch.int64)
    R, A, vm = _t(return_), _t(advantage), _t(valid)
    ratio = likelihood_ratio(a, po, p)                                        # ppo.py:136-137
    surr_1 = ratio * A                                                        # :138
    clipped = torch.clamp(ratio, 1. - ratio_clip, 1. + ratio_clip)           # :139-140
    surr_2 = clipped * A                                                      # :141
    surrogate = torch.min(surr_1, surr_2)                                     # :142
    pi_loss = -valid_mean(surrogate, vm)                                      # :143
    value_error = 0.5 * (v - R) ** 2                                          # :145
    value_loss = value_loss_coeff * valid_mean(value_error, vm)               # :146
    ent_i = entropy(p)
    ent = valid_mean(ent_i, vm)                                               # :148 (base.py:61-64)
    loss = pi_loss + value_loss + (-entropy_loss_coeff * ent)                 # :149-151
    perplexity = valid_mean(torch.exp(ent_i), vm)                             # :153 (base.py:66-68)
    loss.backward()
    return dict(loss=loss.item(), entropy=ent.item(), perplexity=perplexity.item(),
                pi_loss=pi_loss.item(), value_loss=value_loss.item(),
                grad_prob=p.grad.numpy().copy(), grad_value=v.grad.numpy().copy())


def a2c_loss(prob, value, action, return_, advantage, valid, value_loss_coeff, entropy_loss_coeff):
    """A2C.loss arithmetic after the network forward (rlpyt/algos/pg/a2c.py:88-103)."""
    p = _t(prob).clone().requires_grad_(True)
    v = _t(value).clone().requires_grad_(True)
    a = _t(action, torch.int64)
    R, A, vm = _t(return_), _t(advantage), _t(valid)
    logli = log_likelihood(a, p)                                              # a2c.py:89
    pi_loss = -valid_mean(logli * A, vm)                                      # :90
    value_error = 0.5 * (v - R) ** 2                                          # :92
    value_loss = value_loss_coeff * valid_mean(value_error, vm)               # :93
    ent_i = entropy(p)
    ent = valid_mean(ent_i, vm)                                               # :95
    loss = pi_loss + value_loss + (-entropy_loss_coeff * ent)                 # :96-98
    perplexity = valid_mean(torch.exp(ent_i), vm)                             # :100
    loss.backward()
    return dict(loss=loss.item(), entropy=ent.item(), perplexity=perplexity.item(),
                pi_loss=pi_loss.item(), value_loss=value_loss.item(),
                grad_prob=p.grad.numpy().copy(), grad_value=v.grad.numpy().copy())


def sample_categorical(prob, uniform):
    """Inverse-CDF draw with an injected uniform per row (the oracle's definition of
    ``Categorical.sample`` for parity tests: ``torch.multinomial``'s CPU and CUDA generators
    differ, so rlpyt/distributions/categorical.py:25-30 cannot be matched draw-for-draw;
    SURVEY.md section 7 "RNG parity").  action = #{k : cumsum_fp32(p)[k] <= u}, clipped to A-1."""
    p = np.asarray(prob, dtype=np.float32)
    u = np.asarray(uniform, dtype=np.float32)
    acc = np.zeros(p.shape[:-1], dtype=np.float32)
    act = np.zeros(p.shape[:-1], dtype=np.int64)
    for k in range(p.shape[-1] - 1):
        acc = acc + p[..., k]
        act += (u >= acc)
    return act
