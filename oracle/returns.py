"""Oracle: discounted / GAE / n-step returns, valid mask, advantage normalisation.

Test infrastructure only (see oracle/__init__.py).  numpy float32 arithmetic in the
reference's operation order, so results are bit-identical to the reference's numpy
and torch-CPU paths (both are IEEE fp32 element-wise ops, no fused multiply-add).

Reference: rlpyt/algos/utils.py (discount_return :8-21, generalized_advantage_estimation
:24-40, discount_return_n_step :67-101, valid_from_done :104-112) and
rlpyt/algos/pg/base.py (process_returns :41-75).
"""
import numpy as np

F32 = np.float32


def _nd(done):
    """``1 - done`` as fp32 (rlpyt/algos/utils.py:15-16, :33-34)."""
    return (F32(1) - np.asarray(done).astype(F32)).astype(F32)


def discount_return(reward, done, bootstrap_value, discount):
    """R[T-1] = r + g*bv*nd ; R[t] = r[t] + R[t+1]*g*nd[t].

    Follows rlpyt/algos/utils.py:8-21 (note the operand order of :20:
    ``return_[t + 1] * discount * nd[t]`` = (R*g)*nd, while :18 is (g*bv)*nd).
    """
    reward = np.asarray(reward, dtype=F32)
    T = reward.shape[0]
    g = F32(discount)
    nd = _nd(done)
    bv = np.asarray(bootstrap_value, dtype=F32).reshape(reward.shape[1:])
    ret = np.zeros_like(reward)
    ret[T - 1] = reward[T - 1] + (g * bv) * nd[T - 1]
    for t in range(T - 2, -1, -1):
        ret[t] = reward[t] + (ret[t + 1] * g) * nd[t]
    return ret


def generalized_advantage_estimation(reward, value, done, bootstrap_value,
                                     discount, gae_lambda):
    """A[T-1] = r + g*bv*nd - v ; delta = r + g*v[t+1]*nd - v ;
    A[t] = delta + (g*lam)*nd*A[t+1] ; R = A + v.

    Follows rlpyt/algos/utils.py:24-40.  ``discount * gae_lambda`` is a python
    double product that is cast to fp32 once when it meets the fp32 array (:38).
    """
    reward = np.asarray(reward, dtype=F32)
    value = np.asarray(value, dtype=F32)
    T = reward.shape[0]
    g = F32(discount)
    gl = F32(float(discount) * float(gae_lambda))
    nd = _nd(done)
    bv = np.asarray(bootstrap_value, dtype=F32).reshape(reward.shape[1:])
    adv = np.zeros_like(reward)
    adv[T - 1] = (reward[T - 1] + (g * bv) * nd[T - 1]) - value[T - 1]
    for t in range(T - 2, -1, -1):
        delta = (reward[t] + (g * value[t + 1]) * nd[t]) - value[t]
        adv[t] = delta + ((gl * nd[t]) * adv[t + 1])
    ret = adv + value
    return adv, ret


def discount_return_n_step(reward, done, n_step, discount, do_truncated=False):
    """n-step return and "done within n steps" flag.

    Follows rlpyt/algos/utils.py:67-101: ``return_ = reward[:rlen]``,
    ``done_n = done[:rlen]`` then for n in 1..n_step-1:
    ``return_ += (discount**n) * reward[n:n+rlen] * (1 - done_n)`` (mask taken
    BEFORE the tap, :97) and ``done_n = max(done_n, done[n:n+rlen])`` (:98).
    Returns (return_ f32 [rlen,...], done_n bool [rlen,...]).
    """
    reward = np.asarray(reward, dtype=F32)
    done = np.asarray(done).astype(bool)
    rlen = reward.shape[0]
    if not do_truncated:
        rlen -= (n_step - 1)
    ret = reward[:rlen].copy()
    done_n = done[:rlen].copy()
    for n in range(1, n_step):
        gk = F32(float(discount) ** n)
        if do_truncated:
            m = rlen - n
            if m <= 0:
                continue
            nd = F32(1) - done_n[:m].astype(F32)
            ret[:m] = ret[:m] + (gk * reward[n:n + m]) * nd
            done_n[:m] = np.maximum(done_n[:m], done[n:n + m])
        else:
            nd = F32(1) - done_n.astype(F32)
            ret = ret + (gk * reward[n:n + rlen]) * nd
            done_n = np.maximum(done_n, done[n:n + rlen])
    return ret.astype(F32), done_n


def valid_from_done(done):
    """valid[0]=1; valid[t] = 1 - min(cumsum(done[:t]), 1)  (rlpyt/algos/utils.py:104-112).
    The step on which ``done`` first fires is itself still valid."""
    d = np.asarray(done).astype(F32)
    valid = np.ones_like(d)
    if d.shape[0] > 1:
        valid[1:] = F32(1) - np.minimum(np.cumsum(d[:-1], axis=0, dtype=F32), F32(1))
    return valid


def process_returns(reward, done, value, bootstrap_value, discount, gae_lambda,
                    use_valid, normalize_advantage):
    """rlpyt/algos/pg/base.py:41-75.  Returns (return_, advantage, valid-or-None).

    ``use_valid`` = ``not mid_batch_reset or agent.recurrent`` (:60).  Normalisation
    uses the LOCAL mean and UNBIASED std over valid elements, floor 1e-6 (:65-73).
    The mean/std are evaluated with torch-CPU so they are the very numbers the
    reference produces (torch's pairwise fp32 reduction).
    """
    import torch
    if gae_lambda == 1:  # base.py:53-55
        ret = discount_return(reward, done, bootstrap_value, discount)
        adv = ret - np.asarray(value, dtype=F32)
    else:  # base.py:56-58
        adv, ret = generalized_advantage_estimation(
            reward, value, done, bootstrap_value, discount, gae_lambda)
    valid = valid_from_done(done) if use_valid else None
    if normalize_advantage:
        a = torch.from_numpy(adv.copy())
        if valid is not None:
            m = torch.from_numpy(valid) > 0
            mean, std = a[m].mean(), a[m].std()
        else:
            mean, std = a.mean(), a.std()
        adv = ((a - mean) / max(std, 1e-6)).numpy()
    return ret, adv, valid
