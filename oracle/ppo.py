"""Oracle: one PPO / A2C iteration on the CPU (torch-CPU fp32), the arithmetic of
``PPO.optimize_agent`` (rlpyt/algos/pg/ppo.py:59-115) and ``A2C.optimize_agent``
(rlpyt/algos/pg/a2c.py:41-61) with the AtariFf network of oracle/atari_ff.py.

Test infrastructure and the ``cpu_baseline`` / ``--impl reference`` legs of bench.py only.
Uses: oracle.returns.process_returns (pg/base.py:41-75), oracle.pg_loss arithmetic, the GLOBAL numpy
RNG for the minibatch shuffles (rlpyt/utils/misc.py:6-17), torch.nn.utils.clip_grad_norm_ and
torch.optim.Adam exactly as the reference calls them (ppo.py:101-104).
"""
import numpy as np
import torch

from oracle import atari_ff, pg_loss, returns


def _params(sd):
    return {k: v.clone().requires_grad_(True) for k, v in sd.items()}


class PpoOracle:
    """Holds the network parameters + Adam state across iterations."""

    def __init__(self, state_dict, discount=0.99, learning_rate=1e-3, value_loss_coeff=1., entropy_loss_coeff=0.01,
                 clip_grad_norm=1., gae_lambda=1, minibatches=4, epochs=4, ratio_clip=0.1,
                 linear_lr_schedule=True, normalize_advantage=False, n_itr=1, mid_batch_reset=True):
        self.p = _params(state_dict)
        self.opt = torch.optim.Adam(list(self.p.values()), lr=learning_rate)
        self.discount, self.gae_lambda = discount, gae_lambda
        self.c_v, self.c_ent, self.clip_norm = value_loss_coeff, entropy_loss_coeff, clip_grad_norm
        self.minibatches, self.epochs = minibatches, epochs
        self.ratio_clip = self._ratio_clip = ratio_clip
        self.linear_lr_schedule, self.n_itr = linear_lr_schedule, n_itr
        self.normalize_advantage, self.mid_batch_reset = normalize_advantage, mid_batch_reset
        if linear_lr_schedule:
            self.sched = torch.optim.lr_scheduler.LambdaLR(self.opt, lambda itr: (n_itr - itr) / n_itr)

    def state_dict(self):
        return {k: v.detach().clone() for k, v in self.p.items()}

    def optimize_agent(self, itr, obs, action, reward, done, value, old_prob, bootstrap_value, max_updates=None):
        """obs [T,B,C,H,W] u8, action [T,B] i64, reward/value [T,B] f32, done [T,B] bool,
        old_prob [T,B,A], bootstrap_value [1,B].  Returns dict of per-update lists (OptInfo).
        ``max_updates`` (tests only): stop after that many minibatch updates - the full-size parity test checks the
        first update of a [128,256] iteration without paying for the other fifteen on the CPU."""
        T, B = reward.shape
        ret, adv, valid = returns.process_returns(
            np.asarray(reward), np.asarray(done), np.asarray(value), np.asarray(bootstrap_value),
            self.discount, self.gae_lambda, use_valid=not self.mid_batch_reset,
            normalize_advantage=self.normalize_advantage)                           # ppo.py:75
        obs_t = torch.as_tensor(np.asarray(obs))
        act_t = torch.as_tensor(np.asarray(action))
        ret_t, adv_t = torch.from_numpy(ret), torch.from_numpy(adv)
        valid_t = None if valid is None else torch.from_numpy(valid)
        oldp_t = torch.as_tensor(np.asarray(old_prob))
        batch_size = T * B
        mb_size = batch_size // self.minibatches                                     # ppo.py:90-91
        info = dict(loss=[], gradNorm=[], entropy=[], perplexity=[])
        for _ in range(self.epochs):
            indexes = np.arange(batch_size)
            np.random.shuffle(indexes)                                               # misc.py:10-12
            for start in range(0, batch_size - mb_size + 1, mb_size):
                if max_updates is not None and len(info["loss"]) >= max_updates:
                    return info
                idxs = indexes[start:start + mb_size]
                Ti, Bi = idxs % T, idxs // T                                         # ppo.py:94-95
                self.opt.zero_grad()
                pi, v = atari_ff.forward(self.p, obs_t[Ti, Bi])
                vm = None if valid_t is None else valid_t[Ti, Bi]
                ratio = pg_loss.likelihood_ratio(act_t[Ti, Bi], oldp_t[Ti, Bi], pi)
                A = adv_t[Ti, Bi]
                surr = torch.min(ratio * A, torch.clamp(ratio, 1. - self.ratio_clip, 1. + self.ratio_clip) * A)
                pi_loss = -pg_loss.valid_mean(surr, vm)
                value_loss = self.c_v * pg_loss.valid_mean(0.5 * (v - ret_t[Ti, Bi]) ** 2, vm)
                ent_i = pg_loss.entropy(pi)
                ent = pg_loss.valid_mean(ent_i, vm)
                loss = pi_loss + value_loss + (-self.c_ent * ent)
                perplexity = pg_loss.valid_mean(torch.exp(ent_i), vm)
                loss.backward()                                                      # ppo.py:101
                gn = torch.nn.utils.clip_grad_norm_(list(self.p.values()), self.clip_norm)
                self.opt.step()
                info["loss"].append(loss.item())
                info["gradNorm"].append(float(gn))
                info["entropy"].append(ent.item())
                info["perplexity"].append(perplexity.item())
        if self.linear_lr_schedule:                                                  # ppo.py:110-113
            self.sched.step()
            self.ratio_clip = self._ratio_clip * (self.n_itr - itr) / self.n_itr
        return info


def gae_plus_loss_cpu(reward, value, done, bv, discount, gae_lambda, loss_case, n_updates=16):
    """The "reference CPU GAE + PPO-loss" unit of north_star: one GAE over [T,B] with torch-CPU
    tensors (the path PPO.optimize_agent really takes) + n_updates x loss fwd+bwd arithmetic on a
    minibatch (no network).  Returns nothing; bench.py times it."""
    r, v = torch.from_numpy(reward), torch.from_numpy(value)
    d, b = torch.from_numpy(done.astype(np.float32)), torch.from_numpy(bv)
    T = r.shape[0]
    g, gl = discount, discount * gae_lambda
    nd = 1 - d
    adv = torch.zeros_like(r)
    adv[-1] = r[-1] + g * b * nd[-1] - v[-1]
    for t in reversed(range(T - 1)):                                                 # utils.py:36-38
        delta = r[t] + g * v[t + 1] * nd[t] - v[t]
        adv[t] = delta + gl * nd[t] * adv[t + 1]
    ret = adv + v
    for _ in range(n_updates):
        pg_loss.ppo_loss(*loss_case)
    return adv, ret
