#!/usr/bin/env python
"""Generate golden fixtures by CALLING the unmodified reference (never copying it).

Run in the build container only (the GPU box has no /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py [--only returns,loss,...]

Writes tests/golden/<group>.npz (compressed).  Each group stores the exact inputs and
the reference's outputs, so the tests never need the reference at run time.
"""
import argparse
import os
import sys
import types

import numpy as np

REF = os.environ.get("RLPYT_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def _import_reference():
    sys.dont_write_bytecode = True
    if REF not in sys.path:
        sys.path.insert(0, REF)
    # rlpyt.utils.prog_bar needs pyprind (absent, no network): 6-line stub.
    if "pyprind" not in sys.modules:
        stub = types.ModuleType("pyprind")

        class ProgBar:  # noqa: D401 - stub
            def __init__(self, *a, **k):
                pass

            def update(self, *a, **k):
                pass

            def stop(self):
                pass
        stub.ProgBar = ProgBar
        sys.modules["pyprind"] = stub


def _save(name, arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"wrote {path}: {len(arrays)} arrays, {os.path.getsize(path)/1024:.1f} KiB")


# --------------------------------------------------------------------------- returns
def returns_inputs(seed, T, B, p_done=0.01, sparse=False, pattern=None):
    """Synthetic inputs of SURVEY.md section 8(d) row 2."""
    rng = np.random.default_rng(seed)
    if sparse:
        reward = rng.choice(np.array([-1, 0, 1], dtype=np.float32), size=(T, B),
                            p=[0.02, 0.96, 0.02]).astype(np.float32)
    else:
        reward = rng.standard_normal((T, B), dtype=np.float32)
    value = rng.standard_normal((T, B), dtype=np.float32)
    done = rng.random((T, B)) < p_done
    bv = rng.standard_normal((1, B), dtype=np.float32)
    if pattern == "edges":  # done at t=0, t=T-1, an all-done column, a never-done column
        done[:] = False
        done[0, 0] = True
        done[T - 1, 1 % B] = True
        done[:, 2 % B] = True
        if B > 3:
            done[T // 2, 3] = True
            done[T // 2 + 1 if T // 2 + 1 < T else T - 1, 3] = True
    return reward, value, done, bv


RETURNS_CASES = [
    # name, seed, T, B, p_done, sparse, pattern
    ("kat", None, 4, 2, None, False, None),
    ("t1", 11, 1, 3, 0.3, False, None),
    ("t2b1", 12, 2, 1, 0.3, False, None),
    ("cfg1", 13, 5, 8, 0.1, False, None),
    ("edges", 14, 16, 7, 0.0, False, "edges"),
    ("ragged", 15, 37, 61, 0.05, False, None),
    ("sparse", 16, 64, 40, 0.01, True, None),
    ("cfg2", 0, 128, 256, 0.01, False, None),
]
GAMMAS_LAMBDAS = [(0.99, 1.0), (0.99, 0.98), (0.99, 0.95), (1.0, 0.9), (0.9, 0.8)]
NSTEPS = [1, 2, 3, 5]


def gen_returns():
    import torch
    from rlpyt.algos.utils import (discount_return, generalized_advantage_estimation,
                                   discount_return_n_step, valid_from_done)
    out = {}
    for name, seed, T, B, p_done, sparse, pattern in RETURNS_CASES:
        if name == "kat":  # SURVEY.md 9.1 hand-checkable vector
            reward = np.array([[1, 0], [0, 2], [1, 1], [0.5, -1]], dtype=np.float32)
            value = np.array([[0.5, 0.1], [0.2, 0.3], [0.0, -0.2], [1.0, 0.4]], dtype=np.float32)
            done = np.array([[0, 0], [1, 0], [0, 0], [0, 1]], dtype=bool)
            bv = np.array([[2.0, 3.0]], dtype=np.float32)
        else:
            reward, value, done, bv = returns_inputs(seed, T, B, p_done, sparse, pattern)
        out[f"{name}/reward"], out[f"{name}/value"] = reward, value
        out[f"{name}/done"], out[f"{name}/bv"] = done, bv
        done_f = done.astype(np.float32)  # pg/base.py:51 casts done to reward dtype
        big = T * B > 8192  # keep the committed fixture small: fewer variants at full size
        for (g, lam) in (GAMMAS_LAMBDAS[:3] if big else GAMMAS_LAMBDAS):
            key = f"{name}/g{g}_l{lam}"
            # numpy path
            adv, ret = generalized_advantage_estimation(reward, value, done_f, bv, g, lam)
            # torch-CPU path (the one PPO really takes) must agree bit-for-bit
            adv_t, ret_t = generalized_advantage_estimation(
                torch.from_numpy(reward), torch.from_numpy(value), torch.from_numpy(done_f),
                torch.from_numpy(bv), g, lam)
            assert np.array_equal(adv, adv_t.numpy()) and np.array_equal(ret, ret_t.numpy())
            out[key + "/gae_adv"], out[key + "/gae_ret"] = adv, ret
            dr = discount_return(reward, done_f, bv, g)
            dr_t = discount_return(torch.from_numpy(reward), torch.from_numpy(done_f),
                                   torch.from_numpy(bv), g)
            assert np.array_equal(dr, dr_t.numpy())
            out[key + "/disc_ret"] = dr
        out[f"{name}/valid"] = valid_from_done(torch.from_numpy(done_f)).numpy()
        for n in ([3] if big else NSTEPS):
            for trunc in (False, True):
                if not trunc and T - (n - 1) < 1:
                    continue
                for g in ((0.99,) if big else (0.99, 0.5)):
                    r_, dn_ = discount_return_n_step(reward, done, n, g, do_truncated=trunc)
                    out[f"{name}/n{n}_t{int(trunc)}_g{g}/ret"] = np.asarray(r_, dtype=np.float32)
                    out[f"{name}/n{n}_t{int(trunc)}_g{g}/done_n"] = np.asarray(dn_)
    # process_returns through the reference PolicyGradientAlgo (normalisation, valid mask)
    from rlpyt.algos.pg.base import PolicyGradientAlgo
    from collections import namedtuple
    S = namedtuple("S", "env agent")
    E = namedtuple("E", "reward done")
    A = namedtuple("A", "agent_info bootstrap_value")
    I = namedtuple("I", "value")
    for name in ("cfg1", "edges", "ragged", "cfg2"):
        reward, value, done, bv = (out[f"{name}/{k}"] for k in ("reward", "value", "done", "bv"))
        samples = S(env=E(torch.from_numpy(reward), torch.from_numpy(done)),
                    agent=A(I(torch.from_numpy(value)), torch.from_numpy(bv)))
        for lam in ((0.98,) if name == "cfg2" else (1.0, 0.98)):
            for mid_batch_reset in (True, False):
                for norm in (False, True):
                    algo = PolicyGradientAlgo()
                    algo.discount, algo.gae_lambda = 0.99, lam
                    algo.normalize_advantage = norm
                    algo.mid_batch_reset = mid_batch_reset
                    algo.agent = types.SimpleNamespace(recurrent=False)
                    ret, adv, valid = algo.process_returns(samples)
                    key = f"{name}/pr_l{lam}_m{int(mid_batch_reset)}_n{int(norm)}"
                    out[key + "/ret"], out[key + "/adv"] = ret.numpy(), adv.numpy()
                    if valid is not None:
                        out[key + "/valid"] = valid.numpy()
    _save("returns", out)


GROUPS = {"returns": gen_returns}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    _import_reference()
    only = [s for s in args.only.split(",") if s]
    for name, fn in GROUPS.items():
        if only and name not in only:
            continue
        fn()


if __name__ == "__main__":
    main()
