"""Shared loader for the DQN-loss golden cases (tests/golden/dqn.npz, written by make_golden.py)."""
import numpy as np

CASES = ["dqn_small", "dqn_double_pri", "dqn_pri", "dqn_mse", "dqn_clip_frac", "dqn_n1"]


def load_case(g, name):
    d = {k: g[f"{name}/{k}"] for k in ("qs", "target_qs", "next_qs", "action", "return_", "done_n", "is_weights")}
    double, pri, clip, n_step, discount = (float(x) for x in g[f"{name}/hyper"])
    d["double"], d["pri"] = bool(double), bool(pri)
    d["clip"] = None if clip < 0 else clip
    d["n_step"], d["discount"] = int(n_step), discount
    d["loss"] = float(g[f"{name}/loss"][0])
    d["td_abs_errors"] = g[f"{name}/td_abs_errors"]
    d["grad_qs"] = g[f"{name}/grad_qs"]
    return d


def kernel_arithmetic(c):
    """numpy float32 restatement of csrc/dqn_loss.cu, operation by operation (each numpy float32 op
    is one correctly rounded fp32 operation, like the __f*_rn intrinsics) - documents why the kernel's
    td_abs_errors are bit-identical to the reference's."""
    f = np.float32
    qs, tq, nq = c["qs"], c["target_qs"], c["next_qs"]
    N, A = qs.shape
    idx = np.arange(N)
    if c["double"]:
        target_q = tq[idx, np.argmax(nq, axis=-1)]          # first maximal index, like the strict '>' scan
    else:
        target_q = tq.max(axis=-1)
    q = qs[idx, c["action"]]
    disc_n = f(c["discount"] ** c["n_step"])
    not_done = np.where(c["done_n"], f(0), f(1)).astype(f)
    y = c["return_"] + not_done * (disc_n * target_q)
    delta = y - q
    ad = np.abs(delta)
    loss = f(0.5) * (delta * delta)
    dl = delta.copy()
    td = ad.copy()
    if c["clip"] is not None:
        clip = f(c["clip"])
        lin = ~(ad <= clip)
        loss = np.where(lin, clip * (ad - clip * f(0.5)), loss)
        dl = np.where(lin, np.where(delta > 0, clip, -clip), dl)
        td = np.minimum(np.maximum(ad, f(0)), clip)
    w = c["is_weights"] if c["pri"] else np.ones(N, f)
    loss = loss * w
    grad = np.zeros((N, A), f)
    grad[idx, c["action"]] = -(w / f(N)) * dl
    return float(loss.astype(np.float64).sum() / N), td.astype(f), grad
