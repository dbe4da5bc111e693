"""GPU parity of the fused PPO/A2C loss kernel (forward scalars and both gradients) against the
reference's outputs in tests/golden/loss.npz: 1e-5 relative (north_star) with a 1e-7 absolute
floor on the gradients (entries that are exactly 0 in the reference, e.g. masked rows)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from test_oracle_loss import CASES, load_case  # noqa: E402

RTOL = 1e-5


def _c(x, dtype=None):
    if x is None:
        return None
    t = torch.from_numpy(np.ascontiguousarray(x)).cuda()
    return t if dtype is None else t.to(dtype)


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("kind", ["ppo", "a2c"])
def test_fused_loss_vs_reference(golden, name, kind):
    from rlpyt_b200.algos.pg import loss_ops
    g = golden("loss")
    c = load_case(g, name)
    p = _c(c["p_new"]).requires_grad_(True)
    v = _c(c["value"]).requires_grad_(True)
    if kind == "ppo":
        loss, sc = loss_ops.ppo_loss(p, v, _c(c["p_old"]), _c(c["action"]), _c(c["ret"]), _c(c["adv"]),
                                     _c(c["valid"]), c["clip"], c["c_v"], c["c_ent"])
    else:
        loss, sc = loss_ops.a2c_loss(p, v, _c(c["action"]), _c(c["ret"]), _c(c["adv"]), _c(c["valid"]),
                                     c["c_v"], c["c_ent"])
    loss.backward()
    want = g[f"{name}/{kind}/scalars"]
    got = sc.cpu().numpy()
    np.testing.assert_allclose(got[:3], want, rtol=RTOL, atol=1e-7)
    assert got[0] == loss.item()
    gp, gv = g[f"{name}/{kind}/grad_prob"], g[f"{name}/{kind}/grad_value"]
    scale = max(1e-7, 1e-5 * float(np.abs(gp).max()))
    np.testing.assert_allclose(p.grad.cpu().numpy(), gp, rtol=RTOL, atol=scale)
    np.testing.assert_allclose(v.grad.cpu().numpy(), gv, rtol=RTOL, atol=1e-9)
    if c["valid"] is not None:  # masked rows get exactly zero gradient
        dead = c["valid"] == 0
        assert np.all(p.grad.cpu().numpy()[dead] == 0) and np.all(v.grad.cpu().numpy()[dead] == 0)


def test_upstream_gradient_scales():
    from rlpyt_b200.algos.pg import loss_ops
    from oracle import pg_loss as L
    rng = np.random.default_rng(5)
    N, A = 300, 5
    p_np = rng.dirichlet(np.ones(A), N).astype(np.float32)
    po_np = rng.dirichlet(np.ones(A), N).astype(np.float32)
    v_np, R, Ad = (rng.standard_normal(N).astype(np.float32) for _ in range(3))
    a_np = rng.integers(0, A, N)
    o = L.ppo_loss(p_np, v_np, po_np, a_np, R, Ad, None, 0.2, 0.7, 0.02)
    p = _c(p_np).requires_grad_(True)
    v = _c(v_np).requires_grad_(True)
    loss, _ = loss_ops.ppo_loss(p, v, _c(po_np), _c(a_np), _c(R), _c(Ad), None, 0.2, 0.7, 0.02)
    (3.0 * loss).backward()
    np.testing.assert_allclose(p.grad.cpu().numpy(), 3 * o["grad_prob"], rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(v.grad.cpu().numpy(), 3 * o["grad_value"], rtol=1e-5, atol=1e-9)
