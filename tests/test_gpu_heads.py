"""Policy + value heads as one op (csrc/pg_heads.cu, models/heads_op.py; reference rlpyt/models/pg/atari_ff_model.py:56-61:
two torch.nn.Linear + softmax): forward and every gradient against the same arithmetic in float64."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,F,A", [(8192, 512, 6), (1, 512, 6), (37, 63, 18), (300, 360, 32), (130, 1000, 10), (1025, 256, 4), (64, 512, 9)])
def test_policy_value_heads_forward_backward_vs_float64(N, F, A):
    from rlpyt_b200.models.heads_op import PgHeads
    g = torch.Generator(device="cuda").manual_seed(N + F + A)
    rn = lambda *s: torch.randn(*s, device="cuda", generator=g)
    h = torch.relu(rn(N, F)).requires_grad_(True)
    w_pi, b_pi = (rn(A, F) / F ** 0.5).requires_grad_(True), (rn(A) * 0.1).requires_grad_(True)
    w_v, b_v = (rn(1, F) / F ** 0.5).requires_grad_(True), (rn(1) * 0.1).requires_grad_(True)
    g_pi, g_v = rn(N, A), rn(N)
    pi, v = PgHeads.apply(h, w_pi, b_pi, w_v, b_v)
    torch.autograd.backward([pi, v], [g_pi, g_v])
    got = [pi, v, h.grad, w_pi.grad, b_pi.grad, w_v.grad, b_v.grad]
    d = lambda t: t.detach().double().requires_grad_(True)
    h2, wp2, bp2, wv2, bv2 = d(h), d(w_pi), d(b_pi), d(w_v), d(b_v)
    pi2 = torch.softmax(h2 @ wp2.t() + bp2, -1)
    v2 = (h2 @ wv2.t() + bv2).squeeze(-1)
    torch.autograd.backward([pi2, v2], [g_pi.double(), g_v.double()])
    want = [pi2, v2, h2.grad, wp2.grad, bp2.grad, wv2.grad, bv2.grad]
    names = ["pi", "v", "grad_h", "grad_w_pi", "grad_b_pi", "grad_w_v", "grad_b_v"]
    assert torch.allclose(pi.sum(-1), torch.ones(N, device="cuda"), atol=1e-6)
    for name, a, b in zip(names, got, want):
        a, b = a.detach().double().cpu().numpy(), b.detach().cpu().numpy()
        assert a.shape == b.shape, name
        scale = max(1e-30, float(np.abs(b).max()))
        # fp32 dot products / row sums of up to 8192 terms against float64: a few 1e-6 of the largest entry
        assert float(np.abs(a - b).max()) <= 2e-5 * scale, (name, float(np.abs(a - b).max()), scale)


def test_heads_with_one_output_unused_and_determinism():
    """Only the policy (or only the value) receives a gradient: the missing upstream gradient counts as zero; two runs are
    bit-identical (fixed-order reductions)."""
    from rlpyt_b200.models.heads_op import PgHeads
    g = torch.Generator(device="cuda").manual_seed(3)
    h = torch.randn(513, 512, device="cuda", generator=g)
    w_pi, b_pi = torch.randn(6, 512, device="cuda", generator=g) / 22, torch.zeros(6, device="cuda")
    w_v, b_v = torch.randn(1, 512, device="cuda", generator=g) / 22, torch.zeros(1, device="cuda")
    outs = []
    for rep in range(2):
        leaves = [t.clone().requires_grad_(True) for t in (h, w_pi, b_pi, w_v, b_v)]
        pi, v = PgHeads.apply(*leaves)
        pi[:, 0].sum().backward()
        outs.append([t.grad.clone() for t in leaves])
    assert all(torch.equal(a, b) for a, b in zip(*outs))
    assert float(outs[0][3].abs().max()) == 0.0 and float(outs[0][4].abs().max()) == 0.0      # value head untouched
    leaves = [t.clone().requires_grad_(True) for t in (h, w_pi, b_pi, w_v, b_v)]
    pi, v = PgHeads.apply(*leaves)
    v.sum().backward()
    assert float(leaves[1].grad.abs().max()) == 0.0
    np.testing.assert_allclose(leaves[3].grad.cpu().numpy().reshape(-1), h.double().sum(0).cpu().numpy(), rtol=1e-5, atol=1e-4)


def test_model_with_fused_heads_matches_torch_heads(monkeypatch):
    from rlpyt_b200.models.pg.atari_ff_model import AtariFfModel
    torch.manual_seed(2)
    model = AtariFfModel((4, 84, 84), 6).cuda()
    g = torch.Generator(device="cuda").manual_seed(1)
    obs = torch.randint(0, 256, (640, 4, 84, 84), dtype=torch.uint8, device="cuda", generator=g)
    gp, gv = torch.randn(640, 6, device="cuda", generator=g), torch.randn(640, device="cuda", generator=g)

    def run(fused):
        monkeypatch.setenv("RLPYT_B200_FUSED_HEADS", "1" if fused else "0")
        model.zero_grad(set_to_none=True)
        pi, v = model(obs, None, None)
        torch.autograd.backward([pi, v], [gp, gv])
        return [pi.detach(), v.detach()] + [p.grad.clone() for p in model.parameters()]

    a, b = run(False), run(True)
    for x, y in zip(a, b):
        scale = max(1e-30, float(x.abs().max()))
        assert float((x - y).abs().max()) <= 2e-5 * scale
