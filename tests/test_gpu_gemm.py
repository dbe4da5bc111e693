"""The tcgen05 3xTF32 GEMM (csrc/gemm_tf32x3.cu) against an fp64 reference: fp32-level accuracy
(<= 2e-6 of the row/column scale; plain TF32 would be ~5e-4), ragged M/N/K tails, bias + ReLU
epilogue, and the autograd Linear built on it against torch's fp32 Linear."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(a, b, bias, relu):
    y = a.double() @ b.double().t()
    if bias is not None:
        y = y + bias.double()
    return torch.relu(y) if relu else y


@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (256, 512, 3200), (8192, 512, 3200), (300, 200, 100),
                                   (1, 6, 512), (129, 4, 36), (512, 3200, 8192)])
@pytest.mark.parametrize("bias,relu", [(False, False), (True, True)])
def test_gemm_tf32x3_accuracy(M, N, K, bias, relu):
    from rlpyt_b200.models.gemm_op import gemm_tn
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N)
    a = torch.randn(M, K, device="cuda", generator=g)
    b = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
    bv = torch.randn(N, device="cuda", generator=g) if bias else None
    y = gemm_tn(a, b, bv, relu)
    want = _ref(a, b, bv, relu)
    # error budget: each product carries ~2^-20 relative; sums of K terms of magnitude |a||b|
    scale = float((a.double().abs() @ b.double().abs().t()).max())
    err = float((y.double() - want).abs().max())
    assert err <= 3e-6 * scale, (err, scale)
    # and it is far better than single-pass TF32 would be
    assert err <= 1e-4 * float(want.abs().max() + 1e-6)


def test_linear_autograd_matches_torch_fp32():
    from rlpyt_b200.models.gemm_op import linear_tf32x3
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(1024, 3200, device="cuda", generator=g)
    w = (torch.randn(512, 3200, device="cuda", generator=g) / 56).requires_grad_(True)
    b = torch.randn(512, device="cuda", generator=g).requires_grad_(True)
    x1 = x.clone().requires_grad_(True)
    x2 = x.clone().requires_grad_(True)
    w2, b2 = w.detach().clone().requires_grad_(True), b.detach().clone().requires_grad_(True)
    y = linear_tf32x3(x1, w, b, relu=True)
    y_ref = torch.relu(torch.nn.functional.linear(x2, w2, b2))
    np.testing.assert_allclose(y.detach().cpu().numpy(), y_ref.detach().cpu().numpy(), rtol=1e-5, atol=2e-5)
    go = torch.randn(y.shape, device="cuda", generator=g)
    (y * go).sum().backward()
    (y_ref * go * (y.detach() > 0)).sum().backward()   # same ReLU mask (see conv1 test)
    for got, want in ((x1.grad, x2.grad), (w.grad, w2.grad), (b.grad, b2.grad)):
        s = float(want.abs().max())
        np.testing.assert_allclose(got.cpu().numpy(), want.cpu().numpy(), rtol=1e-4, atol=2e-5 * s)
