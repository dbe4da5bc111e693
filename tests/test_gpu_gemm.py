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


@pytest.mark.parametrize("M,N,K,a_mmajor,c_trans", [
    (128, 128, 64, False, False), (128, 128, 64, True, False), (128, 128, 64, False, True),
    (200, 72, 100, False, False), (200, 72, 100, True, True),          # ragged tiles and a K tail
    (1000, 384, 1056, False, False),                                   # several tiles per CTA, k-split
    (256, 512, 3200, False, False),                                    # agent.step shape (k-split 6)
    (8192, 512, 3200, False, False),                                   # fc forward
    (8192, 3200, 512, False, False),                                   # fc input gradient
    (3200, 512, 8192, True, True),                                     # fc weight gradient: x as it lies, gw stored directly
])
@pytest.mark.parametrize("bias,relu", [(False, False), (True, True)])
def test_gemm_ts_accuracy(M, N, K, a_mmajor, c_trans, bias, relu):
    """The TMEM-operand kernel (csrc/gemm_ts.cuh) against fp64, same budget as the first kernel."""
    from rlpyt_b200.models.gemm_op import gemm_ts, split_lo
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N + K)
    a = torch.randn(M, K, device="cuda", generator=g)
    b = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
    bv = torch.randn(N, device="cuda", generator=g) if bias else None
    b_lo = split_lo(b)
    assert torch.equal(b_lo, b - (b.view(torch.int32) & -8192).view(torch.float32))
    y = gemm_ts(a.t().contiguous() if a_mmajor else a, b, b_lo, bv, relu, a_mmajor=a_mmajor, c_trans=c_trans)
    if c_trans:
        assert y.shape == (N, M)
        y = y.t()
    want = _ref(a, b, bv, relu)
    scale = float((a.double().abs() @ b.double().abs().t()).max())
    err = float((y.double() - want).abs().max())
    assert err <= 3e-6 * scale, (err, scale)
    assert err <= 1e-4 * float(want.abs().max() + 1e-6)


def test_gemm_ts_is_deterministic_and_rejects_bad_pitch():
    from rlpyt_b200 import _lib
    from rlpyt_b200.models.gemm_op import gemm_ts, split_lo, transpose_split
    g = torch.Generator(device="cuda").manual_seed(5)
    a = torch.randn(3200, 640, device="cuda", generator=g)
    b = torch.randn(512, 640, device="cuda", generator=g)
    bt, bt_lo = transpose_split(b.t().contiguous())       # back to [512, 640]
    assert torch.equal(bt, b) and torch.equal(bt_lo, split_lo(b))
    y1, y2 = gemm_ts(a, b, bt_lo), gemm_ts(a, b, bt_lo)
    assert torch.equal(y1, y2)
    with pytest.raises(_lib.B200LibraryError):                     # K % 4 != 0: the TMA row pitch is not a multiple of 16 bytes
        gemm_ts(a[:, :639].contiguous(), b[:, :639].contiguous(), bt_lo[:, :639].contiguous())


@pytest.mark.parametrize("impl", ["ts", "ss"])
def test_linear_autograd_both_kernels_vs_fp64(impl, monkeypatch):
    """forward / grad_input / grad_weight / grad_bias of the Linear op on each GEMM kernel against fp64 autograd."""
    from rlpyt_b200.models import gemm_op
    monkeypatch.setattr(gemm_op, "GEMM_IMPL", impl)
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.randn(2048, 3200, device="cuda", generator=g)
    w = torch.randn(512, 3200, device="cuda", generator=g) / 56
    b = torch.randn(512, device="cuda", generator=g)
    go = torch.randn(2048, 512, device="cuda", generator=g)
    x1, w1, b1 = (t.clone().requires_grad_(True) for t in (x, w, b))
    y = gemm_op.linear_tf32x3(x1, w1, b1, relu=True)
    (y * go).sum().backward()
    xd, wd, bd = (t.double().requires_grad_(True) for t in (x, w, b))
    pre = torch.nn.functional.linear(xd, wd, bd)
    mask = (y.detach() > 0).double()                      # the op's own ReLU mask (ties at 0 are measure-zero but not nil)
    (pre * mask * go.double()).sum().backward()
    yd = pre.detach() * mask
    assert float((y.double() - yd).abs().max()) <= 1e-5 * float(yd.abs().max())
    for got, want in ((x1.grad, xd.grad), (w1.grad, wd.grad), (b1.grad, bd.grad)):
        assert float((got.double() - want).abs().max()) <= 1e-5 * float(want.abs().max())


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


@pytest.mark.parametrize("N,plane", [(1, (20, 20)), (5, (20, 20)), (64, (20, 20)), (300, (25, 19)), (3, (8, 8)), (2, (7, 5))])
def test_conv2_forward_tcgen05_vs_torch(N, plane):
    """tcgen05 implicit-GEMM Conv2d(16->32,k4,s2,p1)+ReLU vs torch fp32 (1e-5 of the row scale)."""
    import torch.nn.functional as F
    from rlpyt_b200.models.conv2_op import conv2_relu
    g = torch.Generator(device="cuda").manual_seed(N)
    x = torch.relu(torch.randn((N, 16) + plane, device="cuda", generator=g))
    w = torch.randn(32, 16, 4, 4, device="cuda", generator=g) / 16
    b = torch.randn(32, device="cuda", generator=g) / 4
    y = conv2_relu(x, w, b)
    ref = F.relu(F.conv2d(x.double(), w.double(), b.double(), stride=2, padding=1))
    assert y.shape == ref.shape
    scale = float(F.conv2d(x.double().abs(), w.double().abs(), None, stride=2, padding=1).max())
    assert float((y.double() - ref).abs().max()) <= 3e-6 * scale
    x1, w1, b1 = (t.clone().requires_grad_(True) for t in (x, w, b))
    x2, w2, b2 = (t.clone().requires_grad_(True) for t in (x, w, b))
    go = torch.randn(y.shape, device="cuda", generator=g)
    y1 = conv2_relu(x1, w1, b1)
    y1.backward(go)
    pre = F.conv2d(x2, w2, b2, stride=2, padding=1)
    pre.backward(go * (y1.detach() > 0))
    for got, want in ((x1.grad, x2.grad), (w1.grad, w2.grad), (b1.grad, b2.grad)):
        s = float(want.abs().max()) + 1e-12
        np.testing.assert_allclose(got.cpu().numpy(), want.cpu().numpy(), rtol=1e-4, atol=2e-5 * s)


@pytest.mark.parametrize("impl", ["s2d", "tc"])
@pytest.mark.parametrize("N,plane", [(1, (20, 20)), (149, (20, 20)), (300, (25, 19)), (7, (8, 6)), (2, (7, 5)), (600, (20, 20))])
def test_conv2_forward_and_dgrad_vs_fp64(impl, N, plane):
    """Both implementations of the second layer through the C ABI - "s2d" (csrc/conv2_s2d.cuh: cell rows, row-shifted
    descriptors, bulk-copied images, input gradient assembled in shared memory) and "tc" (csrc/conv_tc.cu: im2col
    implicit GEMMs) - against fp64: forward <= 3e-6 of sum|x||w|, input gradient <= 3e-6 of sum|g||w|; every
    pixel of the input gradient is written (the buffer starts as NaN)."""
    import torch.nn.functional as F
    from rlpyt_b200 import _lib
    lib = _lib.load()
    IH, IW = plane
    if impl == "s2d" and not lib.rl_conv2_s2d_supported(16, IH, IW):
        pytest.skip("geometry not supported by the s2d kernels")
    g = torch.Generator(device="cuda").manual_seed(N + IH)
    x = torch.relu(torch.randn((N, 16) + plane, device="cuda", generator=g))
    w = torch.randn(32, 16, 4, 4, device="cuda", generator=g) / 16
    b = torch.randn(32, device="cuda", generator=g) / 4
    OH, OW = (IH - 2) // 2 + 1, (IW - 2) // 2 + 1
    y = torch.full((N, 32, OH, OW), float("nan"), device="cuda")
    _lib.call("rl_conv2_forward_" + impl, _lib.ptr(x), _lib.ptr(w), _lib.ptr(b), _lib.ptr(y), N, 16, IH, IW, 1, _lib.stream())
    ref = F.relu(F.conv2d(x.double(), w.double(), b.double(), stride=2, padding=1))
    scale = F.conv2d(x.double().abs(), w.double().abs(), None, stride=2, padding=1) + b.double().abs().view(1, 32, 1, 1)
    assert torch.isfinite(y).all()
    assert float(((y.double() - ref).abs() / scale).max()) <= 3e-6
    go = torch.randn(N, 32, OH, OW, device="cuda", generator=g) * (torch.rand(N, 32, OH, OW, device="cuda", generator=g) < 0.7)
    gx = torch.full((N, 16, IH, IW), float("nan"), device="cuda")
    if impl == "s2d":
        _lib.call("rl_conv2_dgrad_s2d", _lib.ptr(go), _lib.ptr(w), _lib.ptr(gx), N, 16, IH, IW, _lib.stream())
    else:
        sc = torch.empty(int(lib.rl_conv2_dgrad_tc_scratch_bytes()) // 4 + 4, device="cuda")
        _lib.call("rl_conv2_dgrad_tc", _lib.ptr(go), _lib.ptr(w), _lib.ptr(gx), N, 16, IH, IW, _lib.ptr(sc), _lib.stream(), n_launch=2)
    # fp64 reference on the HOST: cuDNN's double-precision transposed convolution takes about a minute per case here
    opad = (IH + 2 - 4) - 2 * (OH - 1), (IW + 2 - 4) - 2 * (OW - 1)
    go_h, w_h = go.cpu().double(), w.cpu().double()
    refx = F.conv_transpose2d(go_h, w_h, stride=2, padding=1, output_padding=opad)
    sx = F.conv_transpose2d(go_h.abs(), w_h.abs(), stride=2, padding=1, output_padding=opad)
    assert refx.shape == gx.shape and torch.isfinite(gx).all()
    assert float(((gx.cpu().double() - refx).abs() / sx.clamp_min(1e-30)).max()) <= 3e-6


@pytest.mark.parametrize("shape", [(4, 84, 84), (4, 36, 36), (4, 104, 80)])
@pytest.mark.parametrize("use_rows", [False, True])
def test_conv1_forward_tcgen05_vs_simt(shape, use_rows):
    """The tensor-core first layer (integer pixels exact in TF32, weights split hi/lo, 1/255 in the
    epilogue) against the fp32 SIMT kernel and an fp64 reference."""
    import torch.nn.functional as F
    from rlpyt_b200.models import conv1_op
    g = torch.Generator(device="cuda").manual_seed(9)
    obs = torch.randint(0, 256, (257,) + shape, dtype=torch.uint8, device="cuda", generator=g)
    rows = torch.randint(0, 257, (130,), device="cuda", generator=g) if use_rows else None
    w = (torch.rand(16, 4, 8, 8, device="cuda", generator=g) - 0.5) / 8
    b = (torch.rand(16, device="cuda", generator=g) - 0.5) / 8
    old = conv1_op.FORWARD_IMPL
    try:
        conv1_op.FORWARD_IMPL = "tc"
        y_tc = conv1_op.conv1_u8_relu(w, b, obs, rows)
        conv1_op.FORWARD_IMPL = "simt"
        y_simt = conv1_op.conv1_u8_relu(w, b, obs, rows)
    finally:
        conv1_op.FORWARD_IMPL = old
    x = (obs if rows is None else obs[rows]).double() / 255
    ref = F.relu(F.conv2d(x, w.double(), b.double(), stride=4))
    scale = float(F.conv2d(x.abs(), w.double().abs(), None, stride=4).max())
    assert float((y_tc.double() - ref).abs().max()) <= 3e-6 * scale
    np.testing.assert_allclose(y_tc.cpu().numpy(), y_simt.cpu().numpy(), rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize("shape,n_obs,n_rows", [((4, 84, 84), 70, None), ((4, 84, 84), 300, 130), ((4, 36, 36), 5, None),
                                                ((4, 104, 80), 33, 40), ((4, 8, 8), 3, None)])
def test_conv1_wgrad_tcgen05_vs_fp64(shape, n_obs, n_rows):
    """Weight/bias gradient of the uint8 first layer as a tcgen05 GEMM over positions (pixels exact in
    TF32, gradient split hi/lo, fp32 promotion every 4 k-blocks) against an fp64 reference with the
    same ReLU mask; 1e-5 of the term-magnitude scale."""
    import torch.nn.functional as F
    from rlpyt_b200 import _lib
    from rlpyt_b200.models.conv2_op import wgrad_scratch
    g = torch.Generator(device="cuda").manual_seed(n_obs)
    obs = torch.randint(0, 256, (n_obs,) + shape, dtype=torch.uint8, device="cuda", generator=g)
    rows = torch.randint(0, n_obs, (n_rows,), device="cuda", generator=g) if n_rows else None
    N = n_rows or n_obs
    OH, OW = (shape[1] - 8) // 4 + 1, (shape[2] - 8) // 4 + 1
    out = torch.randn(N, 16, OH, OW, device="cuda", generator=g)          # sign = ReLU mask
    go = torch.randn(N, 16, OH, OW, device="cuda", generator=g)
    gw = torch.full((16, 4, 8, 8), float("nan"), device="cuda")
    gb = torch.full((16,), float("nan"), device="cuda")
    _lib.call("rl_conv1_u8_wgrad_tc", _lib.ptr(obs), _lib.ptr(rows), _lib.ptr(out), _lib.ptr(go), _lib.ptr(gw),
              _lib.ptr(gb), N, 4, shape[1], shape[2], _lib.ptr(wgrad_scratch(obs.device)), _lib.stream(), n_launch=2)
    x = ((obs if rows is None else obs[rows]).double() / 255).requires_grad_(False)
    gm = (go * (out > 0)).double()
    w = torch.zeros(16, 4, 8, 8, dtype=torch.float64, device="cuda", requires_grad=True)
    b = torch.zeros(16, dtype=torch.float64, device="cuda", requires_grad=True)
    F.conv2d(x, w, b, stride=4).backward(gm)
    wa = torch.zeros_like(w, requires_grad=True)
    F.conv2d(x, wa, None, stride=4).backward(gm.abs())
    scale_w = float(wa.grad.max())
    assert float((gw.double() - w.grad).abs().max()) <= 1e-5 * scale_w
    assert float((gb.double() - b.grad).abs().max()) <= 1e-5 * float(gm.abs().sum((0, 2, 3)).max())


@pytest.mark.parametrize("shape", [(4, 84, 84), (4, 36, 36), (4, 104, 80), (4, 8, 8), (4, 64, 128)])
@pytest.mark.parametrize("n_obs,n_rows,relu", [(257, 130, 1), (3, None, 0), (400, None, 1)])
def test_conv1_forward_i8_vs_fp64(shape, n_obs, n_rows, relu):
    """The integer tensor-core first layer (csrc/conv1_i8.cuh: uint8 frames as exact kind::i8 operands, filter
    bank as four base-128 digits, exact int32 accumulation) against an fp64 convolution: <= 3e-6 of sum|x||w|
    (the bar of the TF32 kernel; measured ~5e-8), with and without the minibatch row gather, ragged last tiles,
    weights spanning six orders of magnitude within a channel."""
    import torch.nn.functional as F
    from rlpyt_b200 import _lib
    assert _lib.load().rl_conv1_u8_i8_supported(*shape)
    g = torch.Generator(device="cuda").manual_seed(11 + n_obs)
    obs = torch.randint(0, 256, (n_obs,) + shape, dtype=torch.uint8, device="cuda", generator=g)
    rows = torch.randint(0, n_obs, (n_rows,), device="cuda", generator=g) if n_rows else None
    N = n_rows or n_obs
    w = (torch.rand(16, 4, 8, 8, device="cuda", generator=g) - 0.5) / 8
    w.view(-1)[::7] *= 1e-4                                   # low digits
    w[3] *= 50.0                                              # per-channel scales
    w[5] = 0.0                                                # an all-zero channel (scale 2^0)
    b = (torch.rand(16, device="cuda", generator=g) - 0.5) / 8
    OH, OW = (shape[1] - 8) // 4 + 1, (shape[2] - 8) // 4 + 1
    y = torch.full((N, 16, OH, OW), float("nan"), device="cuda")
    _lib.call("rl_conv1_u8_forward_i8", _lib.ptr(obs), _lib.ptr(rows), _lib.ptr(w), _lib.ptr(b), _lib.ptr(y), N, 4,
              shape[1], shape[2], relu, _lib.stream())
    x = (obs if rows is None else obs[rows]).double() / 255
    ref = F.conv2d(x, w.double(), b.double(), stride=4)
    ref = F.relu(ref) if relu else ref
    scale = F.conv2d(x.abs(), w.double().abs(), None, stride=4) + b.double().abs().view(1, 16, 1, 1)
    assert torch.isfinite(y).all()
    assert float(((y.double() - ref).abs() / scale.clamp_min(1e-30)).max()) <= 3e-6


def test_conv1_i8_rejects_unsupported_geometry():
    from rlpyt_b200 import _lib
    lib = _lib.load()
    assert not lib.rl_conv1_u8_i8_supported(4, 210, 160) and not lib.rl_conv1_u8_i8_supported(4, 84, 82)
    obs = torch.zeros(2, 4, 84, 82, dtype=torch.uint8, device="cuda")
    w, b, y = torch.zeros(16, 4, 8, 8, device="cuda"), torch.zeros(16, device="cuda"), torch.zeros(2, 16, 20, 19, device="cuda")
    with pytest.raises(_lib.B200LibraryError):
        _lib.call("rl_conv1_u8_forward_i8", _lib.ptr(obs), None, _lib.ptr(w), _lib.ptr(b), _lib.ptr(y), 2, 4, 84, 82, 1,
                  _lib.stream())


@pytest.mark.parametrize("shape,n_obs,n_rows,mask", [((4, 84, 84), 70, None, True), ((4, 84, 84), 300, 130, True),
                                                     ((4, 36, 36), 5, None, False), ((4, 104, 80), 33, 40, True),
                                                     ((4, 8, 8), 3, None, True), ((4, 64, 128), 9, None, True), ((4, 84, 84), 1000, None, True)])
def test_conv1_wgrad_i8_vs_fp64(shape, n_obs, n_rows, mask):
    """Weight/bias gradient of the uint8 first layer on the integer tensor cores (gradient as four base-128 digits
    against a per-channel power-of-two scale, exact int32 accumulation over every frame of a CTA, fp64 reduction)
    against an fp64 reference with the same ReLU mask.  Heavy-tailed gradients (a few elements 300x the rest,
    one channel 1e-4 of the others): 1e-6 of the term-magnitude scale (the TF32 kernel's bar is 1e-5)."""
    import torch.nn.functional as F
    from rlpyt_b200 import _lib
    g = torch.Generator(device="cuda").manual_seed(n_obs)
    obs = torch.randint(0, 256, (n_obs,) + shape, dtype=torch.uint8, device="cuda", generator=g)
    rows = torch.randint(0, n_obs, (n_rows,), device="cuda", generator=g) if n_rows else None
    N = n_rows or n_obs
    OH, OW = (shape[1] - 8) // 4 + 1, (shape[2] - 8) // 4 + 1
    out = torch.randn(N, 16, OH, OW, device="cuda", generator=g)          # sign = ReLU mask
    go = torch.randn(N, 16, OH, OW, device="cuda", generator=g) * 1e-3
    go = torch.where(torch.rand(go.shape, device="cuda", generator=g) < 1 / 64, go * 300, go)
    go[:, 5] *= 1e-4
    go[:, 9] = 0.0
    gw = torch.full((16, 4, 8, 8), float("nan"), device="cuda")
    gb = torch.full((16,), float("nan"), device="cuda")
    sc = torch.empty(int(_lib.load().rl_conv1_u8_wgrad_i8_scratch_bytes()) // 4 + 4, device="cuda")
    _lib.call("rl_conv1_u8_wgrad_i8", _lib.ptr(obs), _lib.ptr(rows), _lib.ptr(out) if mask else None, _lib.ptr(go),
              _lib.ptr(gw), _lib.ptr(gb), N, 4, shape[1], shape[2], _lib.ptr(sc), _lib.stream(), n_launch=3)
    x = (obs if rows is None else obs[rows]).double() / 255
    gm = (go * (out > 0)).double() if mask else go.double()
    w = torch.zeros(16, 4, 8, 8, dtype=torch.float64, device="cuda", requires_grad=True)
    b = torch.zeros(16, dtype=torch.float64, device="cuda", requires_grad=True)
    F.conv2d(x, w, b, stride=4).backward(gm)
    wa = torch.zeros_like(w, requires_grad=True)
    F.conv2d(x, wa, None, stride=4).backward(gm.abs())
    assert torch.isfinite(gw).all() and torch.isfinite(gb).all()
    assert float(((gw.double() - w.grad).abs() / wa.grad.clamp_min(1e-300)).max()) <= 1e-6
    assert float((gb.double() - b.grad).abs().max()) <= 1e-5 * float(gm.abs().sum((0, 2, 3)).max())
    assert float(gw[9].abs().max()) == 0.0 and float(gb[9]) == 0.0
    # deterministic: a second call gives the same bits
    gw2, gb2 = torch.empty_like(gw), torch.empty_like(gb)
    _lib.call("rl_conv1_u8_wgrad_i8", _lib.ptr(obs), _lib.ptr(rows), _lib.ptr(out) if mask else None, _lib.ptr(go),
              _lib.ptr(gw2), _lib.ptr(gb2), N, 4, shape[1], shape[2], _lib.ptr(sc), _lib.stream(), n_launch=3)
    assert torch.equal(gw, gw2) and torch.equal(gb, gb2)


@pytest.mark.parametrize("N,plane", [(1, (20, 20)), (64, (20, 20)), (700, (20, 20)), (300, (25, 19)), (2, (7, 5))])
def test_conv2_wgrad_tcgen05_vs_fp64(N, plane):
    import torch.nn.functional as F
    from rlpyt_b200 import _lib
    from rlpyt_b200.models.conv2_op import wgrad_scratch
    g = torch.Generator(device="cuda").manual_seed(N)
    x = torch.relu(torch.randn((N, 16) + plane, device="cuda", generator=g))
    OH, OW = (plane[0] - 2) // 2 + 1, (plane[1] - 2) // 2 + 1
    out = torch.randn(N, 32, OH, OW, device="cuda", generator=g)
    go = torch.randn(N, 32, OH, OW, device="cuda", generator=g)
    for masked_in_kernel in (True, False):
        gw = torch.full((32, 16, 4, 4), float("nan"), device="cuda")
        gb = torch.full((32,), float("nan"), device="cuda")
        gin = go if masked_in_kernel else (go * (out > 0)).contiguous()
        _lib.call("rl_conv2_wgrad_tc", _lib.ptr(x), _lib.ptr(out) if masked_in_kernel else None, _lib.ptr(gin),
                  _lib.ptr(gw), _lib.ptr(gb), N, 16, plane[0], plane[1], _lib.ptr(wgrad_scratch(x.device)),
                  _lib.stream(), n_launch=2)
        gm = (go * (out > 0)).double()
        w = torch.zeros(32, 16, 4, 4, dtype=torch.float64, device="cuda", requires_grad=True)
        b = torch.zeros(32, dtype=torch.float64, device="cuda", requires_grad=True)
        F.conv2d(x.double(), w, b, stride=2, padding=1).backward(gm)
        wa = torch.zeros_like(w, requires_grad=True)
        F.conv2d(x.double(), wa, None, stride=2, padding=1).backward(gm.abs())
        assert float((gw.double() - w.grad).abs().max()) <= 1e-5 * float(wa.grad.max())
        assert float((gb.double() - b.grad).abs().max()) <= 1e-5 * float(gm.abs().sum((0, 2, 3)).max())


@pytest.mark.parametrize("N,plane", [(1, (20, 20)), (149, (20, 20)), (700, (20, 20)), (300, (25, 19)), (9, (8, 6))])
def test_conv2_wgrad_s2d_vs_fp64(N, plane):
    """The cell-space weight gradient (csrc/conv2_s2d.cuh wg2: both operands MN-major in the SWIZZLE_128B_BASE32B
    layout, K = cells, M = 64 accumulators per tap promoted to fp32 per image, fixed-order reduction) against fp64 on
    an already masked gradient: 1e-5 of the term-magnitude sum (measured ~5e-7), deterministic."""
    import torch.nn.functional as F
    from rlpyt_b200 import _lib
    lib = _lib.load()
    if not lib.rl_conv2_s2d_supported(16, plane[0], plane[1]):
        pytest.skip("geometry not supported by the s2d kernels")
    g = torch.Generator(device="cuda").manual_seed(N + 3)
    x = torch.relu(torch.randn((N, 16) + plane, device="cuda", generator=g))
    OH, OW = (plane[0] - 2) // 2 + 1, (plane[1] - 2) // 2 + 1
    gm = (torch.randn(N, 32, OH, OW, device="cuda", generator=g) * (torch.rand(N, 32, OH, OW, device="cuda", generator=g) < 0.7)).contiguous()
    sc = torch.empty(int(lib.rl_conv2_wgrad_s2d_scratch_bytes()) // 4 + 4, device="cuda")
    outs = []
    for _ in range(2):
        gw = torch.full((32, 16, 4, 4), float("nan"), device="cuda")
        gb = torch.full((32,), float("nan"), device="cuda")
        _lib.call("rl_conv2_wgrad_s2d", _lib.ptr(x), _lib.ptr(gm), _lib.ptr(gw), _lib.ptr(gb), N, 16, plane[0], plane[1], _lib.ptr(sc),
                  _lib.stream(), n_launch=2)
        outs.append((gw, gb))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    w = torch.zeros(32, 16, 4, 4, dtype=torch.float64, device="cuda", requires_grad=True)
    b = torch.zeros(32, dtype=torch.float64, device="cuda", requires_grad=True)
    F.conv2d(x.double(), w, b, stride=2, padding=1).backward(gm.double())
    wa = torch.zeros_like(w, requires_grad=True)
    F.conv2d(x.double(), wa, None, stride=2, padding=1).backward(gm.double().abs())
    gw, gb = outs[0]
    assert torch.isfinite(gw).all() and torch.isfinite(gb).all()
    assert float(((gw.double() - w.grad).abs() / wa.grad.clamp_min(1e-300)).max()) <= 1e-5
    assert float((gb.double() - b.grad).abs().max()) <= 1e-5 * float(gm.double().abs().sum((0, 2, 3)).max())


@pytest.mark.parametrize("shape", [(0,), (1,), (7,), (8192, 512), (33, 5, 7)])
def test_relu_backward_bit_exact(shape):
    from rlpyt_b200.models.gemm_op import relu_backward
    g = torch.Generator(device="cuda").manual_seed(3)
    grad = torch.randn(shape, device="cuda", generator=g)
    out = torch.relu(torch.randn(shape, device="cuda", generator=g))
    got = relu_backward(grad, out)
    assert torch.equal(got, grad * (out > 0))
    if grad.numel() > 8:      # 4-byte-aligned views take the scalar path
        flat_g, flat_o = grad.reshape(-1)[1:], out.reshape(-1)[1:]
        assert torch.equal(relu_backward(flat_g, flat_o), flat_g * (flat_o > 0))


@pytest.mark.parametrize("rows,cols", [(1, 1), (5, 3), (32, 32), (33, 65), (8192, 512), (512, 3200), (1000, 17)])
def test_transpose_bit_exact(rows, cols):
    from rlpyt_b200.models.gemm_op import transpose2d
    x = torch.randn(rows, cols, device="cuda", generator=torch.Generator(device="cuda").manual_seed(rows))
    got = transpose2d(x)
    assert got.shape == (cols, rows) and got.is_contiguous()
    assert torch.equal(got, x.t().contiguous())
