"""``torch.nn.Linear`` (+ReLU) over the fp32-accurate tensor-core GEMM of csrc/gemm_tf32x3.cu
(tcgen05 kind::tf32, 3-term hi/lo split, TMA-staged, TMEM accumulators): forward, input gradient
and weight gradient all run on the same "TN" kernel (C = A B^T, K contiguous in both operands);
the backward operands are re-laid-out with the tiled transpose of csrc/layers.cu (HBM-bound, small
next to the GEMM) and the ReLU backward is one fused pass."""
import torch

from rlpyt_b200 import _lib

_WS = {}  # split-K workspaces, keyed by (device, bytes)


def gemm_tn(a, b, bias=None, relu=False):
    """a [M,K] @ b[N,K]^T (+bias) (+relu) -> [M,N]; fp32 CUDA, K % 4 == 0."""
    _lib.require_cuda(a, b, bias)
    a, b = a.contiguous(), b.contiguous()
    M, K = a.shape
    N = b.shape[0]
    assert b.shape[1] == K and a.dtype == torch.float32 and b.dtype == torch.float32
    out = torch.empty((M, N), dtype=torch.float32, device=a.device)
    ws_bytes = int(_lib.load().rl_gemm_tf32x3_workspace_bytes(M, N, K))
    ws = None
    if ws_bytes:
        key = (str(a.device), ws_bytes)
        ws = _WS.get(key)
        if ws is None:
            ws = torch.empty(ws_bytes // 4, dtype=torch.float32, device=a.device)
            _WS[key] = ws
    with torch.cuda.device(a.device):
        _lib.call("rl_gemm_tf32x3_f32", _lib.ptr(a), _lib.ptr(b), _lib.ptr(bias), _lib.ptr(out), M, N, K,
                  int(bool(relu)), _lib.ptr(ws), _lib.stream(), n_launch=2 if ws is not None else 1)
    return out


def relu_backward(grad, out):
    """grad * (out > 0) in one pass (csrc/layers.cu)."""
    _lib.require_cuda(grad, out)
    grad, out = grad.contiguous(), out.contiguous()
    dst = torch.empty_like(grad)
    with torch.cuda.device(grad.device):
        _lib.call("rl_relu_backward_f32", _lib.ptr(grad), _lib.ptr(out), _lib.ptr(dst), grad.numel(), _lib.stream())
    return dst


def transpose2d(x):
    """Contiguous transpose of a 2-D fp32 CUDA tensor."""
    _lib.require_cuda(x)
    x = x.contiguous()
    rows, cols = x.shape
    dst = torch.empty((cols, rows), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.call("rl_transpose_f32", _lib.ptr(x), _lib.ptr(dst), rows, cols, _lib.stream())
    return dst


def usable(in_features, out_features):
    return in_features % 4 == 0 and out_features % 4 == 0


class LinearTf32x3(torch.autograd.Function):

    @staticmethod
    def forward(ctx, x, weight, bias, relu):
        y = gemm_tn(x.detach(), weight.detach(), None if bias is None else bias.detach().contiguous(), relu)
        ctx.relu = relu
        ctx.has_bias = bias is not None
        ctx.save_for_backward(x, weight, y if relu else None)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, weight, y = ctx.saved_tensors
        gy = gy.contiguous()
        if ctx.relu:
            gy = relu_backward(gy, y)
        gx = gw = gb = None
        M = x.shape[0]
        if ctx.needs_input_grad[0]:
            gx = gemm_tn(gy, transpose2d(weight.detach()))            # [M,N] x [K,N]^T
        if ctx.needs_input_grad[1]:
            if M % 4 == 0:
                gw = gemm_tn(transpose2d(gy), transpose2d(x.detach()))   # [N,M] x [K,M]^T
            else:
                gw = gy.t().mm(x.detach())
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = gy.sum(0)
        return gx, gw, gb, None


def linear_tf32x3(x, weight, bias=None, relu=False):
    return LinearTf32x3.apply(x, weight, bias, relu)
