"""``torch.nn.Linear`` (+ReLU) over the fp32-accurate tensor-core GEMMs (tcgen05 kind::tf32, 3-term hi/lo
split, TMA-staged, TMEM accumulators): forward, input gradient and weight gradient.

Two kernels.  ``csrc/gemm_ts.cuh`` (default at M >= TS_MIN_M): the large operand (activations / output
gradient) goes through tensor memory, the small one (weights / transposed gradient) arrives pre-split
(``split_lo`` / ``transpose_split``), the activations of the weight gradient are read as they lie (no
transpose), persistent CTAs.  ``csrc/gemm_tf32x3.cu`` (first kernel, "TN" only, both operands in shared
memory): small M (``agent.step``) and ``RLPYT_B200_GEMM_IMPL=ss``.  The ReLU backward is one fused pass."""
import os

import torch

from rlpyt_b200 import _lib
from rlpyt_b200.algos.optim import grad_destination

_WS = {}  # split-K workspaces, keyed by (device, bytes)
GEMM_IMPL = os.environ.get("RLPYT_B200_GEMM_IMPL", "ts")
TS_MIN_M = int(os.environ.get("RLPYT_B200_GEMM_TS_MIN_M", "1024"))


def _workspace(device, ws_bytes):
    if not ws_bytes:
        return None
    key = (str(device), ws_bytes)
    ws = _WS.get(key)
    if ws is None:
        ws = torch.empty(ws_bytes // 4, dtype=torch.float32, device=device)
        _WS[key] = ws
    return ws


def split_lo(b):
    """b - trunc_tf32(b): the second term of the pre-split B operand of ``gemm_ts``."""
    _lib.require_cuda(b)
    b = b.contiguous()
    lo = torch.empty_like(b)
    with torch.cuda.device(b.device):
        _lib.call("rl_split_lo_f32", _lib.ptr(b), _lib.ptr(lo), b.numel(), _lib.stream())
    return lo


def transpose_split(x):
    """(x^T, x^T - trunc_tf32(x^T)) of a 2-D fp32 CUDA tensor, one pass."""
    _lib.require_cuda(x)
    x = x.contiguous()
    rows, cols = x.shape
    dst = torch.empty((cols, rows), dtype=torch.float32, device=x.device)
    lo = torch.empty_like(dst)
    with torch.cuda.device(x.device):
        _lib.call("rl_transpose_split_f32", _lib.ptr(x), _lib.ptr(dst), _lib.ptr(lo), rows, cols, _lib.stream())
    return dst, lo


def gemm_ts(a, b, b_lo, bias=None, relu=False, a_mmajor=False, c_trans=False, out_mask=None, out=None):
    """a @ b^T (+bias) (+relu).  a: [M,K], or with ``a_mmajor`` the [K,M] matrix a^T; b, b_lo: [N,K];
    result [M,N], or with ``c_trans`` its transpose [N,M].  ``out_mask`` [M,N]: result kept where ``out_mask > 0``, zero
    elsewhere (a preceding ReLU's backward folded into the epilogue; plain form only)."""
    _lib.require_cuda(a, b, b_lo, bias, out_mask)
    a, b, b_lo = a.contiguous(), b.contiguous(), b_lo.contiguous()
    (K, M) = a.shape if a_mmajor else a.shape[::-1]
    N = b.shape[0]
    assert b.shape == (N, K) and b_lo.shape == (N, K) and a.dtype == b.dtype == b_lo.dtype == torch.float32
    if out is None:
        out = torch.empty((N, M) if c_trans else (M, N), dtype=torch.float32, device=a.device)
    else:
        assert tuple(out.shape) == ((N, M) if c_trans else (M, N)) and out.is_contiguous() and out.dtype == torch.float32
    ws = _workspace(a.device, int(_lib.load().rl_gemm_ts_workspace_bytes(M, N, K)))
    if out_mask is not None:
        assert not (a_mmajor or c_trans or relu) and bias is None and tuple(out_mask.shape) == (M, N) and out_mask.is_contiguous()
        with torch.cuda.device(a.device):
            _lib.call("rl_gemm_ts_masked_f32", _lib.ptr(a), _lib.ptr(b), _lib.ptr(b_lo), _lib.ptr(out_mask), _lib.ptr(out), M, N, K,
                      _lib.ptr(ws), _lib.stream(), n_launch=2 if ws is not None else 1)
        return out
    with torch.cuda.device(a.device):
        _lib.call("rl_gemm_ts_f32", _lib.ptr(a), int(bool(a_mmajor)), _lib.ptr(b), _lib.ptr(b_lo), _lib.ptr(bias), _lib.ptr(out),
                  int(bool(c_trans)), M, N, K, int(bool(relu)), _lib.ptr(ws), _lib.stream(), n_launch=2 if ws is not None else 1)
    return out


def gemm_tn(a, b, bias=None, relu=False, out=None):
    """a [M,K] @ b[N,K]^T (+bias) (+relu) -> [M,N]; fp32 CUDA, K % 4 == 0."""
    _lib.require_cuda(a, b, bias)
    a, b = a.contiguous(), b.contiguous()
    M, K = a.shape
    N = b.shape[0]
    assert b.shape[1] == K and a.dtype == torch.float32 and b.dtype == torch.float32
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=a.device)
    else:
        assert tuple(out.shape) == (M, N) and out.is_contiguous() and out.dtype == torch.float32
    ws = _workspace(a.device, int(_lib.load().rl_gemm_tf32x3_workspace_bytes(M, N, K)))
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


def _use_ts(m):
    return GEMM_IMPL == "ts" and m >= TS_MIN_M


class LinearTf32x3(torch.autograd.Function):

    @staticmethod
    def forward(ctx, x, weight, bias, relu, input_is_relu_output=False):
        """``input_is_relu_output``: ``x`` is the output of a ReLU and the caller wants that ReLU's backward applied HERE
        (grad_x zeroed where x <= 0, in the input-gradient GEMM's epilogue) - the producer of ``x`` must then not apply
        it again.  Only honoured on the ``gemm_ts`` path (``fuses_input_relu``)."""
        b = None if bias is None else bias.detach().contiguous()
        ctx.input_relu = bool(input_is_relu_output)
        assert not ctx.input_relu or _use_ts(x.shape[0]), "input ReLU fusion needs the gemm_ts path"
        if _use_ts(x.shape[0]):
            w = weight.detach()
            y = gemm_ts(x.detach(), w, split_lo(w), b, relu)
        else:
            y = gemm_tn(x.detach(), weight.detach(), b, relu)
        ctx.relu = relu
        ctx.has_bias = bias is not None
        ctx.params = (weight, bias)                          # for grad_destination (the flat gradient buffer's slots)
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
        ts = _use_ts(M)
        if ctx.needs_input_grad[0]:
            if ts:
                gx = gemm_ts(gy, *transpose_split(weight.detach()),       # [M,N] x [K,N]^T
                             out_mask=x.detach() if ctx.input_relu else None)
            else:
                gx = gemm_tn(gy, transpose2d(weight.detach()))
        if ctx.needs_input_grad[1]:
            if M % 4 != 0:
                gw = gy.t().mm(x.detach())
            elif ts:
                # gw^T [K,N] = x^T [K,M] gy [M,N]: x is read as it lies (the [M,K] matrix is x^T's "M-major" form),
                # the result is stored transposed, only the small operand gy is transposed (+ split)
                gw = gemm_ts(x.detach(), *transpose_split(gy), a_mmajor=True, c_trans=True, out=grad_destination(ctx.params[0]))
            else:
                gw = gemm_tn(transpose2d(gy), transpose2d(x.detach()), out=grad_destination(ctx.params[0]))   # [N,M] x [K,M]^T
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = torch.sum(gy, 0, out=grad_destination(ctx.params[1]))
        return gx, gw, gb, None, None


def fuses_input_relu(m):
    """Will ``linear_tf32x3(x[m, :], ..., input_is_relu_output=True)`` apply the input's ReLU backward itself?"""
    return _use_ts(m)


def linear_tf32x3(x, weight, bias=None, relu=False, input_is_relu_output=False):
    return LinearTf32x3.apply(x, weight, bias, relu, input_is_relu_output)
