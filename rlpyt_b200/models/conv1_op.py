"""``Conv2d(4->16, k8, s4) + ReLU`` on uint8 frames as one autograd op over the hand-written kernels:
forward (fused with the minibatch row gather and the ``/255`` scaling) and backward (weight and bias
gradients with the ReLU mask folded in - the input is an observation and needs no gradient).
Default implementation "i8": the integer tensor-core kernels of csrc/conv1_i8.cuh (uint8 frames as exact
``tcgen05.mma kind::i8`` operands, the fp32 operand as four base-128 digits, exact int32 accumulation) whenever
the geometry fits (H, W multiples of 4, W <= 128); otherwise "tc", the TF32 implicit-GEMM kernels of
csrc/conv_tc.cu.  ``RLPYT_B200_CONV1_FWD`` / ``RLPYT_B200_CONV_WGRAD`` = i8 | tc | simt select one explicitly
(simt = the fp32 kernels of csrc/conv1.cu, kept for cross-checks)."""
import ctypes
import os

import torch

from rlpyt_b200 import _lib
from rlpyt_b200.algos.optim import grad_destination

_SCRATCH = {}
FORWARD_IMPL = os.environ.get("RLPYT_B200_CONV1_FWD", "i8")
WGRAD_IMPL = os.environ.get("RLPYT_B200_CONV_WGRAD", "i8")
_I8_OK = {}


def i8_supported(C, H, W):
    key = (C, H, W)
    if key not in _I8_OK:
        _I8_OK[key] = bool(_lib.load().rl_conv1_u8_i8_supported(C, H, W))
    return _I8_OK[key]


def _impl(name, C, H, W, N=0):
    if name == "i8" and not (i8_supported(C, H, W) and N <= 256 * 148):
        return "tc"
    return name


def _scratch(dev, key, nbytes_fn):
    k = (str(dev), key)
    if k not in _SCRATCH:
        _SCRATCH[k] = torch.empty(int(nbytes_fn()) // 4 + 4, dtype=torch.float32, device=dev)
    return _SCRATCH[k]


def supported(image_shape, conv_layers):
    """True when the model's first layers are exactly the reference default the kernel implements
    (rlpyt/models/pg/atari_ff_model.py:31-35: channels[0]=16, kernel 8, stride 4, padding 0, ReLU)."""
    c, h, w = image_shape
    first = conv_layers[0] if len(conv_layers) else None
    return (isinstance(first, torch.nn.Conv2d) and c == 4 and w % 4 == 0 and (c * h * w) % 16 == 0
            and first.out_channels == 16 and tuple(first.kernel_size) == (8, 8) and tuple(first.stride) == (4, 4)
            and tuple(first.padding) == (0, 0) and tuple(first.dilation) == (1, 1) and first.groups == 1
            and len(conv_layers) > 1 and isinstance(conv_layers[1], torch.nn.ReLU))


def _producer_absmax(g):
    """max |g[:, c]| per channel if ``g`` is the very tensor conv2's input-gradient kernel just wrote (it leaves the
    maxima of its result in ``conv2_op.LAST_DGRAD_ABSMAX``): same live object, same memory.  Anything else - an
    accumulated or copied gradient, another producer - gets ``None`` and the kernel's own absmax pass."""
    from rlpyt_b200.models import conv2_op
    rec = conv2_op.LAST_DGRAD_ABSMAX.pop(str(g.device), None)
    if rec is None:
        return None
    src = rec["ref"]()
    if src is None or src.data_ptr() != g.data_ptr() or rec["ptr"] != g.data_ptr() or src.shape != g.shape or src._version != g._version:
        return None
    return rec["absmax"]


class Conv1U8Relu(torch.autograd.Function):

    @staticmethod
    def forward(ctx, weight, bias, obs, rows):
        """obs: [R,4,H,W] uint8 CUDA contiguous; rows: int64 [N] or None -> [N,16,OH,OW] fp32."""
        _lib.require_cuda(weight, bias, obs, rows)
        R, C, H, W = obs.shape
        N = R if rows is None else rows.numel()
        OH, OW = (H - 8) // 4 + 1, (W - 8) // 4 + 1
        w, b = weight.detach().contiguous(), bias.detach().contiguous()
        out = torch.empty((N, 16, OH, OW), dtype=torch.float32, device=obs.device)
        fn = {"i8": "rl_conv1_u8_forward_i8", "tc": "rl_conv1_u8_forward_tc", "simt": "rl_conv1_u8_forward"}[
            _impl(FORWARD_IMPL, C, H, W)]
        with torch.cuda.device(obs.device):
            _lib.call(fn, _lib.ptr(obs), _lib.ptr(rows), _lib.ptr(w), _lib.ptr(b), _lib.ptr(out), N, C, H, W, 1, _lib.stream())
        ctx.obs, ctx.rows = obs, rows
        ctx.params = (weight, bias)                          # for grad_destination (the flat gradient buffer's slots)
        ctx.save_for_backward(out)
        ctx.mark_non_differentiable()
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (out,) = ctx.saved_tensors
        obs, rows = ctx.obs, ctx.rows
        R, C, H, W = obs.shape
        N = out.shape[0]
        dev = obs.device
        impl = _impl(WGRAD_IMPL, C, H, W, N)
        if impl == "i8":
            gw, gb = grad_destination(ctx.params[0]), grad_destination(ctx.params[1])
            g = grad_out.contiguous()
            sc = _scratch(dev, "wgrad_i8", _lib.load().rl_conv1_u8_wgrad_i8_scratch_bytes)
            absmax = _producer_absmax(g)
            with torch.cuda.device(dev):
                if absmax is not None:      # the per-channel bound came with the gradient (conv2's input-gradient epilogue)
                    _lib.call("rl_conv1_u8_wgrad_i8_scaled", _lib.ptr(obs), _lib.ptr(rows), _lib.ptr(out), _lib.ptr(g), _lib.ptr(absmax),
                              _lib.ptr(gw), _lib.ptr(gb), N, C, H, W, _lib.ptr(sc), _lib.stream(), n_launch=2)
                else:
                    _lib.call("rl_conv1_u8_wgrad_i8", _lib.ptr(obs), _lib.ptr(rows), _lib.ptr(out), _lib.ptr(g), _lib.ptr(gw),
                              _lib.ptr(gb), N, C, H, W, _lib.ptr(sc), _lib.stream(), n_launch=3)
            return gw, gb, None, None
        if impl == "tc":
            from rlpyt_b200.models.conv2_op import wgrad_scratch
            gw, gb = grad_destination(ctx.params[0]), grad_destination(ctx.params[1])
            g = grad_out.contiguous()
            with torch.cuda.device(dev):
                _lib.call("rl_conv1_u8_wgrad_tc", _lib.ptr(obs), _lib.ptr(rows), _lib.ptr(out), _lib.ptr(g),
                          _lib.ptr(gw), _lib.ptr(gb), N, C, H, W, _lib.ptr(wgrad_scratch(dev)), _lib.stream(),
                          n_launch=2)
            return gw, gb, None, None
        key = str(dev)
        scratch = _SCRATCH.get(key)
        if scratch is None:
            nbytes = int(_lib.load().rl_conv1_u8_wgrad_scratch_bytes())
            scratch = torch.empty(nbytes // 4, dtype=torch.float32, device=dev)
            _SCRATCH[key] = scratch
        gw, gb = grad_destination(ctx.params[0]), grad_destination(ctx.params[1])
        g = grad_out.contiguous()
        with torch.cuda.device(dev):
            _lib.call("rl_conv1_u8_wgrad", _lib.ptr(obs), _lib.ptr(rows), _lib.ptr(out), _lib.ptr(g), _lib.ptr(gw),
                      _lib.ptr(gb), N, C, H, W, 1, _lib.ptr(scratch), _lib.stream(), n_launch=2)
        return gw, gb, None, None


def conv1_u8_relu(weight, bias, obs, rows=None):
    return Conv1U8Relu.apply(weight, bias, obs, rows)


def stream_supported(C, H, W):
    """Can ``conv1_u8_relu_stream`` take this geometry (the kind::i8 kernel must be the dispatched forward)?"""
    return _impl(FORWARD_IMPL, C, H, W) == "i8"


@torch.no_grad()
def conv1_u8_relu_stream(weight, bias, frames):
    """Forward only (``agent.step``): ``frames`` = ``HostMappedFrames``; the kernel reads them from page-locked host memory
    and records them in ``frames.copy_to`` while computing the layer."""
    N, C, H, W = frames.shape
    OH, OW = (H - 8) // 4 + 1, (W - 8) // 4 + 1
    dev = frames.copy_to.device
    w, b = weight.detach().contiguous(), bias.detach().contiguous()
    out = torch.empty((N, 16, OH, OW), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.call("rl_conv1_u8_forward_i8_stream", ctypes.c_void_p(frames.host_ptr), _lib.ptr(w), _lib.ptr(b), _lib.ptr(out),
                  _lib.ptr(frames.copy_to), N, C, H, W, 1, _lib.stream())
    return out
