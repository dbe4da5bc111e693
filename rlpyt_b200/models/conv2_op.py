"""``Conv2d(16->32, k4, s2, p1) + ReLU`` on the tensor cores.  Forward and input gradient: the "s2d" kernels of
csrc/conv2_s2d.cuh (space-to-depth cell rows, row-shifted tcgen05 descriptors, bulk-copied images) whenever the
geometry fits, else the im2col implicit GEMMs of csrc/conv_tc.cu (``RLPYT_B200_CONV2=tc`` forces those).  Weight
gradient: tcgen05 GEMM over all positions (csrc/conv_tc.cu; ``RLPYT_B200_CONV_WGRAD=cudnn`` selects cuDNN's fp32
kernel for comparison)."""
import os

import torch

from rlpyt_b200 import _lib
from rlpyt_b200.algos.optim import grad_destination
from rlpyt_b200.models.gemm_op import relu_backward

_SCRATCH = {}
# max |grad_x[:, c]| per channel of the last input gradient this op produced: {"ref": weakref to that gradient tensor,
# "absmax": 16 device floats}.  The first layer's kind::i8 weight gradient needs exactly this bound for its gradient
# operand and finds it here when the tensor it receives IS that gradient (same object alive, same memory) - else it runs
# its own pass (models/conv1_op.py).
LAST_DGRAD_ABSMAX = {}
FUSE_ABSMAX = os.environ.get("RLPYT_B200_FUSE_ABSMAX", "1") == "1"
WGRAD_IMPL = os.environ.get("RLPYT_B200_CONV_WGRAD", "tc")
CONV2_IMPL = os.environ.get("RLPYT_B200_CONV2", "s2d")
_S2D_OK = {}


def s2d_supported(C, IH, IW):
    key = (C, IH, IW)
    if key not in _S2D_OK:
        _S2D_OK[key] = bool(_lib.load().rl_conv2_s2d_supported(C, IH, IW))
    return CONV2_IMPL == "s2d" and _S2D_OK[key]


def wgrad_scratch(dev):
    key = ("wgrad", str(dev))
    scratch = _SCRATCH.get(key)
    if scratch is None:
        scratch = torch.empty(int(_lib.load().rl_conv_wgrad_tc_scratch_bytes()) // 4, dtype=torch.float32, device=dev)
        _SCRATCH[key] = scratch
    return scratch


def supported(layer, act):
    return (isinstance(layer, torch.nn.Conv2d) and isinstance(act, torch.nn.ReLU) and layer.in_channels == 16
            and layer.out_channels == 32 and tuple(layer.kernel_size) == (4, 4) and tuple(layer.stride) == (2, 2)
            and tuple(layer.padding) == (1, 1) and tuple(layer.dilation) == (1, 1) and layer.groups == 1)


class Conv2ReluTC(torch.autograd.Function):

    @staticmethod
    def forward(ctx, x, weight, bias, grad_is_masked=False):
        """``grad_is_masked``: the consumer of the output applies this layer's ReLU backward itself (the fc layer's
        input-gradient GEMM zeroes its result where this output is <= 0): ``backward`` must not do it a second time."""
        _lib.require_cuda(x, weight, bias)
        ctx.grad_is_masked = bool(grad_is_masked)
        ctx.params = (weight, bias)
        x = x.contiguous()
        N, C, IH, IW = x.shape
        OH, OW = (IH - 2) // 2 + 1, (IW - 2) // 2 + 1
        out = torch.empty((N, 32, OH, OW), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.call("rl_conv2_forward_s2d" if s2d_supported(C, IH, IW) else "rl_conv2_forward_tc", _lib.ptr(x.detach()),
                      _lib.ptr(weight.detach().contiguous()), _lib.ptr(bias.detach().contiguous()), _lib.ptr(out),
                      N, C, IH, IW, 1, _lib.stream())
        ctx.save_for_backward(x, weight, out)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        x, weight, out = ctx.saved_tensors
        g = grad_out.contiguous() if ctx.grad_is_masked else relu_backward(grad_out, out)
        gx = gw = gb = None
        N, C, IH, IW = x.shape
        if ctx.needs_input_grad[0] and s2d_supported(C, IH, IW):
            gx = torch.empty_like(x)
            with torch.cuda.device(x.device):
                if FUSE_ABSMAX:
                    import weakref
                    absmax = torch.empty(16, dtype=torch.float32, device=x.device)
                    _lib.call("rl_conv2_dgrad_s2d_absmax", _lib.ptr(g), _lib.ptr(weight.detach().contiguous()), _lib.ptr(gx),
                              _lib.ptr(absmax), N, C, IH, IW, _lib.stream())
                    LAST_DGRAD_ABSMAX[str(x.device)] = dict(ref=weakref.ref(gx), ptr=gx.data_ptr(), absmax=absmax)
                else:
                    _lib.call("rl_conv2_dgrad_s2d", _lib.ptr(g), _lib.ptr(weight.detach().contiguous()), _lib.ptr(gx), N, C, IH, IW,
                              _lib.stream())
        elif ctx.needs_input_grad[0]:
            dev = x.device
            scratch = _SCRATCH.get(str(dev))
            if scratch is None:
                scratch = torch.empty(int(_lib.load().rl_conv2_dgrad_tc_scratch_bytes()) // 4, dtype=torch.float32,
                                      device=dev)
                _SCRATCH[str(dev)] = scratch
            gx = torch.empty_like(x)
            with torch.cuda.device(dev):
                _lib.call("rl_conv2_dgrad_tc", _lib.ptr(g), _lib.ptr(weight.detach().contiguous()), _lib.ptr(gx),
                          N, C, IH, IW, _lib.ptr(scratch), _lib.stream(), n_launch=2)
        if ((ctx.needs_input_grad[1] or ctx.needs_input_grad[2]) and WGRAD_IMPL == "tc" and s2d_supported(C, IH, IW)
                and ((IH + 2 - 4) // 2 + 2) * ((IW + 2 - 4) // 2 + 2) <= 128):
            # one 128-cell tile per image (20x20 planes): both operands MN-major, K = cells (csrc/conv2_s2d.cuh wg2)
            gw, gb = grad_destination(ctx.params[0]), grad_destination(ctx.params[1])
            sc = _SCRATCH.get(("wgrad_s2d", str(x.device)))
            if sc is None:
                sc = torch.empty(int(_lib.load().rl_conv2_wgrad_s2d_scratch_bytes()) // 4 + 4, dtype=torch.float32, device=x.device)
                _SCRATCH[("wgrad_s2d", str(x.device))] = sc
            with torch.cuda.device(x.device):
                _lib.call("rl_conv2_wgrad_s2d", _lib.ptr(x), _lib.ptr(g), _lib.ptr(gw), _lib.ptr(gb), N, C, IH, IW, _lib.ptr(sc),
                          _lib.stream(), n_launch=2)
        elif (ctx.needs_input_grad[1] or ctx.needs_input_grad[2]) and WGRAD_IMPL == "tc":
            gw, gb = grad_destination(ctx.params[0]), grad_destination(ctx.params[1])
            with torch.cuda.device(x.device):
                _lib.call("rl_conv2_wgrad_tc", _lib.ptr(x), None, _lib.ptr(g), _lib.ptr(gw), _lib.ptr(gb),
                          N, C, IH, IW, _lib.ptr(wgrad_scratch(x.device)), _lib.stream(), n_launch=2)
        elif ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
            _gx, gw, gb = torch.ops.aten.convolution_backward(
                g, x, weight, [32], [2, 2], [1, 1], [1, 1], False, [0, 0], 1,
                [False, ctx.needs_input_grad[1], ctx.needs_input_grad[2]])
        return gx, gw, gb, None


def conv2_relu(x, weight, bias, grad_is_masked=False):
    return Conv2ReluTC.apply(x, weight, bias, grad_is_masked)
