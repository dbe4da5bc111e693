"""``Conv2d(16->32, k4, s2, p1) + ReLU`` with the forward on the tensor cores (tcgen05 implicit GEMM,
csrc/conv_tc.cu).  The backward (input and weight gradients) stays on cuDNN's fp32 kernels this
round - see DESIGN.md section 6."""
import torch

from rlpyt_b200 import _lib


def supported(layer, act):
    return (isinstance(layer, torch.nn.Conv2d) and isinstance(act, torch.nn.ReLU) and layer.in_channels == 16
            and layer.out_channels == 32 and tuple(layer.kernel_size) == (4, 4) and tuple(layer.stride) == (2, 2)
            and tuple(layer.padding) == (1, 1) and tuple(layer.dilation) == (1, 1) and layer.groups == 1)


class Conv2ReluTC(torch.autograd.Function):

    @staticmethod
    def forward(ctx, x, weight, bias):
        _lib.require_cuda(x, weight, bias)
        x = x.contiguous()
        N, C, IH, IW = x.shape
        OH, OW = (IH - 2) // 2 + 1, (IW - 2) // 2 + 1
        out = torch.empty((N, 32, OH, OW), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.call("rl_conv2_forward_tc", _lib.ptr(x.detach()), _lib.ptr(weight.detach().contiguous()),
                      _lib.ptr(bias.detach().contiguous()), _lib.ptr(out), N, C, IH, IW, 1, _lib.stream())
        ctx.save_for_backward(x, weight, out)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        x, weight, out = ctx.saved_tensors
        g = grad_out * (out > 0)
        gx, gw, gb = torch.ops.aten.convolution_backward(
            g, x, weight, [32], [2, 2], [1, 1], [1, 1], False, [0, 0], 1,
            [ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]])
        return gx, gw, gb


def conv2_relu(x, weight, bias):
    return Conv2ReluTC.apply(x, weight, bias)
