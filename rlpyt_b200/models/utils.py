"""Shape / state-dict / gradient helpers for the models (mirror of ``rlpyt/models/utils.py``)."""
import torch


def conv2d_output_shape(h, w, kernel_size=1, stride=1, padding=0, dilation=1):
    """Output (h, w) of a Conv2d/MaxPool2d layer (rlpyt/models/utils.py:5-15)."""
    def pair(x):
        return x if isinstance(x, (tuple, list)) else (x, x)
    k, s, p, d = pair(kernel_size), pair(stride), pair(padding), pair(dilation)
    oh = (h + 2 * p[0] - d[0] * (k[0] - 1) - 1) // s[0] + 1
    ow = (w + 2 * p[1] - d[1] * (k[1] - 1) - 1) // s[1] + 1
    return oh, ow


def strip_ddp_state_dict(state_dict):
    """Drop the ``module.`` prefix DistributedDataParallel adds (rlpyt/models/utils.py:57-70)."""
    return type(state_dict)((k[7:] if k.startswith("module.") else k, v) for k, v in state_dict.items())


def update_state_dict(model, state_dict, tau=1, strip_ddp=True):
    """Hard (``tau == 1``) or soft ``tau * new + (1 - tau) * old`` update of ``model``'s state
    (rlpyt/models/utils.py:42-54)."""
    if strip_ddp:
        state_dict = strip_ddp_state_dict(state_dict)
    if tau == 1:
        model.load_state_dict(state_dict)
    elif tau > 0:
        model.load_state_dict({k: tau * state_dict[k] + (1 - tau) * v for k, v in model.state_dict().items()})


class _ScaleGrad(torch.autograd.Function):

    @staticmethod
    def forward(ctx, x, scale):
        ctx.scale = scale
        return x

    @staticmethod
    def backward(ctx, grad_output):
        return grad_output * ctx.scale, None


def scale_grad(tensor, scale):
    """Identity in the forward pass, gradient multiplied by ``scale`` (rlpyt/models/utils.py:18-39)."""
    return _ScaleGrad.apply(tensor, scale)
