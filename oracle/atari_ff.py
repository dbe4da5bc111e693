"""Oracle: the AtariFf actor-critic network as a pure function of a ``state_dict`` (torch-CPU fp32).

Test infrastructure only (see oracle/__init__.py).  Follows rlpyt/models/pg/atari_ff_model.py:40-63
(u8 -> f32 * (1/255), conv stack, fc, softmax pi, value squeeze), rlpyt/models/conv2d.py:36-44
(Conv2d + ReLU per layer; defaults channels [16,32], kernels [8,4], strides [4,2], paddings [0,1]
from atari_ff_model.py:31-35) and rlpyt/models/mlp.py:30-36 (Linear + ReLU head, fc 512).
State-dict keys are the reference module names: conv.conv.conv.{0,2}.{weight,bias},
conv.head.model.0.{weight,bias}, pi.{weight,bias}, value.{weight,bias}.
"""
import torch
import torch.nn.functional as F

STRIDES = (4, 2)
PADDINGS = (0, 1)


def init_state_dict(image_shape, n_actions, seed=0, fc=512):
    """Random-init weights with the reference's layer shapes (torch default init ranges)."""
    g = torch.Generator().manual_seed(seed)
    c, h, w = image_shape

    def uni(shape, fan_in):
        bound = 1.0 / fan_in ** 0.5
        return (torch.rand(shape, generator=g) * 2 - 1) * bound

    h1, w1 = (h - 8) // 4 + 1, (w - 8) // 4 + 1
    h2, w2 = (h1 + 2 - 4) // 2 + 1, (w1 + 2 - 4) // 2 + 1
    flat = 32 * h2 * w2
    sd = {
        "conv.conv.conv.0.weight": uni((16, c, 8, 8), c * 64), "conv.conv.conv.0.bias": uni((16,), c * 64),
        "conv.conv.conv.2.weight": uni((32, 16, 4, 4), 256), "conv.conv.conv.2.bias": uni((32,), 256),
        "conv.head.model.0.weight": uni((fc, flat), flat), "conv.head.model.0.bias": uni((fc,), flat),
        "pi.weight": uni((n_actions, fc), fc), "pi.bias": uni((n_actions,), fc),
        "value.weight": uni((1, fc), fc), "value.bias": uni((1,), fc),
    }
    return sd


def forward(sd, image):
    """image: [N,C,H,W] uint8 (or float already scaled) -> (pi [N,A], v [N])."""
    # fp32 as the reference (atari_ff_model.py:50-51); a float64 state dict runs the same arithmetic in double
    # (tests: the "exact" side when two fp32 implementations differ by their summation order)
    img = image.type(sd["conv.conv.conv.0.weight"].dtype)
    if image.dtype == torch.uint8:
        img = img.mul_(1. / 255)                                                   # atari_ff_model.py:50-51
    x = F.relu(F.conv2d(img, sd["conv.conv.conv.0.weight"], sd["conv.conv.conv.0.bias"],
                        stride=STRIDES[0], padding=PADDINGS[0]))
    x = F.relu(F.conv2d(x, sd["conv.conv.conv.2.weight"], sd["conv.conv.conv.2.bias"],
                        stride=STRIDES[1], padding=PADDINGS[1]))
    x = F.relu(F.linear(x.reshape(x.shape[0], -1), sd["conv.head.model.0.weight"],
                        sd["conv.head.model.0.bias"]))                             # conv2d.py:110-111
    pi = F.softmax(F.linear(x, sd["pi.weight"], sd["pi.bias"]), dim=-1)            # atari_ff_model.py:55
    v = F.linear(x, sd["value.weight"], sd["value.bias"]).squeeze(-1)              # :56
    return pi, v
