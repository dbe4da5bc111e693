"""Shape helpers for the conv models (mirror of ``rlpyt/models/utils.py:5-15, :57``)."""


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
