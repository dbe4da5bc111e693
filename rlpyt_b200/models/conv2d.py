"""Conv stacks (mirror of ``rlpyt/models/conv2d.py:8-116``; same submodule names:
``conv.<i>`` inside Conv2dModel, ``conv`` + ``head`` inside Conv2dHeadModel)."""
import torch

from rlpyt_b200.models.mlp import MlpModel
from rlpyt_b200.models.utils import conv2d_output_shape


class Conv2dModel(torch.nn.Module):
    """Conv2d+nonlinearity layers; with ``use_maxpool`` strides become max-pools."""

    def __init__(self, in_channels, channels, kernel_sizes, strides, paddings=None,
                 nonlinearity=torch.nn.ReLU, use_maxpool=False, head_sizes=None):
        super().__init__()
        n = len(channels)
        paddings = [0] * n if paddings is None else paddings
        assert len(kernel_sizes) == len(strides) == len(paddings) == n
        pools = list(strides) if use_maxpool else [1] * n
        conv_strides = [1] * n if use_maxpool else list(strides)
        layers, ic = [], in_channels
        for oc, k, s, p, mp in zip(channels, kernel_sizes, conv_strides, paddings, pools):
            layers += [torch.nn.Conv2d(ic, oc, kernel_size=k, stride=s, padding=p), nonlinearity()]
            if mp > 1:
                layers.append(torch.nn.MaxPool2d(mp))
            ic = oc
        self.conv = torch.nn.Sequential(*layers)

    def forward(self, input):
        return self.conv(input)

    def conv_out_size(self, h, w, c=None):
        for layer in self.conv.children():
            if isinstance(layer, (torch.nn.Conv2d, torch.nn.MaxPool2d)):
                h, w = conv2d_output_shape(h, w, layer.kernel_size, layer.stride, layer.padding)
            if isinstance(layer, torch.nn.Conv2d):
                c = layer.out_channels
        return h * w * c


class Conv2dHeadModel(torch.nn.Module):
    """Conv2dModel followed by an MlpModel head on the flattened features."""

    def __init__(self, image_shape, channels, kernel_sizes, strides, hidden_sizes, output_size=None,
                 paddings=None, nonlinearity=torch.nn.ReLU, use_maxpool=False):
        super().__init__()
        c, h, w = image_shape
        self.conv = Conv2dModel(c, channels, kernel_sizes, strides, paddings=paddings,
                                nonlinearity=nonlinearity, use_maxpool=use_maxpool)
        flat = self.conv.conv_out_size(h, w)
        if hidden_sizes or output_size:
            self.head = MlpModel(flat, hidden_sizes, output_size=output_size, nonlinearity=nonlinearity)
            self._output_size = self.head.output_size
        else:
            self.head = lambda x: x
            self._output_size = flat

    def forward(self, input):
        return self.head(self.conv(input).view(input.shape[0], -1))

    @property
    def output_size(self):
        return self._output_size
