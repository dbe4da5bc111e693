"""Convolutional Q-network for Atari (mirror of ``rlpyt/models/dqn/atari_dqn_model.py:9-67``):
conv(C->32,k8,s4) conv(32->64,k4,s2,p1) conv(64->64,k3,s1,p1), each + ReLU, fc(->512) ReLU -> Q[A]
(or the dueling head).  Submodule names equal the reference's, so state_dicts are interchangeable.

The hidden layer of the non-dueling head (3136/7744 -> 512, the GEMM-shaped part) runs on the
fp32-accurate tcgen05 GEMM for training-sized batches; the conv stack is torch/cuDNN this round - the
tcgen05 conv kernels of csrc/conv_tc.cu are specialised to the AtariFf channel counts."""
import torch

from rlpyt_b200.models import gemm_op
from rlpyt_b200.models.conv2d import Conv2dModel
from rlpyt_b200.models.dqn.dueling import DuelingHeadModel
from rlpyt_b200.models.mlp import MlpModel
from rlpyt_b200.utils.tensor import infer_leading_dims, restore_leading_dims


class AtariDqnModel(torch.nn.Module):

    TC_GEMM_MIN_ROWS = 64

    def __init__(self, image_shape, output_size, fc_sizes=512, dueling=False, use_maxpool=False, channels=None,
                 kernel_sizes=None, strides=None, paddings=None):
        super().__init__()
        self.dueling = dueling
        c, h, w = image_shape
        self.conv = Conv2dModel(
            in_channels=c,
            channels=channels or [32, 64, 64],
            kernel_sizes=kernel_sizes or [8, 4, 3],
            strides=strides or [4, 2, 1],
            paddings=paddings or [0, 1, 1],
            use_maxpool=use_maxpool,
        )
        conv_out_size = self.conv.conv_out_size(h, w)
        if dueling:
            self.head = DuelingHeadModel(conv_out_size, fc_sizes, output_size)
        else:
            self.head = MlpModel(conv_out_size, fc_sizes, output_size)

    def _head(self, flat):
        mods = list(self.head.model) if isinstance(self.head, MlpModel) else None
        if (mods is not None and len(mods) == 3 and isinstance(mods[0], torch.nn.Linear)
                and isinstance(mods[1], torch.nn.ReLU) and flat.is_cuda and flat.shape[0] >= self.TC_GEMM_MIN_ROWS
                and gemm_op.usable(mods[0].in_features, mods[0].out_features)):
            return mods[2](gemm_op.linear_tf32x3(flat, mods[0].weight, mods[0].bias, relu=True))
        return self.head(flat)

    def forward(self, observation, prev_action, prev_reward):
        """[T,B,C,H,W] / [B,C,H,W] / [C,H,W] uint8 -> Q with the same leading dims."""
        img = observation.type(torch.float)
        img = img.mul_(1. / 255)
        lead_dim, T, B, img_shape = infer_leading_dims(img, 3)
        conv_out = self.conv(img.view(T * B, *img_shape))
        q = self._head(conv_out.reshape(T * B, -1))
        return restore_leading_dims(q, lead_dim, T, B)
