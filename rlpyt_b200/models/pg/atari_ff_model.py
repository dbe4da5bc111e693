"""Feed-forward Atari actor-critic network (mirror of ``rlpyt/models/pg/atari_ff_model.py:8-63``).

conv(C->16,k8,s4) ReLU conv(16->32,k4,s2,p1) ReLU fc(->512) ReLU -> {softmax pi, value}.
Submodule names equal the reference's, so ``state_dict``s are interchangeable.  uint8 images
are converted with ``float(x) * (1/255)`` - one fp32 rounding, exactly the reference's
``img.type(float).mul_(1./255)`` (:50-51).
"""
import torch
import torch.nn.functional as F

from rlpyt_b200.models.conv2d import Conv2dHeadModel
from rlpyt_b200.utils.tensor import infer_leading_dims, restore_leading_dims


class AtariFfModel(torch.nn.Module):

    def __init__(self, image_shape, output_size, fc_sizes=512, use_maxpool=False, channels=None,
                 kernel_sizes=None, strides=None, paddings=None):
        super().__init__()
        self.image_shape = tuple(image_shape)
        self.conv = Conv2dHeadModel(
            image_shape=image_shape,
            channels=channels or [16, 32],
            kernel_sizes=kernel_sizes or [8, 4],
            strides=strides or [4, 2],
            paddings=paddings or [0, 1],
            use_maxpool=use_maxpool,
            hidden_sizes=fc_sizes,
        )
        self.pi = torch.nn.Linear(self.conv.output_size, output_size)
        self.value = torch.nn.Linear(self.conv.output_size, 1)

    def forward(self, image, prev_action, prev_reward):
        """[T,B,C,H,W] / [B,C,H,W] / [C,H,W] uint8 -> (pi, v) with the same leading dims."""
        img = image.type(torch.float)
        img = img.mul_(1. / 255)
        lead_dim, T, B, img_shape = infer_leading_dims(img, 3)
        fc_out = self.conv(img.view(T * B, *img_shape))
        pi = F.softmax(self.pi(fc_out), dim=-1)
        v = self.value(fc_out).squeeze(-1)
        return restore_leading_dims((pi, v), lead_dim, T, B)
