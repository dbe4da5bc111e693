"""Recurrent Atari actor-critic network (mirror of ``rlpyt/models/pg/atari_lstm_model.py:12-89``):
conv(C->16,k8,s4) ReLU conv(16->32,k4,s2,p1) ReLU fc(->512) ReLU -> LSTM(512 + A + 1 -> 512) -> {softmax pi, value}.
Submodule names equal the reference's (``conv``, ``lstm``, ``pi``, ``value``) so ``state_dict``s are interchangeable.
The convolutional trunk is the feed-forward model's (uint8 CUDA frames -> the kind::i8 first layer, the s2d second
layer, the 3xTF32 GEMM); the LSTM itself is ``torch.nn.LSTM`` (cuDNN) - SURVEY.md 8(f) row 3 lists the recurrent
agents as a widening of the path, not one of its kernels.
"""
import torch
import torch.nn.functional as F

from rlpyt_b200.models.pg.atari_ff_model import AtariFfModel
from rlpyt_b200.utils.collections import namedarraytuple
from rlpyt_b200.utils.tensor import infer_leading_dims, restore_leading_dims

RnnState = namedarraytuple("RnnState", ["h", "c"])


class AtariLstmModel(torch.nn.Module):

    def __init__(self, image_shape, output_size, fc_sizes=512, lstm_size=512, use_maxpool=False, channels=None,
                 kernel_sizes=None, strides=None, paddings=None):
        super().__init__()
        # the feed-forward model owns the trunk (and its kernel dispatch); its heads are unused here
        trunk = AtariFfModel(image_shape, output_size, fc_sizes=fc_sizes, use_maxpool=use_maxpool, channels=channels,
                             kernel_sizes=kernel_sizes, strides=strides, paddings=paddings)
        self.image_shape = tuple(image_shape)
        self.conv = trunk.conv
        self._trunk = [trunk]                         # not a submodule: keeps state_dict names equal to the reference's
        self.lstm = torch.nn.LSTM(self.conv.output_size + output_size + 1, lstm_size)
        self.pi = torch.nn.Linear(lstm_size, output_size)
        self.value = torch.nn.Linear(lstm_size, 1)

    def _features(self, image, T, B):
        """[T*B, fc] features of the frames: fused kernels for contiguous uint8 CUDA frames, torch otherwise."""
        trunk = self._trunk[0]
        if (trunk.fused_first_layer and image.dtype == torch.uint8 and image.is_cuda and image.is_contiguous()):
            layers = self.conv.conv.conv
            from rlpyt_b200.models import conv1_op, conv2_op
            x = conv1_op.conv1_u8_relu(layers[0].weight, layers[0].bias, image.view((-1,) + self.image_shape), None)
            x = conv2_op.conv2_relu(x, layers[2].weight, layers[2].bias) if trunk.tc_second_layer else layers[2:](x)
            return trunk._head(x.view(x.shape[0], -1))
        img = image.type(torch.float)
        img = img.mul_(1. / 255)
        return self.conv(img.view(T * B, *self.image_shape))

    def forward(self, image, prev_action, prev_reward, init_rnn_state):
        """image [T,B,C,H,W] / [B,...] / [...] uint8, prev_action one-hot, prev_reward, init_rnn_state (h, c) [N,B,H] or
        None -> (pi, v, RnnState) with the input's leading dims; the state keeps its B dimension."""
        lead_dim, T, B, _ = infer_leading_dims(image, 3)
        fc_out = self._features(image, T, B)
        lstm_input = torch.cat([fc_out.view(T, B, -1), prev_action.view(T, B, -1).to(fc_out.dtype),
                                prev_reward.view(T, B, 1).to(fc_out.dtype)], dim=2)
        init = None if init_rnn_state is None else tuple(t.contiguous() for t in init_rnn_state)
        lstm_out, (hn, cn) = self.lstm(lstm_input, init)
        flat = lstm_out.view(T * B, -1)
        pi = F.softmax(self.pi(flat), dim=-1)
        v = self.value(flat).squeeze(-1)
        pi, v = restore_leading_dims((pi, v), lead_dim, T, B)
        return pi, v, RnnState(h=hn, c=cn)
