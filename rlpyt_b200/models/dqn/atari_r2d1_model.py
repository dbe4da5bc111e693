"""Recurrent Q-network for Atari (mirror of ``rlpyt/models/dqn/atari_r2d1_model.py:13-77``):
conv(C->32,k8,s4) conv(32->64,k4,s2,p1) conv(64->64,k3,s1,p1) each + ReLU, fc(->512) ReLU, LSTM(512 + A + 1 -> 512),
MLP / dueling head -> Q[A].  Submodule names equal the reference's (``conv.conv``, ``conv.head``, ``lstm``, ``head``) so
state_dicts are interchangeable.  The fc layers between conv and LSTM and inside the head run on the fp32-accurate
tcgen05 GEMM for training-sized batches ([T*B] rows); conv and LSTM are torch (cuDNN), as in the feed-forward
Q-network (models/dqn/atari_dqn_model.py)."""
import torch

from rlpyt_b200.models import gemm_op
from rlpyt_b200.models.conv2d import Conv2dHeadModel
from rlpyt_b200.models.dqn.dueling import DuelingHeadModel
from rlpyt_b200.models.mlp import MlpModel
from rlpyt_b200.utils.collections import namedarraytuple
from rlpyt_b200.utils.tensor import infer_leading_dims, restore_leading_dims

RnnState = namedarraytuple("RnnState", ["h", "c"])


def _mlp(mlp, x, min_rows):
    """MlpModel forward with its Linear+ReLU pairs on the tensor-core GEMM when the batch is large enough."""
    mods = list(mlp.model)
    i = 0
    while i < len(mods):
        m = mods[i]
        relu = i + 1 < len(mods) and isinstance(mods[i + 1], torch.nn.ReLU)
        if (isinstance(m, torch.nn.Linear) and x.is_cuda and x.shape[0] >= min_rows
                and gemm_op.usable(m.in_features, m.out_features)):
            x = gemm_op.linear_tf32x3(x, m.weight, m.bias, relu=relu)
            i += 2 if relu else 1
        else:
            x = m(x)
            i += 1
    return x


class AtariR2d1Model(torch.nn.Module):

    TC_GEMM_MIN_ROWS = 64

    def __init__(self, image_shape, output_size, fc_size=512, lstm_size=512, head_size=512, dueling=False,
                 use_maxpool=False, channels=None, kernel_sizes=None, strides=None, paddings=None):
        super().__init__()
        self.dueling = dueling
        self.conv = Conv2dHeadModel(
            image_shape=image_shape,
            channels=channels or [32, 64, 64],
            kernel_sizes=kernel_sizes or [8, 4, 3],
            strides=strides or [4, 2, 1],
            paddings=paddings or [0, 1, 1],
            use_maxpool=use_maxpool,
            hidden_sizes=fc_size,
        )
        self.lstm = torch.nn.LSTM(self.conv.output_size + output_size + 1, lstm_size)
        if dueling:
            self.head = DuelingHeadModel(lstm_size, head_size, output_size)
        else:
            self.head = MlpModel(lstm_size, head_size, output_size=output_size)

    def forward(self, observation, prev_action, prev_reward, init_rnn_state):
        """observation [T,B,C,H,W] / [B,...] / [...] uint8, prev_action one-hot, prev_reward, init_rnn_state (h, c)
        [N,B,H] or None -> (q with the input's leading dims, RnnState [N,B,H])."""
        img = observation.type(torch.float)
        img = img.mul_(1. / 255)
        lead_dim, T, B, img_shape = infer_leading_dims(img, 3)
        feat = self.conv.conv(img.view(T * B, *img_shape)).reshape(T * B, -1)
        conv_out = _mlp(self.conv.head, feat, self.TC_GEMM_MIN_ROWS)
        lstm_input = torch.cat([conv_out.view(T, B, -1), prev_action.view(T, B, -1).to(conv_out.dtype),
                                prev_reward.view(T, B, 1).to(conv_out.dtype)], dim=2)
        init = None if init_rnn_state is None else tuple(t.contiguous() for t in init_rnn_state)
        lstm_out, (hn, cn) = self.lstm(lstm_input, init)
        flat = lstm_out.reshape(T * B, -1)
        q = self.head(flat) if self.dueling else _mlp(self.head, flat, self.TC_GEMM_MIN_ROWS)
        q = restore_leading_dims(q, lead_dim, T, B)
        return q, RnnState(h=hn, c=cn)
