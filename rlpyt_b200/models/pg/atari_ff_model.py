"""Feed-forward Atari actor-critic network (mirror of ``rlpyt/models/pg/atari_ff_model.py:8-63``).

conv(C->16,k8,s4) ReLU conv(16->32,k4,s2,p1) ReLU fc(->512) ReLU -> {softmax pi, value}.
Submodule names equal the reference's, so ``state_dict``s are interchangeable.  uint8 images
are converted with ``float(x) * (1/255)`` - one fp32 rounding, exactly the reference's
``img.type(float).mul_(1./255)`` (:50-51).
"""
import os

import torch
import torch.nn.functional as F

from rlpyt_b200.models import conv1_op, conv2_op, gemm_op, heads_op
from rlpyt_b200.models.conv2d import Conv2dHeadModel
from rlpyt_b200.utils.gather import HostMappedFrames, LazyRows
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
        # uint8 CUDA frames take the hand-written first layer (csrc/conv1.cu) when the layer is the
        # reference default; anything else goes through the generic torch path below.
        layers = list(self.conv.conv.conv)
        self.fused_first_layer = conv1_op.supported(self.image_shape, layers)
        self.tc_second_layer = len(layers) == 4 and conv2_op.supported(layers[2], layers[3])
        self.accepts_lazy_rows = True
        # agent.step may be handed frames that still lie in the page-locked step buffer (HostMappedFrames): the first
        # layer streams them over PCIe itself and records them in HBM on the way (conv1_op.conv1_u8_relu_stream)
        self.accepts_host_mapped_frames = bool(self.fused_first_layer and len(self.image_shape) == 3
                                               and conv1_op.stream_supported(*self.image_shape))

    def _fused_forward(self, obs, rows, lead_shape):
        layers = self.conv.conv.conv
        x = conv1_op.conv1_u8_relu(layers[0].weight, layers[0].bias, obs, rows)
        # the ReLU backward of the second layer rides in the epilogue of the fc layer's input-gradient GEMM when both
        # layers are on this package's kernels (one 315 MB elementwise pass less per update)
        fuse = self.tc_second_layer and torch.is_grad_enabled() and self._head_fuses_input_relu(x.shape[0])
        if self.tc_second_layer:
            x = conv2_op.conv2_relu(x, layers[2].weight, layers[2].bias, grad_is_masked=fuse)
        else:
            x = layers[2:](x)
        fc_out = self._head(x.view(x.shape[0], -1), input_is_relu_output=fuse)
        if (os.environ.get("RLPYT_B200_FUSED_HEADS", "1") == "1" and isinstance(self.pi, torch.nn.Linear)
                and isinstance(self.value, torch.nn.Linear) and heads_op.usable(fc_out, self.pi.in_features, self.pi.out_features)):
            pi, v = heads_op.policy_value_heads(fc_out, self.pi, self.value)      # both heads + softmax: one kernel each way
        else:
            pi = F.softmax(self.pi(fc_out), dim=-1)
            v = self.value(fc_out).squeeze(-1)
        return pi.view(lead_shape + pi.shape[1:]), v.view(lead_shape)

    # tiny batches (the single example step at start-up) stay on cuBLAS; from B=64 up the tcgen05 GEMM
    # is used - with split-K when the 128x128 tile grid cannot fill the 148 SMs (agent.step, M=256)
    TC_GEMM_MIN_ROWS = 64

    def _head_mods(self):
        head = self.conv.head
        mods = list(head.model) if isinstance(head, torch.nn.Module) and hasattr(head, "model") else None
        if (mods is not None and len(mods) == 2 and isinstance(mods[0], torch.nn.Linear) and isinstance(mods[1], torch.nn.ReLU)
                and gemm_op.usable(mods[0].in_features, mods[0].out_features)):
            return mods
        return None

    def _head_fuses_input_relu(self, n_rows):
        # measured on the B200 (profiles/r02_launches_ppo_iter.csv): the masked epilogue's uncoalesced mask reads stall the
        # GEMM's drain warps - 208 us against 130 us + a 55 us ReLU-backward pass - so the fusion stays opt-in
        if os.environ.get("RLPYT_B200_FUSE_RELU_BWD", "0") != "1":
            return False
        return self._head_mods() is not None and n_rows >= self.TC_GEMM_MIN_ROWS and gemm_op.fuses_input_relu(n_rows)

    def _head(self, flat, input_is_relu_output=False):
        """Linear(conv_out -> fc) + ReLU: the fp32-accurate tcgen05 GEMM for minibatch-sized inputs."""
        mods = self._head_mods()
        if mods is not None and flat.is_cuda and flat.shape[0] >= self.TC_GEMM_MIN_ROWS:
            return gemm_op.linear_tf32x3(flat, mods[0].weight, mods[0].bias, relu=True, input_is_relu_output=input_is_relu_output)
        assert not input_is_relu_output
        return self.conv.head(flat)

    @torch.no_grad()
    def forward_step(self, image, prev_action, prev_reward, distribution, uniform=None):
        """``agent.step``'s forward: (pi, v, action).  For contiguous uint8 CUDA frames [B,C,H,W] the policy / value
        heads, softmax and the categorical draw are ONE kernel (csrc/categorical.cu: pg_head_sample_kernel) after the
        fused trunk; anything else takes ``forward`` + ``distribution.sample``."""
        from rlpyt_b200 import _lib
        from rlpyt_b200.distributions.categorical import DistInfo
        A = self.pi.out_features
        mapped = isinstance(image, HostMappedFrames)
        if mapped and not (self.accepts_host_mapped_frames and A <= 32 and hasattr(distribution, "_rng_state")):
            raise ValueError("HostMappedFrames need the fused kind::i8 first layer and the fused policy head")
        if mapped or (self.fused_first_layer and isinstance(image, torch.Tensor) and image.dtype == torch.uint8 and image.is_cuda
                      and image.is_contiguous() and image.dim() == 4 and A <= 32 and hasattr(distribution, "_rng_state")):
            layers = self.conv.conv.conv
            if mapped:
                x = conv1_op.conv1_u8_relu_stream(layers[0].weight, layers[0].bias, image)
            else:
                x = conv1_op.conv1_u8_relu(layers[0].weight, layers[0].bias, image, None)
            x = conv2_op.conv2_relu(x, layers[2].weight, layers[2].bias) if self.tc_second_layer else layers[2:](x)
            h = self._head(x.view(x.shape[0], -1)).contiguous()
            Bn, F_ = h.shape
            pi = torch.empty((Bn, A), dtype=torch.float32, device=h.device)
            v = torch.empty(Bn, dtype=torch.float32, device=h.device)
            action = torch.empty(Bn, dtype=torch.int64, device=h.device)
            if uniform is not None:
                uniform = uniform.reshape(-1).to(device=h.device, dtype=torch.float32).contiguous()
            state = None if uniform is not None else distribution._rng_state(h.device)
            with torch.cuda.device(h.device):
                _lib.call("rl_pg_head_sample_f32", _lib.ptr(h), _lib.ptr(self.pi.weight.detach()), _lib.ptr(self.pi.bias.detach()),
                          _lib.ptr(self.value.weight.detach().view(-1)), _lib.ptr(self.value.bias.detach()), _lib.ptr(uniform),
                          _lib.ptr(state), _lib.ptr(pi), _lib.ptr(v), _lib.ptr(action), Bn, F_, A, _lib.stream())
            return pi, v, action
        pi, v = self.forward(image, prev_action, prev_reward)
        return pi, v, distribution.sample(DistInfo(prob=pi))

    def forward(self, image, prev_action, prev_reward):
        """[T,B,C,H,W] / [B,C,H,W] / [C,H,W] uint8 -> (pi, v) with the same leading dims."""
        if isinstance(image, LazyRows):
            if self.fused_first_layer and image.dtype == torch.uint8:
                return self._fused_forward(image.src.view((-1,) + self.image_shape), image.rows,
                                           (image.rows.numel(),))
            image = image.materialize()
        if (self.fused_first_layer and image.dtype == torch.uint8 and image.is_cuda
                and image.is_contiguous() and image.dim() >= 3):
            lead_shape = tuple(image.shape[:-3])
            return self._fused_forward(image.view((-1,) + self.image_shape), None, lead_shape)
        img = image.type(torch.float)
        img = img.mul_(1. / 255)
        lead_dim, T, B, img_shape = infer_leading_dims(img, 3)
        fc_out = self.conv(img.view(T * B, *img_shape))
        pi = F.softmax(self.pi(fc_out), dim=-1)
        v = self.value(fc_out).squeeze(-1)
        return restore_leading_dims((pi, v), lead_dim, T, B)
