"""rlpyt_b200 - B200-native (sm_100a) implementation of rlpyt's data-parallel inner loop:
rollout collection -> GAE / n-step returns -> PPO / A2C loss + minibatches -> prioritized frame
replay, behind rlpyt's Sampler / Algo / Agent / ReplayBuffer interfaces.  See DESIGN.md.

The arithmetic runs in hand-written CUDA kernels (rlpyt_b200/csrc, C ABI in include/rlpyt_b200.h);
torch supplies device memory, streams, autograd for the network and torch.distributed (NCCL).
There is no CPU fallback: without the built library and a CUDA device the package raises.
"""
import torch as _torch

# fp32 parity with the reference (which predates TF32): keep cuDNN / cuBLAS in true fp32.
_torch.backends.cudnn.allow_tf32 = False
_torch.backends.cuda.matmul.allow_tf32 = False
# Let cuDNN time its fp32 algorithms once per shape: the default heuristic picks a 3x slower
# implicit-GEMM for the 4-channel 8x8/stride-4 first layer (profiles/r01_profile_step_first.txt).
_torch.backends.cudnn.benchmark = True

__version__ = "0.1.0"
