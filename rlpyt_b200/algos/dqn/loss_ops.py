"""Fused DQN loss as a ``torch.autograd.Function`` over csrc/dqn_loss.cu: one launch pair produces the
loss, the TD-error priorities and dLoss/dqs; backward is a scale by the upstream gradient.
Replaces the ~20 torch-CPU ops of rlpyt/algos/dqn/dqn.py:230-263 (and the D2H copies of both
networks' outputs that precede them in the reference)."""
import torch

from rlpyt_b200 import _lib

_SCRATCH = {}


def _scratch(n, device):
    nbytes = int(_lib.load().rl_dqn_loss_scratch_bytes(n))
    key = (str(device), nbytes)
    s = _SCRATCH.get(key)
    if s is None:
        s = torch.empty(nbytes // 8, dtype=torch.float64, device=device)
        _SCRATCH[key] = s
    return s


def _c(t, dtype=None):
    if t is None:
        return None
    _lib.require_cuda(t)
    t = t.detach()
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


class _DqnLoss(torch.autograd.Function):

    @staticmethod
    def forward(ctx, qs, target_qs, next_qs, action, return_, done_n, is_weights, disc_n, delta_clip):
        q, tq, nq = _c(qs, torch.float32), _c(target_qs, torch.float32), _c(next_qs, torch.float32)
        a, R, w = _c(action, torch.int64), _c(return_, torch.float32), _c(is_weights, torch.float32)
        dn = _c(done_n)
        dn = dn.view(torch.uint8) if dn.dtype == torch.bool else dn.to(torch.uint8)
        n_act = q.shape[-1]
        N = q.numel() // n_act
        assert tq.shape == q.shape and (nq is None or nq.shape == q.shape) and a.numel() == N
        scalars = torch.empty(2, dtype=torch.float32, device=q.device)
        td_abs = torch.empty(a.shape, dtype=torch.float32, device=q.device)
        grad = torch.empty_like(q) if ctx.needs_input_grad[0] else None
        with torch.cuda.device(q.device):
            _lib.call("rl_dqn_loss_f32", _lib.ptr(q), _lib.ptr(tq), _lib.ptr(nq), _lib.ptr(a), _lib.ptr(R),
                      _lib.ptr(dn), _lib.ptr(w), N, n_act, float(disc_n),
                      -1.0 if delta_clip is None else float(delta_clip), _lib.ptr(scalars), _lib.ptr(td_abs),
                      _lib.ptr(grad), _lib.ptr(_scratch(N, q.device)), _lib.stream(), n_launch=2)
        if grad is not None:
            ctx.save_for_backward(grad)
        ctx.mark_non_differentiable(td_abs)
        return scalars[0], td_abs

    @staticmethod
    def backward(ctx, g_loss, _g_td):
        (grad,) = ctx.saved_tensors
        return (grad * g_loss,) + (None,) * 8


def dqn_loss(qs, target_qs, next_qs, action, return_, done_n, is_weights, disc_n, delta_clip):
    """-> (loss 0-dim, td_abs_errors [N]); ``next_qs`` None: plain DQN; ``is_weights`` None: uniform
    replay; ``delta_clip`` None: MSE.  All tensors CUDA."""
    return _DqnLoss.apply(qs, target_qs, next_qs, action, return_, done_n, is_weights, disc_n, delta_clip)
