"""Fused PPO / A2C loss as ``torch.autograd.Function``s over the C-ABI kernels
(csrc/pg_loss.cu): one launch produces loss / entropy / perplexity AND dLoss/dprob,
dLoss/dvalue; backward is a scale by the upstream gradient.

Replaces the ~15 torch-CPU ops + autograd graph of rlpyt/algos/pg/ppo.py:136-153 and
rlpyt/algos/pg/a2c.py:88-100.
"""
import torch

from rlpyt_b200 import _lib

_SCRATCH = {}


def _scratch(n, device):
    nbytes = int(_lib.load().rl_pg_loss_scratch_bytes(n))
    key = (str(device), nbytes)
    s = _SCRATCH.get(key)
    if s is None:
        s = torch.empty(nbytes // 8, dtype=torch.float64, device=device)
        _SCRATCH[key] = s
    return s


def _prep(*tensors):
    out = []
    for t in tensors:
        if t is None:
            out.append(None)
            continue
        _lib.require_cuda(t)
        out.append(t.detach().contiguous())
    return out


class _PpoLoss(torch.autograd.Function):

    @staticmethod
    def forward(ctx, prob_new, value, prob_old, action, return_, advantage, valid,
                ratio_clip, value_loss_coeff, entropy_loss_coeff):
        p, v, po, a, R, A, m = _prep(prob_new, value, prob_old, action, return_, advantage, valid)
        assert p.dtype == torch.float32 and a.dtype == torch.int64
        n_act = p.shape[-1]
        N = p.numel() // n_act
        if m is not None and m.dtype != torch.float32:
            m = m.float()
        scalars = torch.empty(8, dtype=torch.float32, device=p.device)
        need_grad = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        gp = torch.empty_like(p) if need_grad else None
        gv = torch.empty_like(v) if need_grad else None
        with torch.cuda.device(p.device):
            if isinstance(ratio_clip, torch.Tensor):      # device scalar: the value is read when the kernel RUNS (graph replays)
                assert ratio_clip.is_cuda and ratio_clip.dtype == torch.float32 and ratio_clip.numel() == 1
                _lib.call("rl_ppo_loss_devclip_f32", _lib.ptr(p), _lib.ptr(v), _lib.ptr(po), _lib.ptr(a), _lib.ptr(R),
                          _lib.ptr(A), _lib.ptr(m), N, n_act, _lib.ptr(ratio_clip), float(value_loss_coeff),
                          float(entropy_loss_coeff), _lib.ptr(scalars), _lib.ptr(gp), _lib.ptr(gv),
                          _lib.ptr(_scratch(N, p.device)), _lib.stream(), n_launch=2)
            else:
                _lib.call("rl_ppo_loss_f32", _lib.ptr(p), _lib.ptr(v), _lib.ptr(po), _lib.ptr(a), _lib.ptr(R),
                          _lib.ptr(A), _lib.ptr(m), N, n_act, float(ratio_clip), float(value_loss_coeff),
                          float(entropy_loss_coeff), _lib.ptr(scalars), _lib.ptr(gp), _lib.ptr(gv),
                          _lib.ptr(_scratch(N, p.device)), _lib.stream(), n_launch=2)
        if need_grad:
            ctx.save_for_backward(gp, gv)
        ctx.mark_non_differentiable(scalars)
        return scalars[0], scalars

    @staticmethod
    def backward(ctx, g_loss, _g_scalars):
        gp, gv = ctx.saved_tensors
        return (gp * g_loss, gv * g_loss) + (None,) * 8


class _A2cLoss(torch.autograd.Function):

    @staticmethod
    def forward(ctx, prob, value, action, return_, advantage, valid, value_loss_coeff,
                entropy_loss_coeff):
        p, v, a, R, A, m = _prep(prob, value, action, return_, advantage, valid)
        assert p.dtype == torch.float32 and a.dtype == torch.int64
        n_act = p.shape[-1]
        N = p.numel() // n_act
        if m is not None and m.dtype != torch.float32:
            m = m.float()
        scalars = torch.empty(8, dtype=torch.float32, device=p.device)
        need_grad = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        gp = torch.empty_like(p) if need_grad else None
        gv = torch.empty_like(v) if need_grad else None
        with torch.cuda.device(p.device):
            _lib.call("rl_a2c_loss_f32", _lib.ptr(p), _lib.ptr(v), _lib.ptr(a), _lib.ptr(R), _lib.ptr(A),
                      _lib.ptr(m), N, n_act, float(value_loss_coeff), float(entropy_loss_coeff),
                      _lib.ptr(scalars), _lib.ptr(gp), _lib.ptr(gv), _lib.ptr(_scratch(N, p.device)),
                      _lib.stream(), n_launch=2)
        if need_grad:
            ctx.save_for_backward(gp, gv)
        ctx.mark_non_differentiable(scalars)
        return scalars[0], scalars

    @staticmethod
    def backward(ctx, g_loss, _g_scalars):
        gp, gv = ctx.saved_tensors
        return (gp * g_loss, gv * g_loss) + (None,) * 6


def ppo_loss(prob_new, value, prob_old, action, return_, advantage, valid, ratio_clip,
             value_loss_coeff, entropy_loss_coeff):
    """-> (loss [differentiable 0-dim], scalars[8] = loss, entropy, perplexity, pi_loss,
    value_loss, n_valid, 0, 0).  All on device, no host sync."""
    return _PpoLoss.apply(prob_new, value, prob_old, action, return_, advantage, valid,
                          ratio_clip, value_loss_coeff, entropy_loss_coeff)


def a2c_loss(prob, value, action, return_, advantage, valid, value_loss_coeff, entropy_loss_coeff):
    return _A2cLoss.apply(prob, value, action, return_, advantage, valid, value_loss_coeff,
                          entropy_loss_coeff)
