"""Policy + value heads of the policy-gradient models as ONE autograd op over csrc/pg_heads.cu:
``pi = softmax(h W_pi^T + b_pi)``, ``v = h w_v^T + b_v`` (rlpyt/models/pg/atari_ff_model.py:56-61).  The heads are 6
(actions) and 1 columns wide - no GEMM tile fits them: per update torch spends ~120 us in five SIMT GEMM / GEMV launches,
split-K and bias reductions, softmax forward / backward and the add of the two input gradients; here the forward is one
kernel and the backward one kernel + a small fixed-order reduction, reading ``h`` once each."""
import torch

from rlpyt_b200 import _lib
from rlpyt_b200.algos.optim import grad_destination

_SCRATCH = {}
MAX_F, MAX_A = 1024, 32


def usable(h, n_features, n_actions):
    return (isinstance(h, torch.Tensor) and h.is_cuda and h.dtype == torch.float32 and h.dim() == 2 and n_features <= MAX_F
            and 1 <= n_actions <= MAX_A and (n_actions + 1) * n_features * 4 <= 48 * 1024)


class PgHeads(torch.autograd.Function):

    @staticmethod
    def forward(ctx, h, w_pi, b_pi, w_v, b_v):
        _lib.require_cuda(h, w_pi, b_pi, w_v, b_v)
        h = h.contiguous()
        N, F = h.shape
        A = w_pi.shape[0]
        wp, bp = w_pi.detach().contiguous(), b_pi.detach().contiguous()
        wv, bv = w_v.detach().contiguous().view(-1), b_v.detach().contiguous().view(-1)
        pi = torch.empty((N, A), dtype=torch.float32, device=h.device)
        v = torch.empty(N, dtype=torch.float32, device=h.device)
        with torch.cuda.device(h.device):
            _lib.call("rl_pg_heads_forward_f32", _lib.ptr(h.detach()), _lib.ptr(wp), _lib.ptr(bp), _lib.ptr(wv), _lib.ptr(bv),
                      _lib.ptr(pi), _lib.ptr(v), N, F, A, _lib.stream())
        ctx.save_for_backward(h, w_pi, w_v, pi)
        ctx.params = (w_pi, b_pi, w_v, b_v)
        return pi, v

    @staticmethod
    def backward(ctx, g_pi, g_v):
        h, w_pi, w_v, pi = ctx.saved_tensors
        N, F = h.shape
        A = w_pi.shape[0]
        dev = h.device
        g_pi = None if g_pi is None else g_pi.contiguous()
        g_v = None if g_v is None else g_v.contiguous()
        gh = torch.empty_like(h)
        gwp, gbp, gwv, gbv = (grad_destination(p) for p in ctx.params)     # the parameters' slots of the flat gradient buffer
        nbytes = int(_lib.load().rl_pg_heads_backward_scratch_bytes(N, F, A))
        key = (str(dev), nbytes)
        sc = _SCRATCH.get(key)
        if sc is None:
            sc = _SCRATCH[key] = torch.empty(nbytes // 4 + 4, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.call("rl_pg_heads_backward_f32", _lib.ptr(h.detach()), _lib.ptr(pi), _lib.ptr(g_pi), _lib.ptr(g_v),
                      _lib.ptr(w_pi.detach().contiguous()), _lib.ptr(w_v.detach().contiguous().view(-1)), _lib.ptr(gh), _lib.ptr(gwp),
                      _lib.ptr(gbp), _lib.ptr(gwv), _lib.ptr(gbv), N, F, A, _lib.ptr(sc), _lib.stream(), n_launch=2)
        return gh, gwp, gbp, gwv, gbv


def policy_value_heads(h, pi_linear, value_linear):
    """``(softmax(pi_linear(h)), value_linear(h).squeeze(-1))`` for two ``torch.nn.Linear`` modules."""
    return PgHeads.apply(h, pi_linear.weight, pi_linear.bias, value_linear.weight, value_linear.bias)
