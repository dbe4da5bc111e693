"""``FlatAdam``: Adam over ONE flat fp32 parameter buffer with the global-norm clip fused in
(csrc/optim.cu), and the single gradient all-reduce of the data-parallel path.

It replaces three things of the reference's update step (rlpyt/algos/pg/ppo.py:101-104):
``DistributedDataParallel``'s bucketed all-reduce (rlpyt/agents/base.py:118-136),
``torch.nn.utils.clip_grad_norm_`` and ``torch.optim.Adam.step``.  ``state_dict`` /
``load_state_dict`` use torch.optim.Adam's format so ``initial_optim_state_dict`` snapshots
written by the reference load unchanged (and vice versa).
"""
import os

import torch

from rlpyt_b200 import _lib

DIRECT_GRADS = os.environ.get("RLPYT_B200_DIRECT_GRADS", "1") == "1"


def grad_destination(param, shape=None):
    """Where a hand-written backward should WRITE the gradient of ``param``: the parameter's slot of ``FlatAdam``'s flat
    gradient buffer when the optimizer manages it and nothing has claimed the slot since the last ``zero_grad`` (autograd
    then adopts the returned view as ``param.grad`` without a copy: the slot was zeroed and ``param.grad`` is ``None``) -
    otherwise a fresh tensor, which autograd accumulates as usual.  Saves one add kernel per parameter and update (ten
    per PPO update) over accumulating into ``.grad`` views of the flat buffer."""
    st = getattr(param, "_flat_grad", None)
    if st is not None and param.grad is None and st["stamp"] != st["owner"]._stamp and st["owner"].direct_grads:
        st["stamp"] = st["owner"]._stamp
        return st["slot"].view(param.shape)                  # a new view object: autograd may steal it
    return torch.empty(param.shape if shape is None else shape, dtype=param.dtype, device=param.device)


class FlatAdam(torch.optim.Optimizer):

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        params = list(params)
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        assert len(self.param_groups) == 1, "FlatAdam keeps one flat group"
        ps = self.param_groups[0]["params"]
        if not ps or not all(p.is_cuda and p.dtype == torch.float32 for p in ps):
            raise _lib.B200LibraryError("FlatAdam needs fp32 CUDA parameters (call agent.to_device first)")
        dev = ps[0].device
        self._offsets, total = [], 0
        for p in ps:
            self._offsets.append(total)
            total += (p.numel() + 3) // 4 * 4  # keep every view 16-byte aligned
        self.numel = total
        self.flat_param = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(total, dtype=torch.float32, device=dev)
        self.direct_grads = DIRECT_GRADS
        self._stamp = 0
        for p, off in zip(ps, self._offsets):
            self.flat_param[off:off + p.numel()].copy_(p.data.reshape(-1))
            p.data = self.flat_param[off:off + p.numel()].view(p.shape)
            p.grad = self.flat_grad[off:off + p.numel()].view(p.shape)
            p._flat_grad = dict(owner=self, slot=self.flat_grad[off:off + p.numel()], stamp=-1)   # see grad_destination
        self.step_count = 0
        nbytes = int(_lib.load().rl_clip_adam_scratch_bytes(total))
        self._scratch = torch.empty(nbytes // 8, dtype=torch.float64, device=dev)
        self._norm = torch.zeros(1, dtype=torch.float32, device=dev)
        self.world_size = 1

    # ---- gradient buffer management --------------------------------------------------------
    def zero_grad(self, set_to_none=False):
        """One memset of the flat buffer.  ``direct_grads``: every ``.grad`` is dropped, so that backward kernels write
        their result straight into the parameter's slot (``grad_destination``) and autograd adopts that view instead of
        adding into it; ``_collect_grads`` repairs whatever arrived some other way.  Otherwise the per-parameter
        ``.grad`` views of the flat buffer stay attached and autograd accumulates into them."""
        self.flat_grad.zero_()
        self._stamp += 1
        if self.direct_grads:
            for p in self.param_groups[0]["params"]:
                p.grad = None
            return
        for p, off in zip(self.param_groups[0]["params"], self._offsets):
            if p.grad is None or p.grad.data_ptr() != self.flat_grad.data_ptr() + 4 * off:
                p.grad = self.flat_grad[off:off + p.numel()].view(p.shape)

    def _collect_grads(self):
        """Make the flat buffer hold every gradient: a ``.grad`` that is not the parameter's slot (torch-native layers,
        accumulated or cloned gradients) is copied into it; a missing one leaves the slot zero.  The foreign tensor
        STAYS attached: when the backward is a replayed CUDA graph it is the graph's own buffer, rewritten by every
        replay without any host code running - it has to be collected again each time."""
        base = self.flat_grad.data_ptr()
        for p, off in zip(self.param_groups[0]["params"], self._offsets):
            g = p.grad
            if g is None or g.data_ptr() == base + 4 * off:
                continue
            self.flat_grad[off:off + p.numel()].view(p.shape).copy_(g)

    def set_world_size(self, world_size):
        self.world_size = int(world_size)

    # ---- update ----------------------------------------------------------------------------
    @torch.no_grad()
    def clip_and_step(self, max_norm=None, out=None):
        """[all-reduce ->] grad-norm -> clip -> Adam in two kernels.  Returns a 1-element CUDA
        tensor holding the pre-clip gradient norm (no host sync); ``out`` (1-element fp32 CUDA tensor) receives it
        in place when given."""
        g = self.param_groups[0]
        if self.direct_grads:
            self._collect_grads()
        if self.world_size > 1:
            import torch.distributed as dist
            dist.all_reduce(self.flat_grad)  # SUM over ranks; the mean is folded into grad_scale
        self.step_count += 1
        self._opt_called = True  # what torch's lr_scheduler step-order check looks at
        norm = out if out is not None else torch.empty(1, dtype=torch.float32, device=self.flat_param.device)
        with torch.cuda.device(self.flat_param.device):
            _lib.call("rl_clip_adam_f32", _lib.ptr(self.flat_param), _lib.ptr(self.flat_grad),
                      _lib.ptr(self.exp_avg), _lib.ptr(self.exp_avg_sq), self.numel, float(g["lr"]),
                      float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]), float(g["weight_decay"]),
                      self.step_count, float(max_norm) if max_norm else 0.0, 1.0 / self.world_size,
                      _lib.ptr(norm), _lib.ptr(self._scratch), _lib.stream(), n_launch=2)
        return norm

    def step(self, closure=None):
        self.clip_and_step(None)

    # ---- torch.optim.Adam-compatible (de)serialisation ----------------------------------------
    def state_dict(self):
        ps = self.param_groups[0]["params"]
        state = {}
        if self.step_count > 0:
            for i, (p, off) in enumerate(zip(ps, self._offsets)):
                n = p.numel()
                state[i] = dict(step=torch.tensor(float(self.step_count)),
                                exp_avg=self.exp_avg[off:off + n].view(p.shape).clone(),
                                exp_avg_sq=self.exp_avg_sq[off:off + n].view(p.shape).clone())
        g = {k: v for k, v in self.param_groups[0].items() if k != "params"}
        g.setdefault("amsgrad", False)
        g["params"] = list(range(len(ps)))
        return dict(state=state, param_groups=[g])

    def load_state_dict(self, state_dict):
        ps = self.param_groups[0]["params"]
        for k, v in state_dict["param_groups"][0].items():
            if k != "params":
                self.param_groups[0][k] = v
        steps = set()
        for i, (p, off) in enumerate(zip(ps, self._offsets)):
            st = state_dict["state"].get(i)
            if st is None:
                continue
            n = p.numel()
            self.exp_avg[off:off + n].copy_(st["exp_avg"].reshape(-1))
            self.exp_avg_sq[off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
            steps.add(int(float(st["step"])))
        if steps:
            assert len(steps) == 1, "per-parameter step counts differ"
            self.step_count = steps.pop()
