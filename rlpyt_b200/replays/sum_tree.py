"""Device-resident fp64 sum-tree (mirror of ``rlpyt/replays/sum_tree.py:8-222`` ``SumTree``; same
constructor, ``advance`` / ``sample`` / ``update_batch_priorities`` / ``reset``), computed by the
bit-exact kernels of csrc/sumtree.cu.  The cursor / on-off range arithmetic (sum_tree.py:60-99) is
integer bookkeeping and stays on the host; uniforms come from ``np.random.rand`` on the host exactly
like the reference (:107), so with equal seeds the sampled indices are identical.
"""
import numpy as np
import torch

from rlpyt_b200 import _lib


def _pow_alpha_f64(priorities, alpha, device):
    """``priorities ** alpha`` as numpy float32 pow, widened to fp64 (prioritized.py:49,79) - device kernel."""
    p = torch.as_tensor(priorities).to(device, dtype=torch.float32).contiguous()
    out = torch.empty(p.shape, dtype=torch.float64, device=device)
    with torch.cuda.device(device):
        _lib.call("rl_pow_f32_to_f64", _lib.ptr(p), float(np.float32(alpha)), _lib.ptr(out), p.numel(), _lib.stream())
    return out


class SumTree:

    async_ = False

    def __init__(self, T, B, off_backward, off_forward, default_value=1, enable_input_priorities=False,
                 input_priority_shift=0, device=None):
        self.T, self.B, self.size = T, B, T * B
        self.off_backward, self.off_forward = off_backward, off_forward
        self.default_value = default_value
        self.input_priority_shift = input_priority_shift
        self.tree_levels = int(np.ceil(np.log2(self.size + 1)) + 1)             # sum_tree.py:39
        if device is None:
            if not torch.cuda.is_available():
                raise _lib.B200LibraryError("rlpyt_b200 SumTree needs a CUDA device (no CPU fallback)")
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)
        self.tree = torch.zeros(2 ** self.tree_levels - 1, dtype=torch.float64, device=self.device)
        self.low_idx = 2 ** (self.tree_levels - 1) - 1
        self.high_idx = self.size + self.low_idx
        self.priorities = self.tree[self.low_idx:self.high_idx].view(T, B)       # same memory (:43)
        self.input_priorities = (torch.full((T, B), float(default_value), dtype=torch.float64, device=self.device)
                                 if enable_input_priorities else None)
        self._diffs = torch.empty(max(1, min(self.size, 1 << 16)), dtype=torch.float64, device=self.device)
        self.reset()

    def reset(self):
        self.tree.zero_()
        self.t = 0
        self._initial_wrap_guard = True
        self._sampled_unique = False
        self.prev_tree_idxs = None
        if self.input_priorities is not None:
            self.input_priorities.fill_(float(self.default_value))

    # ---- kernels -------------------------------------------------------------------------------
    def _scratch(self, n):
        if self._diffs.numel() < n:
            self._diffs = torch.empty(n, dtype=torch.float64, device=self.device)
        return self._diffs

    def _update_segment(self, n, leaf_idx=None, leaf_base=0, values=None, scalar=0.0):
        if n <= 0:
            return
        with torch.cuda.device(self.device):
            _lib.call("rl_sumtree_update_f64", _lib.ptr(self.tree), self.tree_levels, _lib.ptr(leaf_idx),
                      int(leaf_base), _lib.ptr(values), float(scalar), int(n), _lib.ptr(self._scratch(n)),
                      _lib.stream(), n_launch=2)

    # ---- cursor --------------------------------------------------------------------------------
    def advance(self, T, priorities=None):
        """sum_tree.py:60-99: enable [t-b, t+T-b), zero [t+T-b, t+T+f) (with wrap and the initial
        wrap guard), optionally storing input priorities."""
        if T == 0:
            return
        t, b, f = self.t, self.off_backward, self.off_forward
        low_on_t = (t - b) % self.T
        high_on_t = ((t + T - b - 1) % self.T) + 1
        low_off_t = (t + T - b) % self.T
        high_off_t = ((t + T + f - 1) % self.T) + 1
        if self._initial_wrap_guard:
            low_on_t = max(f, t - b)
            high_on_t = low_off_t = max(low_on_t, t + T - b)
            if t + T - b >= f:
                self._initial_wrap_guard = False
        if priorities is not None:
            assert self.input_priorities is not None, "Must enable input priorities."
            pri = torch.as_tensor(priorities, dtype=torch.float64, device=self.device)
            input_t = t - self.input_priority_shift
            if input_t < 0 or input_t + T > self.T:
                rows = torch.as_tensor(np.arange(input_t, input_t + T) % self.T, device=self.device)
                self.input_priorities[rows] = pri
            else:
                self.input_priorities[input_t:input_t + T] = pri
            if self._initial_wrap_guard and input_t < 0:
                self.input_priorities[input_t:] = float(self.default_value)
        on = [(low_on_t, high_on_t)] if high_on_t > low_on_t else (
            [(low_on_t, self.T), (0, high_on_t)] if high_on_t < low_on_t else [])
        off = [(low_off_t, high_off_t)] if high_off_t > low_off_t else [(low_off_t, self.T), (0, high_off_t)]
        for (a, z) in on:                                                     # reconstruct_advance :160-185
            n = (z - a) * self.B
            if self.input_priorities is None:
                self._update_segment(n, leaf_base=a * self.B + self.low_idx, scalar=self.default_value)
            else:
                self._update_segment(n, leaf_base=a * self.B + self.low_idx,
                                     values=self.input_priorities[a:z].reshape(-1))
        for (a, z) in off:                                                    # :186-200
            self._update_segment((z - a) * self.B, leaf_base=a * self.B + self.low_idx, scalar=0.0)
        self.t = (t + T) % self.T

    # ---- sampling ------------------------------------------------------------------------------
    def find(self, random_values):
        """sum_tree.py:211-222 -> (tree_idxs, scaled_random_values) as CUDA tensors."""
        u = torch.as_tensor(np.asarray(random_values, dtype=np.float64) if not isinstance(random_values, torch.Tensor)
                            else random_values, dtype=torch.float64).to(self.device)
        n = u.numel()
        idx = torch.empty(n, dtype=torch.int64, device=self.device)
        scaled = torch.empty(n, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            _lib.call("rl_sumtree_find_f64", _lib.ptr(self.tree), self.tree_levels, _lib.ptr(u), n, self.B,
                      _lib.ptr(idx), None, None, None, _lib.ptr(scaled), _lib.stream())
        return idx, scaled

    def sample(self, n, unique=False, random_values=None):
        """sum_tree.py:101-128.  Returns ``((T_idxs, B_idxs), priorities)`` as CUDA tensors."""
        self._sampled_unique = unique
        u_np = np.random.rand(int(n)) if random_values is None else random_values
        if unique:
            return self._sample_unique(n, u_np)
        u = torch.as_tensor(np.asarray(u_np, dtype=np.float64) if not isinstance(u_np, torch.Tensor) else u_np,
                            dtype=torch.float64).to(self.device, non_blocking=True)
        arena = torch.empty(4 * n, dtype=torch.int64, device=self.device)       # one allocation for the four outputs
        idx, T_idxs, B_idxs = arena[:n], arena[n:2 * n], arena[2 * n:3 * n]
        pri = arena[3 * n:].view(torch.float64)
        with torch.cuda.device(self.device):
            _lib.call("rl_sumtree_find_f64", _lib.ptr(self.tree), self.tree_levels, _lib.ptr(u), int(n), self.B,
                      _lib.ptr(idx), _lib.ptr(T_idxs), _lib.ptr(B_idxs), _lib.ptr(pri), None, _lib.stream())
        self.prev_tree_idxs = idx
        return (T_idxs, B_idxs), pri

    def _sample_unique(self, n, u_np):
        """The resampling loop of sum_tree.py:109-123; the de-duplication runs on the host like the
        reference (np.unique), only ``find`` is on the device.  Rare path (``unique=True``)."""
        idx, scaled = (x.cpu().numpy() for x in self.find(u_np))
        i = 0
        while i < 100:
            idx, first = np.unique(idx, return_index=True)
            scaled = scaled[first]
            if len(idx) < n:
                more, more_scaled = (x.cpu().numpy() for x in self.find(np.random.rand(2 * (n - len(idx)))))
                idx = np.concatenate([idx, more])
                scaled = np.concatenate([scaled, more_scaled])
            else:
                break
            i += 1
        if len(idx) < n:
            raise RuntimeError("After 100 tries, unable to get unique indexes.")
        idx_t = torch.from_numpy(idx[:n]).to(self.device)
        self.prev_tree_idxs = idx_t
        leaf = idx_t - self.low_idx
        return (torch.div(leaf, self.B, rounding_mode="floor"), leaf % self.B), self.tree[idx_t]

    BATCH_KERNEL_MAX = 2048        # csrc/sumtree.cu kBatchMax: one CTA sorts the batch in shared memory

    def update_batch_priorities(self, priorities, alpha=None):
        """sum_tree.py:130-138 + reconstruct :150-153.  ``priorities``: CUDA f32/f64 tensor (or numpy),
        aligned with the last ``sample``; duplicates keep the first occurrence.  With ``alpha`` the leaves become
        ``priorities ** alpha`` evaluated as numpy's float32 power (prioritized.py:79) inside the same call.
        Batches up to ``BATCH_KERNEL_MAX`` take the two-launch path (sort + pow + leaf write in one CTA, then the
        ordered propagation); larger ones the generic sorted-segment path."""
        idx = self.prev_tree_idxs
        n = idx.numel()
        if n <= self.BATCH_KERNEL_MAX:
            pri = torch.as_tensor(priorities)
            if alpha is not None:
                pri = pri.to(self.device, dtype=torch.float32).reshape(-1).contiguous()
                p32, p64, a = _lib.ptr(pri), None, float(np.float32(alpha))
            else:
                pri = pri.to(self.device, dtype=torch.float64).reshape(-1).contiguous()
                p32, p64, a = None, _lib.ptr(pri), 0.0
            assert pri.numel() == n, "priorities must align with the last sample()"
            if getattr(self, "_sorted_idx", None) is None or self._sorted_idx.numel() < n:
                self._sorted_idx = torch.empty(max(n, 512), dtype=torch.int64, device=self.device)
            with torch.cuda.device(self.device):
                _lib.call("rl_sumtree_update_batch", _lib.ptr(self.tree), self.tree_levels, _lib.ptr(idx.contiguous()),
                          p32, a, p64, int(n), _lib.ptr(self._sorted_idx), _lib.ptr(self._scratch(n)), _lib.stream(),
                          n_launch=2)
            return
        pri = torch.as_tensor(priorities)
        if alpha is not None:
            pri = _pow_alpha_f64(pri, alpha, self.device)
        pri = pri.to(self.device, dtype=torch.float64).reshape(-1)
        if not self._sampled_unique:
            idx, perm = torch.sort(idx, stable=True)      # stable => first occurrence leads its run
            pri = pri[perm]
            self.prev_tree_idxs = idx
        self._update_segment(idx.numel(), leaf_idx=idx.contiguous(), values=pri.contiguous())

    def print_tree(self, level=None):
        host = self.tree.cpu().numpy()
        for k in (range(self.tree_levels) if level is None else [level]):
            print(" ".join(str(x) for x in host[2 ** k - 1: 2 ** (k + 1) - 1]))
