"""Oracle: the fp64 sum-tree over a [T,B] priority matrix (numpy, bit-exact restatement).

Test infrastructure only (see oracle/__init__.py).  Reference: rlpyt/replays/sum_tree.py -
layout :39-43 (levels = ceil(log2(size+1))+1, leaves = first T*B slots of the last level viewed
[T,B]), ``advance`` :60-99 with the initial wrap guard :74-83, ``reconstruct_advance`` :155-204,
``propagate_diffs`` :206-209 (``np.add.at``: sequential fp64 adds in array order), ``find``
:211-222 (strict ``>``: r == left goes left), ``sample`` :101-128 (``np.random.rand``; unique
resampling loop :109-123), ``update_batch_priorities`` :130-138 (``np.unique`` keeps the FIRST
duplicate) + ``reconstruct`` :150-153.
"""
import numpy as np


class SumTree:

    def __init__(self, T, B, off_backward, off_forward, default_value=1, enable_input_priorities=False,
                 input_priority_shift=0):
        self.T, self.B, self.size = T, B, T * B
        self.off_backward, self.off_forward = off_backward, off_forward
        self.default_value = default_value
        self.input_priority_shift = input_priority_shift
        self.tree_levels = int(np.ceil(np.log2(self.size + 1)) + 1)          # sum_tree.py:39
        self.tree = np.zeros(2 ** self.tree_levels - 1, dtype=np.float64)    # :51
        self.low_idx = 2 ** (self.tree_levels - 1) - 1                       # :41
        self.high_idx = self.size + self.low_idx
        self.priorities = self.tree[self.low_idx:self.high_idx].reshape(T, B)
        self.input_priorities = (default_value * np.ones((T, B)) if enable_input_priorities else None)
        self.t = 0
        self._initial_wrap_guard = True
        self._sampled_unique = False
        self.prev_tree_idxs = None

    # -- cursor advance ---------------------------------------------------------------------------
    def advance_ranges(self, T):
        """(low_on, high_on, low_off, high_off) exactly as sum_tree.py:72-83; also flips the guard."""
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
        return low_on_t, high_on_t, low_off_t, high_off_t

    def advance(self, T, priorities=None):
        if T == 0:
            return
        guard_before = self._initial_wrap_guard
        t = self.t
        lo_on, hi_on, lo_off, hi_off = self.advance_ranges(T)
        if priorities is not None:                                            # :84-95
            assert self.input_priorities is not None
            input_t = t - self.input_priority_shift
            if input_t < 0 or input_t + T > self.T:
                idxs = np.arange(input_t, input_t + T) % self.T
            else:
                idxs = slice(input_t, input_t + T)
            self.input_priorities[idxs] = priorities
            # :94 tests the guard AFTER the range computation updated it
            if self._initial_wrap_guard and input_t < 0:
                self.input_priorities[input_t:] = self.default_value
        del guard_before
        self._write_ranges(lo_on, hi_on, lo_off, hi_off)
        self.t = (t + T) % self.T

    def _segments(self, lo_t, hi_t, wrap_if_not_greater):
        """Leaf index segments for a [lo_t, hi_t) time range, possibly wrapped (:159-200)."""
        if hi_t > lo_t:
            return [(lo_t, hi_t)]
        if hi_t < lo_t or wrap_if_not_greater:
            return [(lo_t, self.T), (0, hi_t)]
        return []

    def _write_ranges(self, lo_on, hi_on, lo_off, hi_off):
        idxs, diffs = [], []
        for (a, z) in self._segments(lo_on, hi_on, wrap_if_not_greater=False):     # on: nothing if equal
            new = self.default_value if self.input_priorities is None else self.input_priorities[a:z]
            diffs.append((new - self.priorities[a:z]).reshape(-1))
            self.priorities[a:z] = new
            idxs.append(np.arange(a * self.B + self.low_idx, z * self.B + self.low_idx))
        for (a, z) in self._segments(lo_off, hi_off, wrap_if_not_greater=True):    # off: else-branch wraps
            diffs.append((-self.priorities[a:z]).reshape(-1))
            self.priorities[a:z] = 0
            idxs.append(np.arange(a * self.B + self.low_idx, z * self.B + self.low_idx))
        if diffs:
            self._propagate(np.concatenate(idxs), np.concatenate(diffs))

    def _propagate(self, tree_idxs, diffs):
        for _ in range(1, self.tree_levels):                                  # :206-209
            tree_idxs = (tree_idxs - 1) // 2
            np.add.at(self.tree, tree_idxs, diffs)

    # -- sampling ---------------------------------------------------------------------------------
    def find(self, random_values):
        r = self.tree[0] * np.asarray(random_values, dtype=np.float64)        # :213
        scaled = r.copy()
        idx = np.zeros(len(r), dtype=np.int64)
        for _ in range(self.tree_levels - 1):
            idx = 2 * idx + 1
            left = self.tree[idx]
            right = r > left                                                  # strict (:219)
            idx = idx + right
            r = np.where(right, r - left, r)
        return idx, scaled

    def sample(self, n, unique=False, random_values=None):
        """``random_values``: inject the uniforms (else ``np.random.rand`` like :107)."""
        self._sampled_unique = unique
        u = np.random.rand(int(n)) if random_values is None else np.asarray(random_values, dtype=np.float64)
        tree_idxs, scaled = self.find(u)
        if unique:                                                            # :109-123
            i = 0
            while i < 100:
                tree_idxs, first = np.unique(tree_idxs, return_index=True)
                scaled = scaled[first]
                if len(tree_idxs) < n:
                    more, more_scaled = self.find(np.random.rand(2 * (n - len(tree_idxs))))
                    tree_idxs = np.concatenate([tree_idxs, more])
                    scaled = np.concatenate([scaled, more_scaled])
                else:
                    break
                i += 1
            if len(tree_idxs) < n:
                raise RuntimeError("After 100 tries, unable to get unique indexes.")
            tree_idxs = tree_idxs[:n]
        priorities = self.tree[tree_idxs]
        self.prev_tree_idxs = tree_idxs
        T_idxs, B_idxs = np.divmod(tree_idxs - self.low_idx, self.B)          # :127
        return (T_idxs, B_idxs), priorities

    def update_batch_priorities(self, priorities):
        priorities = np.asarray(priorities)
        if not self._sampled_unique:                                          # :133-137
            self.prev_tree_idxs, first = np.unique(self.prev_tree_idxs, return_index=True)
            priorities = priorities[first]
        diffs = priorities - self.tree[self.prev_tree_idxs]                   # :151 (upcast to f64)
        self.tree[self.prev_tree_idxs] = priorities
        self._propagate(self.prev_tree_idxs, diffs)
