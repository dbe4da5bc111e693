// K6/K7: the fp64 sum-tree of prioritized replay, bit-exact against the reference.
//
// Reference (restated in oracle/sum_tree.py): rlpyt/replays/sum_tree.py - find :211-222,
// reconstruct :150-153, reconstruct_advance :155-204, propagate_diffs :206-209.
//
// Layout: one implicit binary tree in a single fp64 array of 2^levels - 1 nodes, root at 0,
// children of i at 2i+1 / 2i+2, leaves at [2^(levels-1)-1, ...) (sum_tree.py:39-43).  At the
// 1M-frame config: 21 levels, 2 097 151 nodes = 16.8 MB - L2 resident on B200 (126 MB L2).
//
// * sumtree_find_kernel: one thread per sample, `levels-1` dependent fp64 loads; latency bound
//   (512 x 20 x 8 B = 82 KB of dependent reads per batch).  Uses non-fused __dmul_rn/__dsub_rn and
//   the reference's strict `>` so every sample lands on the same leaf as numpy.
// * update (leaf write + propagate): the reference adds the per-leaf differences to every ancestor
//   with np.add.at, i.e. SEQUENTIAL fp64 adds in array order.  fp64 addition is not associative,
//   and node values are history dependent, so atomics / tree reductions would drift in the last
//   bit and eventually change a sampled index.  Here each (level, node) accumulates its run of
//   differences in array order, and all (level, node) pairs run in parallel:
//     - the caller passes one ASCENDING segment of leaf indices (a contiguous range for `advance`,
//       the sorted batch for `update_batch_priorities`), so the leaves below any node form one
//       contiguous run of the array;
//     - one warp per (level, 32-element chunk): lanes load 32 differences coalesced, run heads are
//       found with a ballot, and the warp walks its runs adding `__shfl_sync`-broadcast values in
//       order (every lane carries the same accumulator); a run that extends past the chunk is
//       continued by its head warp through the following chunks, with the next chunk prefetched.
//     The critical path is the root: n sequential adds (n=512: ~3 us; n=33 k: ~0.2 ms).
//   Duplicate leaves in the batch keep the FIRST value (np.unique(return_index), :135-137): later
//   duplicates are turned into +0.0 differences, which leave every sum bit-identical.
#include "common.cuh"

namespace rl {

__global__ void sumtree_find_kernel(const double* __restrict__ tree, int levels,
                                    const double* __restrict__ uniforms, int64_t n, int64_t low_idx,
                                    int64_t B, int64_t* __restrict__ tree_idx, int64_t* __restrict__ T_idx,
                                    int64_t* __restrict__ B_idx, double* __restrict__ priority,
                                    double* __restrict__ scaled) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double r = __dmul_rn(tree[0], uniforms[i]);           // sum_tree.py:213
    if (scaled != nullptr) scaled[i] = r;
    int64_t idx = 0;
    for (int l = 1; l < levels; ++l) {
        idx = 2 * idx + 1;
        const double left = tree[idx];
        if (r > left) {                                    // strict: r == left goes left (:219)
            idx += 1;
            r = __dsub_rn(r, left);
        }
    }
    tree_idx[i] = idx;
    if (priority != nullptr) priority[i] = tree[idx];
    const int64_t leaf = idx - low_idx;
    if (T_idx != nullptr) T_idx[i] = leaf / B;            // np.divmod (:127)
    if (B_idx != nullptr) B_idx[i] = leaf % B;
}

struct LeafSeg {
    const int64_t* idx;  // ascending leaf tree-indices, or nullptr => base + i
    int64_t base;
    int64_t n;
};

__device__ __forceinline__ int64_t seg_leaf(const LeafSeg& s, int64_t i) {
    return s.idx != nullptr ? s.idx[i] : s.base + i;
}

// diffs[i] = value_i - tree[leaf_i]; tree[leaf_i] = value_i   (reconstruct :151-152 / :167-168,192-193)
__global__ void sumtree_set_leaves_kernel(double* __restrict__ tree, LeafSeg s,
                                          const double* __restrict__ values, double scalar,
                                          double* __restrict__ diffs) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= s.n) return;
    const int64_t leaf = seg_leaf(s, i);
    const bool first = (s.idx == nullptr) || i == 0 || s.idx[i - 1] != leaf;
    if (!first) {  // later duplicate: dropped by np.unique -> contributes +0.0
        diffs[i] = 0.0;
        return;
    }
    const double v = values != nullptr ? values[i] : scalar;
    diffs[i] = __dsub_rn(v, tree[leaf]);
    tree[leaf] = v;
}

constexpr int kPropWarpsPerBlock = 4;

__global__ void __launch_bounds__(kPropWarpsPerBlock * 32)
sumtree_propagate_kernel(double* __restrict__ tree, int levels, LeafSeg s, const double* __restrict__ diffs,
                         int64_t n_chunks) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = static_cast<int64_t>(blockIdx.x) * kPropWarpsPerBlock + (threadIdx.x >> 5);
    const int64_t total = n_chunks * (levels - 1);
    if (warp >= total) return;
    const int up = static_cast<int>(warp / n_chunks) + 1;   // how many levels above the leaves
    const int64_t chunk = warp % n_chunks;
    const int64_t n = s.n;
    constexpr unsigned kFull = 0xffffffffu;

    // ---- this warp's own chunk: find the run heads
    int64_t i = chunk * 32 + lane;
    bool valid = i < n;
    int64_t anc = valid ? ((seg_leaf(s, i) + 1) >> up) - 1 : -2;
    double d = valid ? diffs[i] : 0.0;
    int64_t prev = __shfl_up_sync(kFull, anc, 1);
    if (lane == 0) prev = (i > 0 && valid) ? ((seg_leaf(s, i - 1) + 1) >> up) - 1 : -3;
    const bool head = valid && anc != prev;
    const unsigned head_mask = __ballot_sync(kFull, head);
    if (head_mask == 0) return;  // the whole chunk continues a run owned by an earlier warp
    const unsigned valid_mask = __ballot_sync(kFull, valid);
    // all run heads of the chunk fetch their node concurrently (one memory round trip)
    const double node_val = head ? tree[anc] : 0.0;

    double acc = 0.0;
    int64_t cur = -1;
    const int first = __ffs(head_mask) - 1;
#pragma unroll 1
    for (int k = first; k < 32; ++k) {
        if (!((valid_mask >> k) & 1u)) break;
        if ((head_mask >> k) & 1u) {
            if (cur >= 0 && lane == 0) tree[cur] = acc;   // previous run ended inside the chunk
            cur = __shfl_sync(kFull, anc, k);
            acc = __shfl_sync(kFull, node_val, k);
        }
        acc = __dadd_rn(acc, __shfl_sync(kFull, d, k));   // np.add.at order (:209)
    }

    // ---- the last run may continue through the following chunks.  Upper levels own runs of thousands of
    // differences (the root: all n): the chain of dependent fp64 adds is the critical path of `advance`.  A chunk
    // that lies entirely inside the run is therefore added by EVERY lane from its own registers: the 32 differences
    // are fetched as 16 broadcast 16-byte loads, one chunk ahead of the adds, so the chain is DADD latency only
    // (the first implementation paid a warp shuffle per element: 0.49 ms for a 32 768-leaf segment).
    int64_t next = (chunk + 1) * 32;
    if (valid_mask == kFull && next < n) {
        i = next + lane;
        valid = i < n;
        int64_t a2 = valid ? ((seg_leaf(s, i) + 1) >> up) - 1 : -2;
        double d2 = valid ? diffs[i] : 0.0;
        double2 v2[16];
        if (next + 32 <= n) {
#pragma unroll
            for (int q = 0; q < 16; ++q) v2[q] = reinterpret_cast<const double2*>(diffs + next)[q];
        }
        while (true) {
            // prefetch the chunk after this one
            const int64_t i3 = next + 32 + lane;
            const bool v3ok = i3 < n;
            const int64_t a3 = v3ok ? ((seg_leaf(s, i3) + 1) >> up) - 1 : -2;
            const double d3 = v3ok ? diffs[i3] : 0.0;
            double2 v3[16];
            if (next + 64 <= n) {
#pragma unroll
                for (int q = 0; q < 16; ++q) v3[q] = reinterpret_cast<const double2*>(diffs + next + 32)[q];
            }
            const unsigned same = __ballot_sync(kFull, a2 == cur);
            if (same == kFull) {                                      // implies next + 32 <= n: v2 is loaded
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    acc = __dadd_rn(acc, v2[q].x);                    // np.add.at order (:209)
                    acc = __dadd_rn(acc, v2[q].y);
                }
            } else {
                const int cnt = __ffs(~same) - 1;                     // leading lanes still in the run
                for (int k = 0; k < cnt; ++k) acc = __dadd_rn(acc, __shfl_sync(kFull, d2, k));
                break;
            }
            next += 32;
            if (next >= n) break;
            a2 = a3;
            d2 = d3;
#pragma unroll
            for (int q = 0; q < 16; ++q) v2[q] = v3[q];
        }
    }
    if (lane == 0) tree[cur] = acc;
}

// ---- update_batch_priorities in two launches -------------------------------------------------------------------
// sum_tree.py:130-138 + prioritized.py:73-79 for one sampled batch (n <= kBatchMax): ONE CTA
//   1. sorts the batch by (leaf index, position in the batch) - bitonic sort of 64-bit composite keys in shared
//      memory; the position tie-break makes the FIRST occurrence of a duplicated leaf lead its run, which is the one
//      np.unique(return_index=True) keeps (:135-137);
//   2. evaluates priority ** alpha as numpy's float32 power through fp64 (or takes fp64 values as given);
//   3. writes the leaves and the differences (later duplicates: +0.0) in sorted order,
// then sumtree_propagate_kernel adds the differences to the ancestors in array order as before.  The round trip
// through torch.sort (radix sort: 4-5 launches), the gather by the permutation and the separate pow / set-leaves
// launches of the first implementation cost ~80 us per update; this path is two launches behind one C call.
constexpr int kBatchMax = 2048;
constexpr int kBatchThreads = 1024;

__global__ void __launch_bounds__(kBatchThreads)
sumtree_batch_leaves_kernel(double* __restrict__ tree, const int64_t* __restrict__ leaf_idx,
                            const float* __restrict__ pri_f32, float alpha, const double* __restrict__ values_f64,
                            int n, int padded, int64_t* __restrict__ sorted_idx, double* __restrict__ diffs) {
    __shared__ unsigned long long key[kBatchMax];
    const int tid = threadIdx.x;
    for (int i = tid; i < padded; i += kBatchThreads)
        key[i] = i < n ? ((static_cast<unsigned long long>(leaf_idx[i]) << 11) | static_cast<unsigned long long>(i))
                       : ~0ull;
    __syncthreads();
    for (int k = 2; k <= padded; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int p = tid; p < (padded >> 1); p += kBatchThreads) {
                const int lo = ((p & ~(j - 1)) << 1) | (p & (j - 1));      // index with bit j clear
                const int hi = lo | j;
                const bool up = (lo & k) == 0;
                const unsigned long long a = key[lo], b = key[hi];
                if ((a > b) == up) {
                    key[lo] = b;
                    key[hi] = a;
                }
            }
            __syncthreads();
        }
    }
    for (int i = tid; i < n; i += kBatchThreads) {
        const unsigned long long kv = key[i];
        const int64_t leaf = static_cast<int64_t>(kv >> 11);
        const int pos = static_cast<int>(kv & 2047ull);
        sorted_idx[i] = leaf;
        const bool first = i == 0 || static_cast<int64_t>(key[i - 1] >> 11) != leaf;
        if (!first) {                                                       // dropped by np.unique -> contributes +0.0
            diffs[i] = 0.0;
            continue;
        }
        double v;
        if (values_f64 != nullptr) {
            v = values_f64[pos];
        } else {
            const float r = static_cast<float>(pow(static_cast<double>(pri_f32[pos]), static_cast<double>(alpha)));
            v = static_cast<double>(r);
        }
        diffs[i] = __dsub_rn(v, tree[leaf]);
        tree[leaf] = v;
    }
}

// out[i] = (double)(float)pow((double)x[i], (double)exponent): numpy's float32 `x ** alpha`
// (rlpyt/replays/non_sequence/prioritized.py:79) evaluated through fp64 so that the fp32 result is
// the correctly rounded one (glibc powf is correctly rounded in all but astronomically rare cases).
__global__ void pow_f32_to_f64_kernel(const float* __restrict__ x, float exponent, double* __restrict__ out,
                                      int64_t n) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float r = static_cast<float>(pow(static_cast<double>(x[i]), static_cast<double>(exponent)));
    out[i] = static_cast<double>(r);
}

// is_weights = (1/(p+1e-6))**beta / max(...)  -> float32   (prioritized.py:68-70)
__global__ void is_weights_kernel(const double* __restrict__ priority, double beta, double eps, float* __restrict__ out,
                                  int n) {
    double mx = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) mx = fmax(mx, pow(1.0 / (priority[i] + eps), beta));
    // block max (n <= a few thousand: one block)
    __shared__ double red[32];
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x < 32) {
        double m = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (threadIdx.x == 0) red[0] = m;
    }
    __syncthreads();
    const double m = red[0];
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double w = pow(1.0 / (priority[i] + eps), beta);
        out[i] = static_cast<float>(w / m);
    }
}

}  // namespace rl

extern "C" {

int rl_sumtree_find_f64(const double* tree, int levels, const double* uniforms, int64_t n, int64_t B,
                        int64_t* tree_idx, int64_t* T_idx, int64_t* B_idx, double* priority,
                        double* scaled, void* stream) {
    RL_REQUIRE(tree && uniforms && tree_idx, RL_EINVAL, "rl_sumtree_find_f64: null pointer");
    RL_REQUIRE(levels >= 2 && levels <= 40 && n >= 0 && B >= 1, RL_EINVAL, "rl_sumtree_find_f64: levels=%d n=%lld",
               levels, (long long)n);
    if (n == 0) return RL_OK;
    const int64_t low_idx = (1LL << (levels - 1)) - 1;
    const unsigned grid = static_cast<unsigned>((n + 127) / 128);
    rl::sumtree_find_kernel<<<grid, 128, 0, rl::as_stream(stream)>>>(tree, levels, uniforms, n, low_idx, B,
                                                                      tree_idx, T_idx, B_idx, priority, scaled);
    return rl::check_launch("sumtree_find_kernel");
}

int rl_sumtree_update_f64(double* tree, int levels, const int64_t* leaf_idx, int64_t leaf_base,
                          const double* values, double value_scalar, int64_t n, double* scratch_diffs,
                          void* stream) {
    RL_REQUIRE(tree && scratch_diffs, RL_EINVAL, "rl_sumtree_update_f64: null pointer");
    RL_REQUIRE(levels >= 2 && levels <= 40 && n >= 0, RL_EINVAL, "rl_sumtree_update_f64: levels=%d n=%lld",
               levels, (long long)n);
    RL_REQUIRE(rl::aligned(scratch_diffs, 16), RL_EALIGN, "rl_sumtree_update_f64: scratch_diffs must be 16-byte aligned");
    if (n == 0) return RL_OK;
    const int64_t low_idx = (1LL << (levels - 1)) - 1;
    if (leaf_idx == nullptr)
        RL_REQUIRE(leaf_base >= low_idx && leaf_base + n <= 2 * low_idx + 1, RL_EINVAL,
                   "rl_sumtree_update_f64: leaf range [%lld,%lld) outside the leaf level", (long long)leaf_base,
                   (long long)(leaf_base + n));
    cudaStream_t st = rl::as_stream(stream);
    rl::LeafSeg s{leaf_idx, leaf_base, n};
    rl::sumtree_set_leaves_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(tree, s, values,
                                                                                        value_scalar, scratch_diffs);
    int rc = rl::check_launch("sumtree_set_leaves_kernel");
    if (rc != RL_OK) return rc;
    const int64_t n_chunks = (n + 31) / 32;
    const int64_t warps = n_chunks * (levels - 1);
    const int64_t blocks = (warps + rl::kPropWarpsPerBlock - 1) / rl::kPropWarpsPerBlock;
    RL_REQUIRE(blocks < (1LL << 31), RL_EINVAL, "rl_sumtree_update_f64: segment too large");
    rl::sumtree_propagate_kernel<<<static_cast<unsigned>(blocks), rl::kPropWarpsPerBlock * 32, 0, st>>>(
        tree, levels, s, scratch_diffs, n_chunks);
    return rl::check_launch("sumtree_propagate_kernel");
}

int rl_sumtree_update_batch(double* tree, int levels, const int64_t* leaf_idx, const float* priorities_f32, float alpha,
                            const double* values_f64, int64_t n, int64_t* scratch_sorted_idx, double* scratch_diffs,
                            void* stream) {
    RL_REQUIRE(tree && leaf_idx && scratch_sorted_idx && scratch_diffs, RL_EINVAL, "rl_sumtree_update_batch: null pointer");
    RL_REQUIRE((priorities_f32 != nullptr) != (values_f64 != nullptr), RL_EINVAL,
               "rl_sumtree_update_batch: give exactly one of priorities_f32 (with alpha) and values_f64");
    RL_REQUIRE(levels >= 2 && levels <= 40 && n >= 0 && n <= rl::kBatchMax, RL_EINVAL,
               "rl_sumtree_update_batch: levels=%d n=%lld (batch limit %d)", levels, (long long)n, rl::kBatchMax);
    RL_REQUIRE(rl::aligned(scratch_diffs, 16), RL_EALIGN, "rl_sumtree_update_batch: scratch_diffs must be 16-byte aligned");
    if (n == 0) return RL_OK;
    cudaStream_t st = rl::as_stream(stream);
    int padded = 32;
    while (padded < n) padded <<= 1;
    rl::sumtree_batch_leaves_kernel<<<1, rl::kBatchThreads, 0, st>>>(tree, leaf_idx, priorities_f32, alpha, values_f64,
                                                                   static_cast<int>(n), padded, scratch_sorted_idx,
                                                                   scratch_diffs);
    int rc = rl::check_launch("sumtree_batch_leaves_kernel");
    if (rc != RL_OK) return rc;
    rl::LeafSeg s{scratch_sorted_idx, 0, n};
    const int64_t n_chunks = (n + 31) / 32;
    const int64_t warps = n_chunks * (levels - 1);
    const int64_t blocks = (warps + rl::kPropWarpsPerBlock - 1) / rl::kPropWarpsPerBlock;
    rl::sumtree_propagate_kernel<<<static_cast<unsigned>(blocks), rl::kPropWarpsPerBlock * 32, 0, st>>>(
        tree, levels, s, scratch_diffs, n_chunks);
    return rl::check_launch("sumtree_propagate_kernel");
}

int rl_pow_f32_to_f64(const float* x, float exponent, double* out, int64_t n, void* stream) {
    RL_REQUIRE(x && out && n >= 0, RL_EINVAL, "rl_pow_f32_to_f64: bad argument");
    if (n == 0) return RL_OK;
    rl::pow_f32_to_f64_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, rl::as_stream(stream)>>>(
        x, exponent, out, n);
    return rl::check_launch("pow_f32_to_f64_kernel");
}

int rl_is_weights_f32(const double* priority, double beta, float* out, int n, void* stream) {
    RL_REQUIRE(priority && out && n >= 1, RL_EINVAL, "rl_is_weights_f32: bad argument");
    rl::is_weights_kernel<<<1, 1024, 0, rl::as_stream(stream)>>>(priority, beta, 1e-6, out, n);
    return rl::check_launch("is_weights_kernel");
}

int rl_is_weights_eps_f32(const double* priority, double beta, double eps, float* out, int n, void* stream) {
    RL_REQUIRE(priority && out && n >= 1 && eps >= 0.0, RL_EINVAL, "rl_is_weights_eps_f32: bad argument");
    rl::is_weights_kernel<<<1, 1024, 0, rl::as_stream(stream)>>>(priority, beta, eps, out, n);
    return rl::check_launch("is_weights_kernel");
}

}  // extern "C"
