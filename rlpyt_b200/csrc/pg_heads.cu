// Policy + value heads of the policy-gradient models in training (rlpyt/models/pg/atari_ff_model.py:56-61):
//     pi = softmax(h W_pi^T + b_pi),   v = h w_v^T + b_v          h: [N, F] (fc output), W_pi: [A, F], w_v: [1, F]
// forward in ONE kernel and backward in ONE kernel + a small fixed-order reduction, instead of what torch.nn.Linear x 2 +
// softmax cost per update at N = 8192, F = 512, A = 6 (ncu launch list, profiles/r02_launches_ppo_iter.csv): five
// cuBLAS/CUTLASS SIMT GEMMs and GEMVs of 7-22 us each (the heads are 6 and 1 columns wide: no tile shape fits), two
// split-K reductions, two bias reductions, softmax forward/backward and the add of the two input gradients - about
// 120 us of launches for 34 MB of traffic.  Here h is read once per pass.
//
// forward: one warp per row (the arithmetic of pg_head_sample_kernel, categorical.cu, without the draw).
// backward: a CTA owns a contiguous block of rows.  Per 8 rows its warps first turn (dL/dpi, pi, dL/dv) into the logit
// gradients gl[a] = pi_a (g_a - sum_k g_k pi_k) (softmax backward) in shared memory; then every thread, which owns the
// feature columns f = t, t + 256, ... of all A + 1 weight rows in registers, (i) writes dL/dh[r, f] = sum_a gl[a] W[a, f]
// and (ii) accumulates dL/dW[a, f] += gl[a] h[r, f].  Per-CTA partial sums are added over CTAs in fp64 in a fixed order
// (deterministic; every rank of a data-parallel job computes bit-identical updates from identical data).
#include "common.cuh"

namespace rl {

constexpr int kPhWarps = 8;
constexpr int kPhThreads = kPhWarps * 32;
constexpr int kPhMaxA = 32;
constexpr int kPhMaxCols = 4;                    // feature columns per thread in the backward: F <= 1024

// Forward.  The A + 1 weight rows are staged in shared memory once per CTA ([A + 1][F] floats: 14 KB at A = 6, F = 512);
// a warp takes rows grid-stride, every lane 4 consecutive features per step (16-byte loads of h, conflict-free 16-byte
// reads of the weights), AMAX = the action count's bucket so that only A + 1 (not 33) accumulators are live.
template <int AMAX>
__global__ void __launch_bounds__(kPhThreads)
pg_heads_fwd_kernel(const float* __restrict__ h, const float* __restrict__ w_pi, const float* __restrict__ b_pi,
                    const float* __restrict__ w_v, const float* __restrict__ b_v, float* __restrict__ prob,
                    float* __restrict__ value, int64_t N, int F, int A) {
    extern __shared__ __align__(16) float ws[];              // [A + 1][F]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int A1 = A + 1;
    for (int i = threadIdx.x; i < A1 * F; i += kPhThreads) ws[i] = i < A * F ? w_pi[i] : w_v[i - A * F];
    __syncthreads();
    const bool vec = (F % 4 == 0) && ((reinterpret_cast<uintptr_t>(h) & 15) == 0);
    for (int64_t row = static_cast<int64_t>(blockIdx.x) * kPhWarps + warp; row < N; row += static_cast<int64_t>(gridDim.x) * kPhWarps) {
        const float* hr = h + row * F;
        float acc[AMAX + 1];
#pragma unroll
        for (int k = 0; k <= AMAX; ++k) acc[k] = 0.0f;
        if (vec) {
            for (int f = 4 * lane; f < F; f += 128) {
                const float4 x = *reinterpret_cast<const float4*>(hr + f);
#pragma unroll
                for (int k = 0; k <= AMAX; ++k)
                    if (k < A1) {
                        const float4 w = *reinterpret_cast<const float4*>(ws + k * F + f);
                        acc[k] = fmaf(x.x, w.x, fmaf(x.y, w.y, fmaf(x.z, w.z, fmaf(x.w, w.w, acc[k]))));
                    }
            }
        } else {
            for (int f = lane; f < F; f += 32) {
                const float x = hr[f];
#pragma unroll
                for (int k = 0; k <= AMAX; ++k)
                    if (k < A1) acc[k] = fmaf(x, ws[k * F + f], acc[k]);
            }
        }
#pragma unroll
        for (int k = 0; k <= AMAX; ++k)
            if (k < A1) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
            }
        if (lane == 0) {
            float mx = -3.402823466e38f, vhead = 0.0f;
#pragma unroll
            for (int k = 0; k <= AMAX; ++k) {
                if (k < A) { acc[k] += b_pi[k]; mx = fmaxf(mx, acc[k]); }
                if (k == A) vhead = acc[k];
            }
            float sum = 0.0f;
#pragma unroll
            for (int k = 0; k <= AMAX; ++k)
                if (k < A) { acc[k] = expf(acc[k] - mx); sum += acc[k]; }
            const float inv = 1.0f / sum;
#pragma unroll
            for (int k = 0; k <= AMAX; ++k)
                if (k < A) prob[row * A + k] = acc[k] * inv;
            value[row] = vhead + b_v[0];
        }
    }
}

// partial_w: [grid][A + 1][F] (row A = the value head), partial_b: [grid][A + 1]
template <int COLS, int AMAX>
__global__ void __launch_bounds__(kPhThreads)
pg_heads_bwd_kernel(const float* __restrict__ h, const float* __restrict__ prob, const float* __restrict__ g_prob,
                    const float* __restrict__ g_value, const float* __restrict__ w_pi, const float* __restrict__ w_v,
                    float* __restrict__ grad_h, float* __restrict__ partial_w, float* __restrict__ partial_b, int64_t N, int F,
                    int A, int rows_per_cta) {
    __shared__ float gl[kPhWarps][kPhMaxA + 1];
    __shared__ float bsum[kPhMaxA + 1];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int A1 = A + 1;
    float w[AMAX + 1][COLS], acc[AMAX + 1][COLS];       // AMAX >= A: register arrays sized for the action count's bucket
#pragma unroll
    for (int a = 0; a <= AMAX; ++a)
#pragma unroll
        for (int c = 0; c < COLS; ++c) {
            const int f = tid + c * kPhThreads;
            acc[a][c] = 0.0f;
            w[a][c] = (a < A1 && f < F) ? (a < A ? w_pi[a * F + f] : w_v[f]) : 0.0f;
        }
    if (tid <= kPhMaxA) bsum[tid] = 0.0f;
    const int64_t r0 = static_cast<int64_t>(blockIdx.x) * rows_per_cta;
    const int64_t r1 = min(N, r0 + rows_per_cta);
    for (int64_t rb = r0; rb < r1; rb += kPhWarps) {
        __syncthreads();                                     // gl of the previous batch has been consumed
        const int64_t row = rb + warp;
        if (row < r1) {
            // softmax backward for this warp's row: lane k < A holds (g_k, p_k)
            const float p = lane < A ? prob[row * A + lane] : 0.0f;
            const float g = (lane < A && g_prob != nullptr) ? g_prob[row * A + lane] : 0.0f;
            float dot = g * p;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
            if (lane < A) gl[warp][lane] = p * (g - dot);
            if (lane == 0) gl[warp][A] = g_value != nullptr ? g_value[row] : 0.0f;
        } else {
            if (lane < A) gl[warp][lane] = 0.0f;
            if (lane == 0) gl[warp][A] = 0.0f;
        }
        __syncthreads();
        if (tid < A1) {                                      // bias gradients: fixed order over the rows of the batch
            float s = bsum[tid];
#pragma unroll
            for (int j = 0; j < kPhWarps; ++j) s += gl[j][tid];
            bsum[tid] = s;
        }
#pragma unroll
        for (int j = 0; j < kPhWarps; ++j) {
            const int64_t r = rb + j;
            if (r >= r1) break;
            float x[COLS], gh[COLS];
#pragma unroll
            for (int c = 0; c < COLS; ++c) {
                const int f = tid + c * kPhThreads;
                x[c] = f < F ? h[r * F + f] : 0.0f;
                gh[c] = 0.0f;
            }
#pragma unroll
            for (int a = 0; a <= AMAX; ++a) {
                if (a < A1) {
                    const float ga = gl[j][a];
#pragma unroll
                    for (int c = 0; c < COLS; ++c) {
                        gh[c] = fmaf(ga, w[a][c], gh[c]);
                        acc[a][c] = fmaf(ga, x[c], acc[a][c]);
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < COLS; ++c) {
                const int f = tid + c * kPhThreads;
                if (f < F) grad_h[r * F + f] = gh[c];
            }
        }
    }
    __syncthreads();
    float* pw = partial_w + static_cast<int64_t>(blockIdx.x) * A1 * F;
#pragma unroll
    for (int a = 0; a <= AMAX; ++a)
        if (a < A1)
#pragma unroll
            for (int c = 0; c < COLS; ++c) {
                const int f = tid + c * kPhThreads;
                if (f < F) pw[a * F + f] = acc[a][c];
            }
    if (tid < A1) partial_b[static_cast<int64_t>(blockIdx.x) * A1 + tid] = bsum[tid];
}

// grad_w_pi [A, F], grad_w_v [F], grad_b_pi [A], grad_b_v [1] = sums over CTAs in fp64 in a FIXED order: one warp per
// output, lane l adds the partials of CTAs l, l + 32, ... (ascending), then a fixed shuffle tree combines the 32 lanes.
__global__ void __launch_bounds__(256)
pg_heads_reduce_kernel(const float* __restrict__ partial_w, const float* __restrict__ partial_b, int n_cta, int F, int A,
                       float* __restrict__ grad_w_pi, float* __restrict__ grad_w_v, float* __restrict__ grad_b_pi,
                       float* __restrict__ grad_b_v) {
    const int A1 = A + 1;
    const int lane = threadIdx.x & 31;
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;          // output index, warp-uniform
    const int n_w = A1 * F;
    if (i >= n_w + A1) return;
    const float* src = i < n_w ? partial_w + i : partial_b + (i - n_w);
    const int64_t stride = i < n_w ? static_cast<int64_t>(n_w) : A1;
    double s = 0.0;
    for (int c = lane; c < n_cta; c += 32) s += static_cast<double>(src[static_cast<int64_t>(c) * stride]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
        if (i < n_w) {
            const int a = i / F, f = i - a * F;
            if (a < A) grad_w_pi[a * F + f] = static_cast<float>(s);
            else grad_w_v[f] = static_cast<float>(s);
        } else {
            const int a = i - n_w;
            if (a < A) grad_b_pi[a] = static_cast<float>(s);
            else grad_b_v[0] = static_cast<float>(s);
        }
    }
}

static inline int heads_grid(int64_t N, int* rows_per_cta) {
    int sms = sm_count();
    if (sms <= 0) sms = 148;
    int64_t grid = (N + 31) / 32;                            // ~32 rows per CTA ...
    if (grid > 4LL * sms) grid = 4LL * sms;                  // ... unless that makes more than 4 CTAs per SM
    if (grid < 1) grid = 1;
    int64_t rpc = (N + grid - 1) / grid;
    rpc = (rpc + kPhWarps - 1) / kPhWarps * kPhWarps;
    *rows_per_cta = static_cast<int>(rpc);
    return static_cast<int>((N + rpc - 1) / rpc);
}

}  // namespace rl

extern "C" {

int rl_pg_heads_forward_f32(const float* h, const float* w_pi, const float* b_pi, const float* w_v, const float* b_v,
                            float* prob, float* value, int64_t N, int F, int A, void* stream) {
    RL_REQUIRE(h && w_pi && b_pi && w_v && b_v && prob && value, RL_EINVAL, "rl_pg_heads_forward_f32: null pointer");
    RL_REQUIRE(N >= 0 && F >= 1 && A >= 1 && A <= rl::kPhMaxA, RL_EINVAL, "rl_pg_heads_forward_f32: needs 1 <= A <= 32 (got N=%lld F=%d A=%d)",
               static_cast<long long>(N), F, A);
    if (N == 0) return RL_OK;
    int sms = rl::sm_count();
    if (sms <= 0) sms = 148;
    RL_REQUIRE(static_cast<size_t>(A + 1) * F * sizeof(float) <= 48 * 1024, RL_EINVAL,
               "rl_pg_heads_forward_f32: (A + 1) * F floats must fit 48 KB of shared memory (A=%d F=%d)", A, F);
    int64_t grid = (N + rl::kPhWarps - 1) / rl::kPhWarps;
    if (grid > 4LL * sms) grid = 4LL * sms;                  // each CTA stages the weights once, then takes rows grid-stride
    const size_t smem = static_cast<size_t>(A + 1) * F * sizeof(float);
    cudaStream_t st = rl::as_stream(stream);
    if (A <= 8) rl::pg_heads_fwd_kernel<8><<<static_cast<unsigned>(grid), rl::kPhThreads, smem, st>>>(h, w_pi, b_pi, w_v, b_v, prob, value, N, F, A);
    else if (A <= 16) rl::pg_heads_fwd_kernel<16><<<static_cast<unsigned>(grid), rl::kPhThreads, smem, st>>>(h, w_pi, b_pi, w_v, b_v, prob, value, N, F, A);
    else rl::pg_heads_fwd_kernel<32><<<static_cast<unsigned>(grid), rl::kPhThreads, smem, st>>>(h, w_pi, b_pi, w_v, b_v, prob, value, N, F, A);
    return rl::check_launch("pg_heads_fwd_kernel");
}

int64_t rl_pg_heads_backward_scratch_bytes(int64_t N, int F, int A) {
    if (N < 1 || F < 1 || A < 1) return 0;
    int rpc = 0;
    const int grid = rl::heads_grid(N, &rpc);
    return static_cast<int64_t>(grid) * (A + 1) * (static_cast<int64_t>(F) + 1) * static_cast<int64_t>(sizeof(float));
}

int rl_pg_heads_backward_f32(const float* h, const float* prob, const float* grad_prob, const float* grad_value,
                             const float* w_pi, const float* w_v, float* grad_h, float* grad_w_pi, float* grad_b_pi,
                             float* grad_w_v, float* grad_b_v, int64_t N, int F, int A, void* scratch, void* stream) {
    RL_REQUIRE(h && prob && w_pi && w_v && grad_h && grad_w_pi && grad_b_pi && grad_w_v && grad_b_v && scratch, RL_EINVAL,
               "rl_pg_heads_backward_f32: null pointer");
    RL_REQUIRE(N >= 1 && F >= 1 && F <= rl::kPhMaxCols * rl::kPhThreads && A >= 1 && A <= rl::kPhMaxA, RL_EINVAL,
               "rl_pg_heads_backward_f32: needs N >= 1, F <= %d, 1 <= A <= 32 (got N=%lld F=%d A=%d)", rl::kPhMaxCols * rl::kPhThreads,
               static_cast<long long>(N), F, A);
    RL_REQUIRE(rl::aligned(scratch, 4), RL_EALIGN, "rl_pg_heads_backward_f32: scratch must be 4-byte aligned");
    int rpc = 0;
    const int grid = rl::heads_grid(N, &rpc);
    float* partial_w = static_cast<float*>(scratch);
    float* partial_b = partial_w + static_cast<int64_t>(grid) * (A + 1) * F;
    cudaStream_t st = rl::as_stream(stream);
    const int cols = (F + rl::kPhThreads - 1) / rl::kPhThreads;
#define RL_PH_LAUNCH(C, AM)                                                                                                           \
    rl::pg_heads_bwd_kernel<C, AM><<<static_cast<unsigned>(grid), rl::kPhThreads, 0, st>>>(h, prob, grad_prob, grad_value, w_pi, w_v, \
                                                                                          grad_h, partial_w, partial_b, N, F, A, rpc)
#define RL_PH_COLS(AM)                    \
    do {                                  \
        if (cols <= 1) RL_PH_LAUNCH(1, AM); \
        else if (cols == 2) RL_PH_LAUNCH(2, AM); \
        else RL_PH_LAUNCH(4, AM);         \
    } while (0)
    if (A <= 8) RL_PH_COLS(8);
    else if (A <= 16) RL_PH_COLS(16);
    else RL_PH_COLS(32);
#undef RL_PH_COLS
#undef RL_PH_LAUNCH
    int rc = rl::check_launch("pg_heads_bwd_kernel");
    if (rc != RL_OK) return rc;
    const int total = (A + 1) * F + (A + 1);                 // one warp per output
    rl::pg_heads_reduce_kernel<<<(total + 7) / 8, 256, 0, st>>>(partial_w, partial_b, grid, F, A, grad_w_pi, grad_w_v, grad_b_pi, grad_b_v);
    return rl::check_launch("pg_heads_reduce_kernel");
}

}  // extern "C"
