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

__global__ void __launch_bounds__(kPhThreads)
pg_heads_fwd_kernel(const float* __restrict__ h, const float* __restrict__ w_pi, const float* __restrict__ b_pi,
                    const float* __restrict__ w_v, const float* __restrict__ b_v, float* __restrict__ prob,
                    float* __restrict__ value, int64_t N, int F, int A) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t row = static_cast<int64_t>(blockIdx.x) * kPhWarps + warp; row < N; row += static_cast<int64_t>(gridDim.x) * kPhWarps) {
        const float* hr = h + row * F;
        float acc[kPhMaxA + 1];
#pragma unroll
        for (int k = 0; k <= kPhMaxA; ++k) acc[k] = 0.0f;
        for (int f = lane; f < F; f += 32) {
            const float x = hr[f];
#pragma unroll
            for (int k = 0; k < kPhMaxA; ++k)
                if (k < A) acc[k] = fmaf(x, w_pi[k * F + f], acc[k]);
            acc[kPhMaxA] = fmaf(x, w_v[f], acc[kPhMaxA]);
        }
#pragma unroll
        for (int k = 0; k <= kPhMaxA; ++k) {
            if (k < A || k == kPhMaxA) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
            }
        }
        if (lane == 0) {
            float mx = -3.402823466e38f;
#pragma unroll
            for (int k = 0; k < kPhMaxA; ++k)
                if (k < A) { acc[k] += b_pi[k]; mx = fmaxf(mx, acc[k]); }
            float sum = 0.0f;
#pragma unroll
            for (int k = 0; k < kPhMaxA; ++k)
                if (k < A) { acc[k] = expf(acc[k] - mx); sum += acc[k]; }
            const float inv = 1.0f / sum;
#pragma unroll
            for (int k = 0; k < kPhMaxA; ++k)
                if (k < A) prob[row * A + k] = acc[k] * inv;
            value[row] = acc[kPhMaxA] + b_v[0];
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

// grad_w_pi [A, F], grad_w_v [F], grad_b_pi [A], grad_b_v [1] = sums over CTAs in fp64, ascending CTA order
__global__ void __launch_bounds__(256)
pg_heads_reduce_kernel(const float* __restrict__ partial_w, const float* __restrict__ partial_b, int n_cta, int F, int A,
                       float* __restrict__ grad_w_pi, float* __restrict__ grad_w_v, float* __restrict__ grad_b_pi,
                       float* __restrict__ grad_b_v) {
    const int A1 = A + 1;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < A1 * F) {
        double s = 0.0;
        for (int c = 0; c < n_cta; ++c) s += static_cast<double>(partial_w[static_cast<int64_t>(c) * A1 * F + i]);
        const int a = i / F, f = i - a * F;
        if (a < A) grad_w_pi[a * F + f] = static_cast<float>(s);
        else grad_w_v[f] = static_cast<float>(s);
    } else if (i < A1 * F + A1) {
        const int a = i - A1 * F;
        double s = 0.0;
        for (int c = 0; c < n_cta; ++c) s += static_cast<double>(partial_b[static_cast<int64_t>(c) * A1 + a]);
        if (a < A) grad_b_pi[a] = static_cast<float>(s);
        else grad_b_v[0] = static_cast<float>(s);
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
    int64_t grid = (N + rl::kPhWarps - 1) / rl::kPhWarps;
    if (grid > 8LL * sms) grid = 8LL * sms;
    rl::pg_heads_fwd_kernel<<<static_cast<unsigned>(grid), rl::kPhThreads, 0, rl::as_stream(stream)>>>(h, w_pi, b_pi, w_v, b_v, prob, value, N, F, A);
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
    const int total = (A + 1) * F + (A + 1);
    rl::pg_heads_reduce_kernel<<<(total + 255) / 256, 256, 0, st>>>(partial_w, partial_b, grid, F, A, grad_w_pi, grad_w_v, grad_b_pi, grad_b_v);
    return rl::check_launch("pg_heads_reduce_kernel");
}

}  // extern "C"
