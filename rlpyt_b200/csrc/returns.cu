// K1-K4: discounted / GAE / n-step returns, valid mask, advantage normalisation.
//
// Reference arithmetic (restated in oracle/returns.py):
//   rlpyt/algos/utils.py:8-21    discount_return
//   rlpyt/algos/utils.py:24-40   generalized_advantage_estimation
//   rlpyt/algos/utils.py:67-101  discount_return_n_step
//   rlpyt/algos/utils.py:104-112 valid_from_done
//   rlpyt/algos/pg/base.py:65-73 advantage normalisation
//
// Layout: [T,B] time-major, element (t,b) at t*B+b, so a warp reading one row touches
// 32*VEC consecutive floats (fully coalesced 128/512 B requests).  Two kernels:
//
//  * returns_stream_kernel  - one thread per VEC(=4|1) columns, sequential backward sweep over
//    t with a U-row register prefetch so >= U*(2*16+4) B per thread are in flight.  HBM-bound
//    at large B: 17 B/element for GAE (r4+v4+done1 read, A4+R4 written), 9 B (13 B with the
//    fused advantage) for discount_return.  Uses the reference's operation order with
//    non-contracted fp32 ops => bit-identical to the reference.
//  * returns_tscan_kernel   - small B (the [128,256] config): T is split over the 32 warps of
//    a CTA (lane = column), each warp composes its R rows into an affine map
//    A_in -> D + C*A_in, the 32 chunk maps are suffix-scanned with warp shuffles through a
//    padded shared-memory transpose, and the carries are applied.  One memory round trip
//    instead of T dependent steps.  Segments (episode ends) need no special casing: a done
//    step has C = 0.  Re-associated => <= 1e-5 relative, not bit-exact.
#include "common.cuh"

namespace rl {

// ------------------------------------------------------------------ streaming kernel
template <int VEC> struct VecF;
template <> struct VecF<4> {
    using F = float4;
    using D = uint32_t;  // 4 packed done bytes
};
template <> struct VecF<1> {
    using F = float;
    using D = uint8_t;
};

template <int VEC>
__device__ __forceinline__ void unpack(const typename VecF<VEC>::F& f, float (&o)[VEC]);
template <>
__device__ __forceinline__ void unpack<4>(const float4& f, float (&o)[4]) {
    o[0] = f.x; o[1] = f.y; o[2] = f.z; o[3] = f.w;
}
template <>
__device__ __forceinline__ void unpack<1>(const float& f, float (&o)[1]) { o[0] = f; }

template <int VEC>
__device__ __forceinline__ typename VecF<VEC>::F pack(const float (&o)[VEC]);
template <>
__device__ __forceinline__ float4 pack<4>(const float (&o)[4]) { return make_float4(o[0], o[1], o[2], o[3]); }
template <>
__device__ __forceinline__ float pack<1>(const float (&o)[1]) { return o[0]; }

template <int VEC>
__device__ __forceinline__ void unpack_nd(typename VecF<VEC>::D d, float (&nd)[VEC]);
template <>
__device__ __forceinline__ void unpack_nd<4>(uint32_t d, float (&nd)[4]) {
#pragma unroll
    for (int k = 0; k < 4; ++k) nd[k] = __fsub_rn(1.0f, static_cast<float>((d >> (8 * k)) & 0xffu));
}
template <>
__device__ __forceinline__ void unpack_nd<1>(uint8_t d, float (&nd)[1]) {
    nd[0] = __fsub_rn(1.0f, static_cast<float>(d));
}

constexpr int kStreamThreads = 128;
constexpr int kStreamU = 4;  // rows prefetched per batch

template <int VEC, bool GAE, bool WITH_VALUE>
__global__ void __launch_bounds__(kStreamThreads)
returns_stream_kernel(const float* __restrict__ reward, const float* __restrict__ value,
                      const uint8_t* __restrict__ done, const float* __restrict__ bootstrap,
                      float* __restrict__ adv, float* __restrict__ ret,
                      int T, int64_t B, float g, float gl) {
    using F = typename VecF<VEC>::F;
    using D = typename VecF<VEC>::D;
    const int64_t col = (static_cast<int64_t>(blockIdx.x) * kStreamThreads + threadIdx.x) * VEC;
    if (col >= B) return;

    float nxt[VEC];  // GAE: value[t+1];  discount_return: return_[t+1]
    float na[VEC];   // GAE: advantage[t+1]
    unpack<VEC>(*reinterpret_cast<const F*>(bootstrap + col), nxt);
#pragma unroll
    for (int k = 0; k < VEC; ++k) na[k] = 0.0f;
    bool first = true;

    for (int t0 = T - 1; t0 >= 0; t0 -= kStreamU) {
        F fr[kStreamU], fv[kStreamU];
        D fd[kStreamU];
        // Issue the whole batch of independent loads before touching the dependent chain.
#pragma unroll
        for (int u = 0; u < kStreamU; ++u) {
            const int t = t0 - u;
            if (t >= 0) {
                const int64_t off = static_cast<int64_t>(t) * B + col;
                fr[u] = ldg_stream(reinterpret_cast<const F*>(reward + off));
                if (GAE || WITH_VALUE) fv[u] = ldg_stream(reinterpret_cast<const F*>(value + off));
                fd[u] = ldg_stream(reinterpret_cast<const D*>(done + off));
            }
        }
#pragma unroll
        for (int u = 0; u < kStreamU; ++u) {
            const int t = t0 - u;
            if (t < 0) break;
            const int64_t off = static_cast<int64_t>(t) * B + col;
            float r[VEC], v[VEC], nd[VEC], oa[VEC], orr[VEC];
            unpack<VEC>(fr[u], r);
            if (GAE || WITH_VALUE) unpack<VEC>(fv[u], v);
            unpack_nd<VEC>(fd[u], nd);
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                if (GAE) {
                    // utils.py:35,37-38 (first row: utils.py:35 has no lambda term)
                    const float delta = __fsub_rn(__fadd_rn(r[k], __fmul_rn(__fmul_rn(g, nxt[k]), nd[k])), v[k]);
                    const float a = first ? delta
                                          : __fadd_rn(delta, __fmul_rn(__fmul_rn(gl, nd[k]), na[k]));
                    oa[k] = a;
                    orr[k] = __fadd_rn(a, v[k]);  // utils.py:39
                    na[k] = a;
                    nxt[k] = v[k];
                } else {
                    // utils.py:18,20
                    const float R = __fadd_rn(r[k], __fmul_rn(__fmul_rn(g, nxt[k]), nd[k]));
                    orr[k] = R;
                    if (WITH_VALUE) oa[k] = __fsub_rn(R, v[k]);  // pg/base.py:55
                    nxt[k] = R;
                }
            }
            first = false;
            stg_stream(reinterpret_cast<F*>(ret + off), pack<VEC>(orr));
            if (GAE || WITH_VALUE) stg_stream(reinterpret_cast<F*>(adv + off), pack<VEC>(oa));
        }
    }
}

// ------------------------------------------------------------------ T-parallel scan kernel
constexpr int kScanWarps = 32;

template <int R, bool GAE, bool WITH_VALUE>
__global__ void __launch_bounds__(kScanWarps * 32)
returns_tscan_kernel(const float* __restrict__ reward, const float* __restrict__ value,
                     const uint8_t* __restrict__ done, const float* __restrict__ bootstrap,
                     float* __restrict__ adv, float* __restrict__ ret,
                     int T, int64_t B, float g, float gl) {
    __shared__ float sC[kScanWarps][33];
    __shared__ float sD[kScanWarps][33];
    __shared__ float sCarry[kScanWarps][33];

    const int lane = threadIdx.x & 31;
    const int w = threadIdx.x >> 5;
    const int64_t col = static_cast<int64_t>(blockIdx.x) * 32 + lane;
    const bool colok = col < B;
    const int t0 = w * R;

    float r[R], v[R], nd[R];
#pragma unroll
    for (int i = 0; i < R; ++i) {
        const int t = t0 + i;
        const bool ok = colok && t < T;
        const int64_t off = static_cast<int64_t>(t) * B + col;
        r[i] = ok ? ldg_stream(reward + off) : 0.0f;
        v[i] = (ok && (GAE || WITH_VALUE)) ? ldg_stream(value + off) : 0.0f;
        nd[i] = ok ? 1.0f - static_cast<float>(ldg_stream(done + off)) : 1.0f;
    }
    float vnext = 0.0f;  // value[t0+R] (GAE): first row of the next chunk, or bootstrap at the end
    if (GAE && colok && t0 < T) {
        vnext = (t0 + R < T) ? ldg_stream(value + static_cast<int64_t>(t0 + R) * B + col)
                             : bootstrap[col];
    }

    // Local backward composition: A_t = Dt[i] + Ct[i] * A_in, A_in = A at the first row of the
    // next chunk.  Rows t >= T are the identity map.
    float Ct[R], Dt[R];
    float C = 1.0f, Dd = 0.0f;
#pragma unroll
    for (int i = R - 1; i >= 0; --i) {
        const int t = t0 + i;
        if (t < T) {
            float c, delta;
            if (GAE) {
                const float vn = (i == R - 1 || t == T - 1) ? vnext : v[(i + 1) % R];
                delta = (r[i] + (g * vn) * nd[i]) - v[i];
                c = gl * nd[i];
            } else {
                delta = r[i];
                c = g * nd[i];
            }
            Dd = delta + c * Dd;
            C = c * C;
        }
        Ct[i] = C;
        Dt[i] = Dd;
    }
    sC[w][lane] = C;
    sD[w][lane] = Dd;
    __syncthreads();

    // Cross-chunk suffix scan: warp j owns column j of the tile, lane i = chunk i.
    {
        const int j = w;
        float c = sC[lane][j], d = sD[lane][j];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const float c2 = __shfl_down_sync(0xffffffffu, c, o);
            const float d2 = __shfl_down_sync(0xffffffffu, d, o);
            if (lane + o < 32) {
                d = d + c * d2;
                c = c * c2;
            }
        }
        // Terminal value beyond the last row: 0 for GAE (utils.py:35), bootstrap for the
        // plain discounted return (utils.py:18).
        const int64_t cj = static_cast<int64_t>(blockIdx.x) * 32 + j;
        const float term = (!GAE && cj < B) ? bootstrap[cj] : 0.0f;
        const float cn = __shfl_down_sync(0xffffffffu, c, 1);
        const float dn = __shfl_down_sync(0xffffffffu, d, 1);
        sCarry[lane][j] = (lane == 31) ? term : dn + cn * term;
    }
    __syncthreads();

    const float carry = sCarry[w][lane];
#pragma unroll
    for (int i = 0; i < R; ++i) {
        const int t = t0 + i;
        if (colok && t < T) {
            const int64_t off = static_cast<int64_t>(t) * B + col;
            const float a = Dt[i] + Ct[i] * carry;
            if (GAE) {
                stg_stream(adv + off, a);
                stg_stream(ret + off, a + v[i]);
            } else {
                stg_stream(ret + off, a);
                if (WITH_VALUE) stg_stream(adv + off, a - v[i]);
            }
        }
    }
}

template <bool GAE, bool WITH_VALUE>
static int launch_returns(const float* reward, const float* value, const uint8_t* done,
                          const float* bootstrap, float* adv, float* ret, int T, int64_t B,
                          float g, float gl, int algo, cudaStream_t st) {
    if (algo == 0) {
        // Small batches are latency bound: fan T out over warps.  Large batches are HBM
        // bound: stream.  Crossover measured on B200 (profiles/).
        algo = (T >= 8 && T <= 8 * kScanWarps && B <= 16384) ? 2 : 1;
    }
    if (algo == 2) {
        RL_REQUIRE(T <= 8 * kScanWarps, RL_EINVAL, "tscan kernel supports T <= %d (got %d)",
                   8 * kScanWarps, T);
        const unsigned grid = static_cast<unsigned>((B + 31) / 32);
        const int need = (T + kScanWarps - 1) / kScanWarps;
#define RL_TSCAN(R_)                                                                         \
    returns_tscan_kernel<R_, GAE, WITH_VALUE><<<grid, kScanWarps * 32, 0, st>>>(             \
        reward, value, done, bootstrap, adv, ret, T, B, g, gl)
        if (need <= 1) RL_TSCAN(1);
        else if (need <= 2) RL_TSCAN(2);
        else if (need <= 4) RL_TSCAN(4);
        else RL_TSCAN(8);
#undef RL_TSCAN
        return check_launch("returns_tscan_kernel");
    }
    const bool vec4 = (B % 4 == 0) && aligned(reward, 16) && aligned(ret, 16) && aligned(done, 4) &&
                      aligned(bootstrap, 16) && (!(GAE || WITH_VALUE) || (aligned(value, 16) && aligned(adv, 16)));
    if (vec4) {
        const int64_t threads = B / 4;
        const unsigned grid = static_cast<unsigned>((threads + kStreamThreads - 1) / kStreamThreads);
        returns_stream_kernel<4, GAE, WITH_VALUE><<<grid, kStreamThreads, 0, st>>>(
            reward, value, done, bootstrap, adv, ret, T, B, g, gl);
    } else {
        const unsigned grid = static_cast<unsigned>((B + kStreamThreads - 1) / kStreamThreads);
        returns_stream_kernel<1, GAE, WITH_VALUE><<<grid, kStreamThreads, 0, st>>>(
            reward, value, done, bootstrap, adv, ret, T, B, g, gl);
    }
    return check_launch("returns_stream_kernel");
}

// ------------------------------------------------------------------ n-step return
// One thread per output element; the n taps are read straight from L2 (n <= ~5).
__global__ void nstep_return_kernel(const float* __restrict__ reward, const uint8_t* __restrict__ done,
                                    const float* __restrict__ gpow, float* __restrict__ ret,
                                    uint8_t* __restrict__ done_n, int T_in, int64_t B, int rlen,
                                    int n_step) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= static_cast<int64_t>(rlen) * B) return;
    const int t = static_cast<int>(i / B);
    float acc = reward[i];        // utils.py:82
    uint8_t dn = done[i] ? 1 : 0; // utils.py:83
    for (int n = 1; n < n_step; ++n) {
        const int tt = t + n;
        if (tt >= T_in) break;  // do_truncated: later taps fall off the end (utils.py:93-95)
        const int64_t j = i + static_cast<int64_t>(n) * B;
        const float nd = __fsub_rn(1.0f, static_cast<float>(dn));
        acc = __fadd_rn(acc, __fmul_rn(__fmul_rn(gpow[n], reward[j]), nd));  // utils.py:97
        dn = (dn | (done[j] ? 1 : 0));                                        // utils.py:98
    }
    ret[i] = acc;
    done_n[i] = dn;
}

// ------------------------------------------------------------------ valid_from_done
__global__ void valid_from_done_kernel(const uint8_t* __restrict__ done, float* __restrict__ valid,
                                       int T, int64_t B) {
    const int64_t col = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (col >= B) return;
    float v = 1.0f;
    for (int t = 0; t < T; ++t) {
        const int64_t off = static_cast<int64_t>(t) * B + col;
        valid[off] = v;              // the step on which done fires is still valid (utils.py:111)
        if (done[off]) v = 0.0f;
    }
}

// ------------------------------------------------------------------ advantage normalisation
// Pass 1: per-block (sum, sum of squares, count) in fp64 -> scratch[3*blk..]; fixed order.
// Pass 2: every block folds the partials in the same order, then rescales its slice.
constexpr int kNormThreads = 256;
constexpr int kNormMaxBlocks = 1024;

static inline int norm_blocks(int64_t n) {
    int64_t b = (n + kNormThreads * 4 - 1) / (kNormThreads * 4);
    if (b < 1) b = 1;
    if (b > kNormMaxBlocks) b = kNormMaxBlocks;
    return static_cast<int>(b);
}

__global__ void __launch_bounds__(kNormThreads)
adv_stats_kernel(const float* __restrict__ adv, const float* __restrict__ valid, int64_t n,
                 double* __restrict__ partials) {
    __shared__ double sh[3][kNormThreads / 32];
    double s = 0.0, q = 0.0, c = 0.0;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * kNormThreads + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * kNormThreads) {
        const bool ok = valid == nullptr || valid[i] > 0.0f;  // pg/base.py:67
        if (ok) {
            const double a = static_cast<double>(adv[i]);
            s += a;
            q += a * a;
            c += 1.0;
        }
    }
    s = warp_sum(s); q = warp_sum(q); c = warp_sum(c);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) { sh[0][w] = s; sh[1][w] = q; sh[2][w] = c; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double ts = 0, tq = 0, tc = 0;
        for (int k = 0; k < kNormThreads / 32; ++k) { ts += sh[0][k]; tq += sh[1][k]; tc += sh[2][k]; }
        partials[3 * blockIdx.x + 0] = ts;
        partials[3 * blockIdx.x + 1] = tq;
        partials[3 * blockIdx.x + 2] = tc;
    }
}

__global__ void __launch_bounds__(kNormThreads)
adv_apply_kernel(float* __restrict__ adv, int64_t n, const double* __restrict__ partials,
                 int nparts, float* __restrict__ stats_out) {
    __shared__ float sh_mean, sh_den;
    if (threadIdx.x == 0) {
        double s = 0, q = 0, c = 0;
        for (int k = 0; k < nparts; ++k) { s += partials[3 * k]; q += partials[3 * k + 1]; c += partials[3 * k + 2]; }
        const double mean = s / c;
        double var = (q - s * mean) / (c - 1.0);  // unbiased (torch.std default)
        if (var < 0.0) var = 0.0;
        const float mean_f = static_cast<float>(mean);
        const float std_f = static_cast<float>(sqrt(var));
        sh_mean = mean_f;
        sh_den = fmaxf(std_f, 1e-6f);  // pg/base.py:73 (NaN std propagates like python max())
        if (std_f != std_f) sh_den = std_f;
        if (stats_out != nullptr && blockIdx.x == 0) { stats_out[0] = mean_f; stats_out[1] = std_f; stats_out[2] = static_cast<float>(c); }
    }
    __syncthreads();
    const float mean = sh_mean, den = sh_den;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * kNormThreads + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * kNormThreads) {
        adv[i] = __fdiv_rn(__fsub_rn(adv[i], mean), den);
    }
}

}  // namespace rl

// ============================================================================ C ABI
extern "C" {

int rl_gae_f32(const float* reward, const float* value, const uint8_t* done,
               const float* bootstrap_value, float* advantage, float* return_,
               int T, int64_t B, float discount, float gamma_lambda, int algo, void* stream) {
    RL_REQUIRE(reward && value && done && bootstrap_value && advantage && return_, RL_EINVAL,
               "rl_gae_f32: null pointer");
    RL_REQUIRE(T >= 1 && B >= 1, RL_EINVAL, "rl_gae_f32: T=%d B=%lld must be >= 1", T, (long long)B);
    RL_REQUIRE(algo >= 0 && algo <= 2, RL_EINVAL, "rl_gae_f32: algo=%d", algo);
    return rl::launch_returns<true, true>(reward, value, done, bootstrap_value, advantage, return_,
                                          T, B, discount, gamma_lambda, algo, rl::as_stream(stream));
}

int rl_discount_return_f32(const float* reward, const uint8_t* done, const float* bootstrap_value,
                           const float* value, float* return_, float* advantage,
                           int T, int64_t B, float discount, int algo, void* stream) {
    RL_REQUIRE(reward && done && bootstrap_value && return_, RL_EINVAL,
               "rl_discount_return_f32: null pointer");
    RL_REQUIRE((value == nullptr) == (advantage == nullptr), RL_EINVAL,
               "rl_discount_return_f32: value and advantage must both be given or both be NULL");
    RL_REQUIRE(T >= 1 && B >= 1, RL_EINVAL, "rl_discount_return_f32: T=%d B=%lld", T, (long long)B);
    RL_REQUIRE(algo >= 0 && algo <= 2, RL_EINVAL, "rl_discount_return_f32: algo=%d", algo);
    if (value != nullptr)
        return rl::launch_returns<false, true>(reward, value, done, bootstrap_value, advantage, return_,
                                               T, B, discount, 0.0f, algo, rl::as_stream(stream));
    return rl::launch_returns<false, false>(reward, nullptr, done, bootstrap_value, nullptr, return_,
                                            T, B, discount, 0.0f, algo, rl::as_stream(stream));
}

int rl_nstep_return_f32(const float* reward, const uint8_t* done, const float* discount_pow,
                        float* return_, uint8_t* done_n, int T_in, int64_t B, int n_step,
                        int do_truncated, void* stream) {
    RL_REQUIRE(reward && done && discount_pow && return_ && done_n, RL_EINVAL,
               "rl_nstep_return_f32: null pointer");
    RL_REQUIRE(n_step >= 1 && B >= 1 && T_in >= 1, RL_EINVAL, "rl_nstep_return_f32: bad extent");
    const int rlen = do_truncated ? T_in : T_in - (n_step - 1);
    RL_REQUIRE(rlen >= 1, RL_EINVAL, "rl_nstep_return_f32: T_in=%d too short for n_step=%d", T_in, n_step);
    const int64_t n = static_cast<int64_t>(rlen) * B;
    const unsigned grid = static_cast<unsigned>((n + 255) / 256);
    rl::nstep_return_kernel<<<grid, 256, 0, rl::as_stream(stream)>>>(reward, done, discount_pow, return_,
                                                                     done_n, T_in, B, rlen, n_step);
    return rl::check_launch("nstep_return_kernel");
}

int rl_valid_from_done_f32(const uint8_t* done, float* valid, int T, int64_t B, void* stream) {
    RL_REQUIRE(done && valid, RL_EINVAL, "rl_valid_from_done_f32: null pointer");
    RL_REQUIRE(T >= 1 && B >= 1, RL_EINVAL, "rl_valid_from_done_f32: bad extent");
    const unsigned grid = static_cast<unsigned>((B + 127) / 128);
    rl::valid_from_done_kernel<<<grid, 128, 0, rl::as_stream(stream)>>>(done, valid, T, B);
    return rl::check_launch("valid_from_done_kernel");
}

int64_t rl_adv_normalize_scratch_bytes(int64_t n) {
    if (n < 1) n = 1;
    return static_cast<int64_t>(rl::norm_blocks(n)) * 3 * sizeof(double);
}

int rl_adv_normalize_f32(float* advantage, const float* valid, int64_t n, void* scratch,
                         float* stats_out, void* stream) {
    RL_REQUIRE(advantage && scratch, RL_EINVAL, "rl_adv_normalize_f32: null pointer");
    RL_REQUIRE(n >= 1, RL_EINVAL, "rl_adv_normalize_f32: n=%lld", (long long)n);
    RL_REQUIRE(rl::aligned(scratch, 8), RL_EALIGN, "rl_adv_normalize_f32: scratch must be 8B aligned");
    const int nb = rl::norm_blocks(n);
    double* partials = static_cast<double*>(scratch);
    cudaStream_t st = rl::as_stream(stream);
    rl::adv_stats_kernel<<<nb, rl::kNormThreads, 0, st>>>(advantage, valid, n, partials);
    int rc = rl::check_launch("adv_stats_kernel");
    if (rc != RL_OK) return rc;
    rl::adv_apply_kernel<<<nb, rl::kNormThreads, 0, st>>>(advantage, n, partials, nb, stats_out);
    return rl::check_launch("adv_apply_kernel");
}

}  // extern "C"
