// Categorical.sample on the device (rlpyt/distributions/categorical.py:25-30 draws with torch.multinomial):
// inverse-CDF draw per row, action = #{k < A-1 : u >= cumsum_fp32(p)[k]} - the definition oracle/pg_loss.py:
// sample_categorical pins - with the uniform either INJECTED (parity tests: bit-exact actions) or generated in the
// kernel by Philox4x32-10 from a (seed, counter) pair that lives in device memory, so that a CUDA graph that
// captured agent.step draws fresh numbers at every replay (the kernel bumps the counter itself).
#include "common.cuh"

namespace rl {

__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
}
// first output word of Philox4x32-10 with counter (row_lo, row_hi, call_lo, call_hi) and key = seed
__device__ __forceinline__ uint32_t philox_word(uint64_t row, uint64_t call, uint64_t seed) {
    uint32_t c[4] = {static_cast<uint32_t>(row), static_cast<uint32_t>(row >> 32), static_cast<uint32_t>(call),
                     static_cast<uint32_t>(call >> 32)};
    uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        philox_round(c, k0, k1);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return c[0];
}

// one block: every thread reads the call counter before anyone bumps it
__global__ void __launch_bounds__(256)
categorical_sample_kernel(const float* __restrict__ prob, const float* __restrict__ uniform, int64_t* __restrict__ rng_state,
                          int64_t* __restrict__ action, float* __restrict__ uniform_out, int64_t N, int A) {
    uint64_t seed = 0, call = 0;
    if (rng_state != nullptr) {
        seed = static_cast<uint64_t>(rng_state[0]);
        call = static_cast<uint64_t>(rng_state[1]);
    }
    __syncthreads();
    for (int64_t i = threadIdx.x; i < N; i += blockDim.x) {
        float u;
        if (uniform != nullptr) u = uniform[i];
        else u = static_cast<float>(philox_word(static_cast<uint64_t>(i), call, seed) >> 8) * (1.0f / 16777216.0f);   // [0, 1), 24 bits
        if (uniform_out != nullptr) uniform_out[i] = u;
        const float* p = prob + i * A;
        float acc = 0.0f;
        int a = 0;
        for (int k = 0; k < A - 1; ++k) {
            acc = __fadd_rn(acc, p[k]);
            a += (u >= acc) ? 1 : 0;
        }
        action[i] = a;
    }
    if (rng_state != nullptr && threadIdx.x == 0) rng_state[1] = static_cast<int64_t>(call + 1);
}

}  // namespace rl

extern "C" {

int rl_categorical_sample_f32(const float* prob, const float* uniform, int64_t* rng_state, int64_t* action,
                              float* uniform_out, int64_t N, int A, void* stream) {
    RL_REQUIRE(prob && action, RL_EINVAL, "rl_categorical_sample_f32: null pointer");
    RL_REQUIRE(N >= 0 && A >= 1, RL_EINVAL, "rl_categorical_sample_f32: bad extents N=%lld A=%d", static_cast<long long>(N), A);
    RL_REQUIRE(uniform != nullptr || rng_state != nullptr, RL_EINVAL,
               "rl_categorical_sample_f32: needs injected uniforms or a device (seed, counter) pair");
    if (N == 0) return RL_OK;
    rl::categorical_sample_kernel<<<1, 256, 0, rl::as_stream(stream)>>>(prob, uniform, rng_state, action, uniform_out, N, A);
    return rl::check_launch("categorical_sample_kernel");
}

}  // extern "C"

// ====================================================================================================
// agent.step's policy head in one launch (rlpyt/models/pg/atari_ff_model.py:56-58 + rlpyt/agents/pg/categorical.py:37-39):
//     pi = softmax(h W_pi^T + b_pi),  v = h w_v + b_v,  action ~ Categorical(pi)   (inverse CDF, Philox as above)
// for the B rows of the sampler's step.  One warp per row: the F features are split over the lanes, the A + 1 dot
// products are reduced with shuffles (fixed tree: deterministic), lane 0 finishes softmax and the draw.  Replaces two
// cuBLAS GEMV-shaped launches, two bias adds, softmax, torch.multinomial's kernels and their glue (~15 launches of a
// latency-bound step) by one.  A <= 32.
namespace rl {

constexpr int kHeadWarps = 8;

__global__ void __launch_bounds__(kHeadWarps * 32)
pg_head_sample_kernel(const float* __restrict__ h, const float* __restrict__ w_pi, const float* __restrict__ b_pi,
                      const float* __restrict__ w_v, const float* __restrict__ b_v, const float* __restrict__ uniform,
                      int64_t* __restrict__ rng_state, float* __restrict__ prob, float* __restrict__ value,
                      int64_t* __restrict__ action, int64_t B, int F, int A) {
    uint64_t seed = 0, call = 0;
    if (rng_state != nullptr) {
        seed = static_cast<uint64_t>(rng_state[0]);
        call = static_cast<uint64_t>(rng_state[1]);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t row = static_cast<int64_t>(blockIdx.x) * kHeadWarps + warp;
    if (row < B) {
        const float* hr = h + row * F;
        float acc[33];
#pragma unroll
        for (int k = 0; k < 33; ++k) acc[k] = 0.0f;
        for (int f = lane; f < F; f += 32) {
            const float x = hr[f];
#pragma unroll
            for (int k = 0; k < 32; ++k)
                if (k < A) acc[k] = fmaf(x, w_pi[k * F + f], acc[k]);
            acc[32] = fmaf(x, w_v[f], acc[32]);
        }
#pragma unroll
        for (int k = 0; k < 33; ++k) {
            if (k < A || k == 32) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
            }
        }
        if (lane == 0) {
            float mx = -3.402823466e38f;
#pragma unroll
            for (int k = 0; k < 32; ++k)
                if (k < A) { acc[k] += b_pi[k]; mx = fmaxf(mx, acc[k]); }
            float sum = 0.0f;
#pragma unroll
            for (int k = 0; k < 32; ++k)
                if (k < A) { acc[k] = expf(acc[k] - mx); sum += acc[k]; }
            const float inv = 1.0f / sum;
            float u;
            if (uniform != nullptr) u = uniform[row];
            else u = static_cast<float>(philox_word(static_cast<uint64_t>(row), call, seed) >> 8) * (1.0f / 16777216.0f);
            float cdf = 0.0f;
            int a = 0;
#pragma unroll
            for (int k = 0; k < 32; ++k)
                if (k < A) {
                    const float p = acc[k] * inv;
                    prob[row * A + k] = p;
                    if (k < A - 1) {
                        cdf = __fadd_rn(cdf, p);
                        a += (u >= cdf) ? 1 : 0;
                    }
                }
            value[row] = acc[32] + b_v[0];
            action[row] = a;
        }
    }
    // the last block to finish bumps the call counter: every block has read it by then
    if (rng_state != nullptr && threadIdx.x == 0) {
        __threadfence();
        const unsigned int done = atomicAdd(reinterpret_cast<unsigned int*>(rng_state + 2), 1u) + 1u;
        if (done == gridDim.x) {
            reinterpret_cast<unsigned int*>(rng_state + 2)[0] = 0u;
            rng_state[1] = static_cast<int64_t>(call + 1);
        }
    }
}

}  // namespace rl

extern "C" {

int rl_pg_head_sample_f32(const float* h, const float* w_pi, const float* b_pi, const float* w_v, const float* b_v,
                          const float* uniform, int64_t* rng_state, float* prob, float* value, int64_t* action, int64_t B,
                          int F, int A, void* stream) {
    RL_REQUIRE(h && w_pi && b_pi && w_v && b_v && prob && value && action, RL_EINVAL, "rl_pg_head_sample_f32: null pointer");
    RL_REQUIRE(B >= 0 && F >= 1 && A >= 1 && A <= 32, RL_EINVAL, "rl_pg_head_sample_f32: needs 1 <= A <= 32 (got B=%lld F=%d A=%d)",
               static_cast<long long>(B), F, A);
    RL_REQUIRE(uniform != nullptr || rng_state != nullptr, RL_EINVAL,
               "rl_pg_head_sample_f32: needs injected uniforms or a device (seed, counter, ticket) triple");
    if (B == 0) return RL_OK;
    const unsigned grid = static_cast<unsigned>((B + rl::kHeadWarps - 1) / rl::kHeadWarps);
    rl::pg_head_sample_kernel<<<grid, rl::kHeadWarps * 32, 0, rl::as_stream(stream)>>>(h, w_pi, b_pi, w_v, b_v, uniform, rng_state, prob,
                                                                                      value, action, B, F, A);
    return rl::check_launch("pg_head_sample_kernel");
}

}  // extern "C"
