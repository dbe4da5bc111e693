// K5: fused PPO / A2C loss - forward scalars AND the gradients w.r.t. the network outputs in
// one pass over the minibatch.
//
// Reference arithmetic (restated in oracle/pg_loss.py):
//   rlpyt/algos/pg/ppo.py:136-153      ratio, clip, min, value error, entropy, perplexity
//   rlpyt/algos/pg/a2c.py:88-100       log-likelihood * advantage
//   rlpyt/distributions/categorical.py:32-43, rlpyt/utils/tensor.py:39-46 (valid_mean)
// plus what torch autograd derives from them (minimum: ties split 1/2-1/2; clamp: mask
// lo <= x <= hi inclusive).
//
// The reference runs ~15 tiny torch-CPU ops forward and as many backward per minibatch; here:
//   kernel 1 (pg_loss_kernel): one thread per sample reads its A probabilities once, produces
//     the per-sample surrogate / value error / entropy / perplexity, writes dL/dprob[N,A] and
//     dL/dvalue[N] (already divided by N when there is no valid mask), and reduces the five
//     sums per block in fp64 -> partials (fixed shuffle tree + fixed block order = deterministic);
//   kernel 2 (pg_loss_finalize_kernel): folds the partials in block order, writes the scalars,
//     and - only when a valid mask is given, where the divisor sum(valid) is not known up front -
//     rescales the gradients by 1/sum(valid).
// Traffic per sample: (2A+5)*4 B read (+ valid) and (A+1)*4 B written; N=8192, A=6 is 0.8 MB,
// L2 resident and launch-latency bound - the win over the reference is removing ~30 op
// dispatches, 2 D2H copies and 4 .item() syncs, not bandwidth.
#include "common.cuh"

namespace rl {

constexpr int kLossThreads = 256;
constexpr int kLossSums = 5;  // surrogate|logli*A, value_error, entropy, perplexity, count
constexpr float kEps = 1e-8f; // rlpyt/distributions/categorical.py:9
constexpr int kMaxA = 64;

static inline int loss_blocks(int64_t N) { return static_cast<int>((N + kLossThreads - 1) / kLossThreads); }

template <bool PPO>
__global__ void __launch_bounds__(kLossThreads)
pg_loss_kernel(const float* __restrict__ prob_new, const float* __restrict__ value,
               const float* __restrict__ prob_old, const int64_t* __restrict__ action,
               const float* __restrict__ ret, const float* __restrict__ adv,
               const float* __restrict__ valid, int64_t N, int A, float clip, float c_v, float c_ent,
               float* __restrict__ grad_prob, float* __restrict__ grad_value,
               double* __restrict__ partials, const float* __restrict__ clip_dev) {
    __shared__ double sh[kLossSums][kLossThreads / 32];
    if (PPO && clip_dev != nullptr) clip = *clip_dev;      // ratio clip as a device scalar: a captured CUDA graph follows the schedule
    const int64_t i = static_cast<int64_t>(blockIdx.x) * kLossThreads + threadIdx.x;
    double acc[kLossSums] = {0, 0, 0, 0, 0};
    if (i < N) {
        const float m = (valid != nullptr) ? valid[i] : 1.0f;
        // With no mask the mean's 1/N is known now; with a mask the finalize kernel divides.
        const float w = (valid != nullptr) ? m : 1.0f / static_cast<float>(N);
        // An action outside [0, A) must not become an out-of-bounds access: index 0 is used for the addresses and
        // the selected probability is poisoned with NaN, so the loss and this row's gradient are NaN and the caller
        // sees it (the reference raises an IndexError in select_at_indexes, rlpyt/utils/tensor.py:5-15).
        const int64_t a_raw = action[i];
        const bool a_bad = a_raw < 0 || a_raw >= A;
        const int a = a_bad ? 0 : static_cast<int>(a_raw);
        const float* p = prob_new + i * A;
        float ent = 0.0f;
        float pa = 0.0f;
        // entropy term + its gradient, one pass over the A probabilities
        for (int k = 0; k < A; ++k) {
            const float pk = p[k];
            const float lg = logf(pk + kEps);
            ent -= pk * lg;                                             // categorical.py:34
            if (k == a) pa = pk;
            if (grad_prob != nullptr)                                    // d(-c_ent*ent)/dp_k
                grad_prob[i * A + k] = (c_ent * w) * (lg + pk / (pk + kEps));
        }
        if (a_bad) pa = __int_as_float(0x7fc00000);
        const float Ai = adv[i];
        float pi_term, g_pa;  // pi_term enters the mean with sign -1; g_pa = d(pi_loss_i)/dp_a / w
        if (PPO) {
            const float den = prob_old[i * A + a] + kEps;
            const float ratio = (pa + kEps) / den;                      // categorical.py:43
            const float s1 = ratio * Ai;                                // ppo.py:138
            const float lo = 1.0f - clip, hi = 1.0f + clip;
            const float cr = fminf(fmaxf(ratio, lo), hi);               // ppo.py:139
            const float s2 = cr * Ai;                                   // ppo.py:141
            pi_term = fminf(s1, s2);                                    // ppo.py:142
            const float inrange = (ratio >= lo && ratio <= hi) ? 1.0f : 0.0f;
            const float dsel = (s1 < s2) ? 1.0f : ((s1 > s2) ? inrange : 0.5f + 0.5f * inrange);
            g_pa = -(Ai * dsel) / den;
        } else {
            const float den = pa + kEps;
            pi_term = logf(den) * Ai;                                   // a2c.py:89-90
            g_pa = -Ai / den;
        }
        const float dv = value[i] - ret[i];
        const float ve = 0.5f * dv * dv;                                // ppo.py:145 / a2c.py:92
        if (grad_prob != nullptr) grad_prob[i * A + a] += w * g_pa;
        if (grad_value != nullptr) grad_value[i] = (c_v * w) * dv;
        acc[0] = static_cast<double>(pi_term) * m;
        acc[1] = static_cast<double>(ve) * m;
        acc[2] = static_cast<double>(ent) * m;
        acc[3] = static_cast<double>(expf(ent)) * m;                    // base.py:66-68
        acc[4] = static_cast<double>(m);
    }
#pragma unroll
    for (int s = 0; s < kLossSums; ++s) acc[s] = warp_sum(acc[s]);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < kLossSums; ++s) sh[s][wid] = acc[s];
    }
    __syncthreads();
    if (threadIdx.x < kLossSums) {
        double t = 0;
        for (int k = 0; k < kLossThreads / 32; ++k) t += sh[threadIdx.x][k];
        partials[static_cast<int64_t>(blockIdx.x) * kLossSums + threadIdx.x] = t;
    }
}

__global__ void __launch_bounds__(kLossThreads)
pg_loss_finalize_kernel(const double* __restrict__ partials, int nparts, int64_t N, int A,
                        float c_v, float c_ent, int masked, float* __restrict__ scalars,
                        float* __restrict__ grad_prob, float* __restrict__ grad_value) {
    __shared__ double tot[kLossSums];
    if (threadIdx.x < kLossSums) {
        double t = 0;
        for (int k = 0; k < nparts; ++k) t += partials[static_cast<int64_t>(k) * kLossSums + threadIdx.x];
        tot[threadIdx.x] = t;
    }
    __syncthreads();
    const double cnt = tot[4];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const double pi_loss = -tot[0] / cnt;
        const double value_loss = static_cast<double>(c_v) * (tot[1] / cnt);
        const double entropy = tot[2] / cnt;
        const double perplexity = tot[3] / cnt;
        scalars[0] = static_cast<float>(pi_loss + value_loss - static_cast<double>(c_ent) * entropy);
        scalars[1] = static_cast<float>(entropy);
        scalars[2] = static_cast<float>(perplexity);
        scalars[3] = static_cast<float>(pi_loss);
        scalars[4] = static_cast<float>(value_loss);
        scalars[5] = static_cast<float>(cnt);
        scalars[6] = 0.0f;
        scalars[7] = 0.0f;
    }
    if (masked) {
        const float inv = static_cast<float>(1.0 / cnt);
        const int64_t stride = static_cast<int64_t>(gridDim.x) * kLossThreads;
        const int64_t start = static_cast<int64_t>(blockIdx.x) * kLossThreads + threadIdx.x;
        if (grad_prob != nullptr)
            for (int64_t j = start; j < N * A; j += stride) grad_prob[j] *= inv;
        if (grad_value != nullptr)
            for (int64_t j = start; j < N; j += stride) grad_value[j] *= inv;
    }
}

template <bool PPO>
static int launch_loss(const float* prob_new, const float* value, const float* prob_old,
                       const int64_t* action, const float* ret, const float* adv, const float* valid,
                       int64_t N, int A, float clip, float c_v, float c_ent, float* scalars,
                       float* grad_prob, float* grad_value, void* scratch, cudaStream_t st,
                       const float* clip_dev = nullptr) {
    const int nb = loss_blocks(N);
    double* partials = static_cast<double*>(scratch);
    pg_loss_kernel<PPO><<<nb, kLossThreads, 0, st>>>(prob_new, value, prob_old, action, ret, adv, valid,
                                                     N, A, clip, c_v, c_ent, grad_prob, grad_value, partials, clip_dev);
    int rc = check_launch("pg_loss_kernel");
    if (rc != RL_OK) return rc;
    const int masked = valid != nullptr;
    const int fb = masked ? nb : 1;
    pg_loss_finalize_kernel<<<fb, kLossThreads, 0, st>>>(partials, nb, N, A, c_v, c_ent, masked, scalars,
                                                         grad_prob, grad_value);
    return check_launch("pg_loss_finalize_kernel");
}

}  // namespace rl

extern "C" {

int64_t rl_pg_loss_scratch_bytes(int64_t N) {
    if (N < 1) N = 1;
    return static_cast<int64_t>(rl::loss_blocks(N)) * rl::kLossSums * sizeof(double);
}

int rl_ppo_loss_f32(const float* prob_new, const float* value, const float* prob_old,
                    const int64_t* action, const float* return_, const float* advantage,
                    const float* valid, int64_t N, int A, float ratio_clip, float value_loss_coeff,
                    float entropy_loss_coeff, float* out_scalars, float* grad_prob, float* grad_value,
                    void* scratch, void* stream) {
    RL_REQUIRE(prob_new && value && prob_old && action && return_ && advantage && out_scalars && scratch,
               RL_EINVAL, "rl_ppo_loss_f32: null pointer");
    RL_REQUIRE(N >= 1 && A >= 1 && A <= rl::kMaxA, RL_EINVAL, "rl_ppo_loss_f32: N=%lld A=%d", (long long)N, A);
    RL_REQUIRE(rl::aligned(scratch, 8), RL_EALIGN, "rl_ppo_loss_f32: scratch must be 8B aligned");
    return rl::launch_loss<true>(prob_new, value, prob_old, action, return_, advantage, valid, N, A,
                                 ratio_clip, value_loss_coeff, entropy_loss_coeff, out_scalars,
                                 grad_prob, grad_value, scratch, rl::as_stream(stream));
}

int rl_ppo_loss_devclip_f32(const float* prob_new, const float* value, const float* prob_old,
                            const int64_t* action, const float* return_, const float* advantage,
                            const float* valid, int64_t N, int A, const float* ratio_clip_dev, float value_loss_coeff,
                            float entropy_loss_coeff, float* out_scalars, float* grad_prob, float* grad_value,
                            void* scratch, void* stream) {
    RL_REQUIRE(prob_new && value && prob_old && action && return_ && advantage && out_scalars && scratch && ratio_clip_dev,
               RL_EINVAL, "rl_ppo_loss_devclip_f32: null pointer");
    RL_REQUIRE(N >= 1 && A >= 1 && A <= rl::kMaxA, RL_EINVAL, "rl_ppo_loss_devclip_f32: N=%lld A=%d", (long long)N, A);
    RL_REQUIRE(rl::aligned(scratch, 8), RL_EALIGN, "rl_ppo_loss_devclip_f32: scratch must be 8B aligned");
    return rl::launch_loss<true>(prob_new, value, prob_old, action, return_, advantage, valid, N, A, 0.0f,
                                 value_loss_coeff, entropy_loss_coeff, out_scalars, grad_prob, grad_value, scratch,
                                 rl::as_stream(stream), ratio_clip_dev);
}

int rl_a2c_loss_f32(const float* prob, const float* value, const int64_t* action, const float* return_,
                    const float* advantage, const float* valid, int64_t N, int A,
                    float value_loss_coeff, float entropy_loss_coeff, float* out_scalars,
                    float* grad_prob, float* grad_value, void* scratch, void* stream) {
    RL_REQUIRE(prob && value && action && return_ && advantage && out_scalars && scratch, RL_EINVAL,
               "rl_a2c_loss_f32: null pointer");
    RL_REQUIRE(N >= 1 && A >= 1 && A <= rl::kMaxA, RL_EINVAL, "rl_a2c_loss_f32: N=%lld A=%d", (long long)N, A);
    RL_REQUIRE(rl::aligned(scratch, 8), RL_EALIGN, "rl_a2c_loss_f32: scratch must be 8B aligned");
    return rl::launch_loss<false>(prob, value, nullptr, action, return_, advantage, valid, N, A, 0.0f,
                                  value_loss_coeff, entropy_loss_coeff, out_scalars, grad_prob,
                                  grad_value, scratch, rl::as_stream(stream));
}

}  // extern "C"
