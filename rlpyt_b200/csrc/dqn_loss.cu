// Fused DQN / Double-DQN loss: target construction, Huber (or MSE) loss, importance weights,
// TD-error priorities AND the gradient w.r.t. the online Q-values in one pass.
//
// Reference arithmetic (restated in oracle/dqn_loss.py): rlpyt/algos/dqn/dqn.py:230-263
//   q        = qs[i, action[i]]                                          select_at_indexes
//   target_q = target_qs[i, argmax_k next_qs[i,k]]   (double_dqn, :236-239; first maximal index)
//            = max_k target_qs[i,k]                   (else, :241)
//   y        = return_ + (1 - done_n) * (discount**n_step * target_q)    (:242-243)
//   delta    = y - q ; losses = 0.5 * delta**2                           (:244-245)
//   Huber    : losses = where(|delta| <= c, losses, c * (|delta| - c/2)) (:247-249)
//   losses  *= is_weights   (prioritized, :250-251)
//   td_abs_errors = clamp(|delta|, 0, c)                                 (:252-254)
//   loss     = mean(losses)                                              (:263)
// and what autograd derives: dloss/dq_i = -(w_i / N) * (|delta| <= c ? delta : c * sign(delta)).
//
// Every per-sample value is computed with explicitly rounded fp32 operations in the reference's
// order (no FMA contraction), so td_abs_errors - which feed the fp64 sum-tree through
// update_batch_priorities - are bit-identical to the reference's; the mean is accumulated in fp64
// with a fixed reduction order (deterministic), so the loss scalar agrees to fp32 rounding.
//
// The reference runs ~20 small torch-CPU ops after copying both networks' outputs to the host
// (rlpyt/agents/dqn/dqn_agent.py:28,75); here: 2 launches, nothing leaves the device, and the
// priorities go straight into rl_pow_f32_to_f64 / rl_sumtree_update_f64.
#include "common.cuh"

namespace rl {

constexpr int kDqnThreads = 256;
constexpr int kDqnMaxA = 64;

static inline int dqn_blocks(int64_t N) { return static_cast<int>((N + kDqnThreads - 1) / kDqnThreads); }

__global__ void __launch_bounds__(kDqnThreads)
dqn_loss_kernel(const float* __restrict__ qs, const float* __restrict__ target_qs,
                const float* __restrict__ next_qs, const int64_t* __restrict__ action,
                const float* __restrict__ ret, const uint8_t* __restrict__ done_n,
                const float* __restrict__ is_weights, int64_t N, int A, float disc_n, float delta_clip,
                float* __restrict__ td_abs, float* __restrict__ grad_qs, double* __restrict__ partials) {
    __shared__ double sh[kDqnThreads / 32];
    const int64_t i = static_cast<int64_t>(blockIdx.x) * kDqnThreads + threadIdx.x;
    double acc = 0.0;
    if (i < N) {
        // An action outside [0, A) (replay rows of the documented "wrong wrap", a model / env action-space
        // mismatch) must not become an out-of-bounds access: index 0 is read instead and the sample's Q-value is
        // poisoned with NaN, so the loss, the TD errors and the gradient of that row are NaN and the caller sees it
        // (the reference raises an IndexError in select_at_indexes).
        const int64_t a_raw = action[i];
        const bool a_bad = a_raw < 0 || a_raw >= A;
        const int a = a_bad ? 0 : static_cast<int>(a_raw);
        const float* tq = target_qs + i * A;
        float target_q;
        if (next_qs != nullptr) {                       // double DQN: online argmax, target value
            const float* nq = next_qs + i * A;
            int best = 0;
            float bv = nq[0];
            for (int k = 1; k < A; ++k) {
                const float v = nq[k];
                if (v > bv) { bv = v; best = k; }       // strict >: first maximal index (torch.argmax)
            }
            target_q = tq[best];
        } else {
            target_q = tq[0];
            for (int k = 1; k < A; ++k) target_q = fmaxf(target_q, tq[k]);
        }
        const float q = a_bad ? __int_as_float(0x7fc00000) : qs[i * A + a];
        const float disc_target_q = __fmul_rn(disc_n, target_q);
        const float not_done = done_n[i] ? 0.0f : 1.0f;
        const float y = __fadd_rn(ret[i], __fmul_rn(not_done, disc_target_q));
        const float delta = __fsub_rn(y, q);
        const float abs_delta = fabsf(delta);
        float loss = __fmul_rn(0.5f, __fmul_rn(delta, delta));
        float dl = delta;                               // d(loss_i)/d(delta)
        float td = abs_delta;
        if (delta_clip >= 0.0f) {
            if (!(abs_delta <= delta_clip)) {           // linear branch of the Huber loss
                loss = __fmul_rn(delta_clip, __fsub_rn(abs_delta, __fmul_rn(delta_clip, 0.5f)));
                dl = delta > 0.0f ? delta_clip : -delta_clip;
            }
            td = fminf(fmaxf(abs_delta, 0.0f), delta_clip);
        }
        float w = 1.0f;
        if (is_weights != nullptr) {
            w = is_weights[i];
            loss = __fmul_rn(loss, w);
        }
        td_abs[i] = td;
        if (grad_qs != nullptr) {
            const float g = -(w / static_cast<float>(N)) * dl;
            for (int k = 0; k < A; ++k) grad_qs[i * A + k] = (k == a) ? g : 0.0f;
        }
        acc = static_cast<double>(loss);
    }
    acc = warp_sum(acc);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) sh[wid] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < kDqnThreads / 32; ++k) t += sh[k];
        partials[blockIdx.x] = t;
    }
}

__global__ void dqn_loss_finalize_kernel(const double* __restrict__ partials, int nparts, int64_t N,
                                         float* __restrict__ scalars) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < nparts; ++k) t += partials[k];
        scalars[0] = static_cast<float>(t / static_cast<double>(N));
        scalars[1] = static_cast<float>(N);
    }
}

}  // namespace rl

extern "C" {

int64_t rl_dqn_loss_scratch_bytes(int64_t N) {
    if (N < 1) N = 1;
    return static_cast<int64_t>(rl::dqn_blocks(N)) * static_cast<int64_t>(sizeof(double));
}

int rl_dqn_loss_f32(const float* qs, const float* target_qs, const float* next_qs, const int64_t* action,
                    const float* return_, const uint8_t* done_n, const float* is_weights, int64_t N, int A,
                    float disc_n, float delta_clip, float* out_scalars, float* td_abs_errors, float* grad_qs,
                    void* scratch, void* stream) {
    RL_REQUIRE(qs && target_qs && action && return_ && done_n && out_scalars && td_abs_errors && scratch, RL_EINVAL,
               "rl_dqn_loss_f32: null pointer");
    RL_REQUIRE(N >= 1 && A >= 1 && A <= rl::kDqnMaxA, RL_EINVAL, "rl_dqn_loss_f32: N=%lld A=%d", (long long)N, A);
    RL_REQUIRE(rl::aligned(scratch, 8), RL_EALIGN, "rl_dqn_loss_f32: scratch must be 8B aligned");
    const int nb = rl::dqn_blocks(N);
    cudaStream_t st = rl::as_stream(stream);
    double* partials = static_cast<double*>(scratch);
    rl::dqn_loss_kernel<<<nb, rl::kDqnThreads, 0, st>>>(qs, target_qs, next_qs, action, return_, done_n, is_weights,
                                                        N, A, disc_n, delta_clip, td_abs_errors, grad_qs, partials);
    int rc = rl::check_launch("dqn_loss_kernel");
    if (rc != RL_OK) return rc;
    rl::dqn_loss_finalize_kernel<<<1, 32, 0, st>>>(partials, nb, N, out_scalars);
    return rl::check_launch("dqn_loss_finalize_kernel");
}

}  // extern "C"
