// Fused "global grad-norm -> clip -> Adam" over ONE flat fp32 parameter buffer.
//
// Replaces, per minibatch update (rlpyt/algos/pg/ppo.py:101-104, a2c.py:50-53):
//   torch.nn.utils.clip_grad_norm_(agent.parameters(), clip)  (per-tensor norms, stack, norm,
//   clamp, per-tensor mul_) and torch.optim.Adam.step() (per-tensor lerp/addcmul/sqrt/addcdiv)
// with two launches over the 1.65 M-element (6.6 MB) flat buffers.  The flat gradient buffer is
// also what the multi-GPU path all-reduces with a single NCCL call;
// the 1/world_size average is folded into grad_scale here, so no extra pass touches the
// gradients.
//
// Arithmetic follows torch 2.x:
//   total_norm = ||g||_2 ; coef = min(1, max_norm / (total_norm + 1e-6))       (clip_grad_norm_)
//   m = m + (g - m) * (1 - b1) ; v = v * b2 + (1 - b2) * g * g
//   p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)              (Adam, no amsgrad)
// Sum of squares is accumulated in fp64 with a fixed reduction order (deterministic).
#include "common.cuh"

namespace rl {

constexpr int kOptThreads = 256;
constexpr int kOptMaxBlocks = 592;  // 4 x 148 SMs

static inline int opt_blocks(int64_t n) {
    int64_t b = (n + kOptThreads * 4 - 1) / (kOptThreads * 4);
    if (b < 1) b = 1;
    if (b > kOptMaxBlocks) b = kOptMaxBlocks;
    return static_cast<int>(b);
}

__global__ void __launch_bounds__(kOptThreads)
grad_sqsum_kernel(const float* __restrict__ grad, int64_t n, float grad_scale, double* __restrict__ partials) {
    __shared__ double sh[kOptThreads / 32];
    double s = 0.0;
    const int64_t nv = n / 4;
    const float4* g4 = reinterpret_cast<const float4*>(grad);
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * kOptThreads + threadIdx.x; i < nv;
         i += static_cast<int64_t>(gridDim.x) * kOptThreads) {
        const float4 g = g4[i];
        const float a = g.x * grad_scale, b = g.y * grad_scale, c = g.z * grad_scale, d = g.w * grad_scale;
        s += static_cast<double>(a) * a + static_cast<double>(b) * b + static_cast<double>(c) * c +
             static_cast<double>(d) * d;
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const float a = grad[nv * 4 + threadIdx.x] * grad_scale;
        s += static_cast<double>(a) * a;
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int k = 0; k < kOptThreads / 32; ++k) t += sh[k];
        partials[blockIdx.x] = t;
    }
}

struct AdamHyper {
    float lr, beta1, beta2, eps, weight_decay, max_norm, grad_scale;
    float bias1, bias2_sqrt;  // 1 - b1^t, sqrt(1 - b2^t), computed on the host in double
};

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, const AdamHyper& h, float coef) {
    g = g * h.grad_scale * coef;
    if (h.weight_decay != 0.0f) g += h.weight_decay * p;
    m = m + (g - m) * (1.0f - h.beta1);
    v = v * h.beta2 + (1.0f - h.beta2) * g * g;
    const float denom = sqrtf(v) / h.bias2_sqrt + h.eps;
    p = p - (h.lr / h.bias1) * (m / denom);
}

__global__ void __launch_bounds__(kOptThreads)
clip_adam_kernel(float* __restrict__ param, const float* __restrict__ grad, float* __restrict__ exp_avg,
                 float* __restrict__ exp_avg_sq, int64_t n, AdamHyper h, const double* __restrict__ partials,
                 int nparts, float* __restrict__ norm_out) {
    __shared__ float sh_coef;
    if (threadIdx.x == 0) {
        double t = 0;
        for (int k = 0; k < nparts; ++k) t += partials[k];
        const float norm = static_cast<float>(sqrt(t));
        float coef = 1.0f;
        if (h.max_norm > 0.0f) coef = fminf(1.0f, h.max_norm / (norm + 1e-6f));
        sh_coef = coef;
        if (blockIdx.x == 0 && norm_out != nullptr) norm_out[0] = norm;
    }
    __syncthreads();
    const float coef = sh_coef;
    const int64_t nv = n / 4;
    float4* p4 = reinterpret_cast<float4*>(param);
    const float4* g4 = reinterpret_cast<const float4*>(grad);
    float4* m4 = reinterpret_cast<float4*>(exp_avg);
    float4* v4 = reinterpret_cast<float4*>(exp_avg_sq);
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * kOptThreads + threadIdx.x; i < nv;
         i += static_cast<int64_t>(gridDim.x) * kOptThreads) {
        float4 p = p4[i], m = m4[i], v = v4[i];
        const float4 g = g4[i];
        adam_one(p.x, g.x, m.x, v.x, h, coef);
        adam_one(p.y, g.y, m.y, v.y, h, coef);
        adam_one(p.z, g.z, m.z, v.z, h, coef);
        adam_one(p.w, g.w, m.w, v.w, h, coef);
        p4[i] = p; m4[i] = m; v4[i] = v;
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const int64_t j = nv * 4 + threadIdx.x;
        adam_one(param[j], grad[j], exp_avg[j], exp_avg_sq[j], h, coef);
    }
}

}  // namespace rl

extern "C" {

int64_t rl_clip_adam_scratch_bytes(int64_t n) {
    if (n < 1) n = 1;
    return static_cast<int64_t>(rl::opt_blocks(n)) * sizeof(double);
}

int rl_clip_adam_f32(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                     float lr, float beta1, float beta2, float eps, float weight_decay, int64_t step,
                     float max_norm, float grad_scale, float* norm_out, void* scratch, void* stream) {
    RL_REQUIRE(param && grad && exp_avg && exp_avg_sq && scratch, RL_EINVAL, "rl_clip_adam_f32: null pointer");
    RL_REQUIRE(n >= 1 && step >= 1, RL_EINVAL, "rl_clip_adam_f32: n=%lld step=%lld", (long long)n, (long long)step);
    RL_REQUIRE(rl::aligned(param, 16) && rl::aligned(grad, 16) && rl::aligned(exp_avg, 16) &&
                   rl::aligned(exp_avg_sq, 16) && rl::aligned(scratch, 8),
               RL_EALIGN, "rl_clip_adam_f32: buffers must be 16B aligned (scratch 8B)");
    rl::AdamHyper h;
    h.lr = lr; h.beta1 = beta1; h.beta2 = beta2; h.eps = eps; h.weight_decay = weight_decay;
    h.max_norm = max_norm; h.grad_scale = grad_scale;
    h.bias1 = static_cast<float>(1.0 - pow(static_cast<double>(beta1), static_cast<double>(step)));
    h.bias2_sqrt = static_cast<float>(sqrt(1.0 - pow(static_cast<double>(beta2), static_cast<double>(step))));
    const int nb = rl::opt_blocks(n);
    double* partials = static_cast<double*>(scratch);
    cudaStream_t st = rl::as_stream(stream);
    rl::grad_sqsum_kernel<<<nb, rl::kOptThreads, 0, st>>>(grad, n, grad_scale, partials);
    int rc = rl::check_launch("grad_sqsum_kernel");
    if (rc != RL_OK) return rc;
    rl::clip_adam_kernel<<<nb, rl::kOptThreads, 0, st>>>(param, grad, exp_avg, exp_avg_sq, n, h, partials, nb,
                                                         norm_out);
    return rl::check_launch("clip_adam_kernel");
}

}  // extern "C"
