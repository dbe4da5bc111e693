// First layer of the AtariFf network, hand-written for its real input: uint8 frames.
//
// Reference: rlpyt/models/pg/atari_ff_model.py:50-53 (img.float().mul_(1/255) then the conv stack)
// with rlpyt/models/conv2d.py:36-44 (Conv2d(C=4 -> 16, k=8, s=4, p=0) + ReLU).  It runs in BOTH
// halves of the path: once per env step in agent.step (N = B = 256) and 16x per PPO iteration on
// the minibatches (N = 8192), where cuDNN's fp32 kernels for this 4-channel 8x8/stride-4 layer
// take 2.1 ms (forward, autotuned; 6.1 ms with the default heuristic) after torch has first
// materialised the gathered minibatch as uint8 (231 MB), converted it to fp32 (925 MB) and scaled it.
//
// conv1_fwd_kernel fuses: [optional row gather from the resident [T*B] batch] -> u8 -> fp32 * (1/255)
// (one rounding, as the reference) -> 8x8/s4 convolution -> + bias -> ReLU.  Reads 28 KB of uint8 per
// image instead of 113 KB of fp32; never writes the gathered or converted image.
//   * persistent CTAs (grid = multiple of the SM count), one image per iteration staged in shared
//     memory as uint8; weights transposed once per CTA to [k][oc] so the 16 output channels of one
//     filter tap are four broadcast LDS.128;
//   * each thread owns 4 horizontally adjacent output positions x 16 channels = 64 fp32
//     accumulators; per (c, ky) it reads 20 input bytes and 8x16 weights for 512 FMAs.
// conv1_wgrad_kernel: dW[oc][k] = sum_n sum_p g[n][oc][p] * x[n][patch(p)][k], g = dY * (Y > 0):
//   thread t owns the 4 filter taps {t, t+64, t+128, t+192} (the 4 input frames at one (ky,kx)) for
//   all 16 channels = 64 accumulators, so each broadcast load of a position's 16 gradients feeds 64
//   FMAs (the first version, 1 tap per thread, was shared-memory-issue bound: FMA pipe 36 % busy,
//   profiles/r01_ppo_kernels_full_summary.json); image and masked gradient are staged in shared
//   memory ([p][oc]); per-CTA partial sums are reduced by conv1_wgrad_reduce_kernel in CTA order
//   => deterministic (no float atomics).
// fp32 SIMT on purpose: parity with the reference is 1e-5 and the inputs are exact; a tensor-core
// version needs a 3-term bf16 split of the weights (pixels are exact in bf16) - noted in DESIGN.md.
#include "common.cuh"

namespace rl {

constexpr int kC1In = 4;      // input channels (frames)
constexpr int kC1K = 8;       // kernel size
constexpr int kC1S = 4;       // stride
constexpr int kC1Out = 16;    // output channels
constexpr int kC1Taps = kC1In * kC1K * kC1K;  // 256
constexpr int kFwdThreads = 128;
constexpr int kWgradThreads = 64;       // each thread owns 4 filter taps x 16 channels
constexpr float kInv255 = 1.0f / 255.0f;      // python 1./255 rounded to fp32 (atari_ff_model.py:51)

__device__ __forceinline__ void stage_image(uint8_t* s_img, const uint8_t* __restrict__ src, int bytes,
                                            int tid, int nthreads) {
    // bytes is a multiple of 16 (checked on the host); 16 B coalesced loads
    const uint4* s4 = reinterpret_cast<const uint4*>(src);
    uint4* d4 = reinterpret_cast<uint4*>(s_img);
    for (int i = tid; i < bytes / 16; i += nthreads) d4[i] = ldg_stream(s4 + i);
}

__global__ void __launch_bounds__(kFwdThreads)
conv1_fwd_kernel(const uint8_t* __restrict__ obs, const int64_t* __restrict__ rows,
                 const float* __restrict__ weight, const float* __restrict__ bias, float* __restrict__ Y,
                 int N, int H, int W, int OH, int OW, int relu) {
    extern __shared__ __align__(16) uint8_t smem[];
    float* s_w = reinterpret_cast<float*>(smem);                 // [256][16]
    uint8_t* s_img = smem + kC1Taps * kC1Out * sizeof(float);    // [4][H][W] (+32 B slack)
    const int img_bytes = kC1In * H * W;
    for (int i = threadIdx.x; i < kC1Taps * kC1Out; i += kFwdThreads) {
        const int oc = i / kC1Taps, k = i % kC1Taps;              // global layout [oc][c][ky][kx]
        s_w[k * kC1Out + oc] = weight[i];
    }
    float b[kC1Out];
#pragma unroll
    for (int oc = 0; oc < kC1Out; ++oc) b[oc] = bias[oc];
    const int groups = (OW + 3) / 4;
    const int tiles = OH * groups;

    for (int n = blockIdx.x; n < N; n += gridDim.x) {
        const int64_t r = rows != nullptr ? rows[n] : n;
        __syncthreads();  // previous image fully consumed (also orders the weight stores first time)
        stage_image(s_img, obs + r * img_bytes, img_bytes, threadIdx.x, kFwdThreads);
        __syncthreads();
        for (int tile = threadIdx.x; tile < tiles; tile += kFwdThreads) {
            const int oy = tile / groups, ox0 = (tile % groups) * 4;
            float acc[4][kC1Out];
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int oc = 0; oc < kC1Out; ++oc) acc[p][oc] = b[oc];
#pragma unroll 1
            for (int c = 0; c < kC1In; ++c) {
#pragma unroll 1
                for (int ky = 0; ky < kC1K; ++ky) {
                    // 20 input bytes: x = 4*ox0 .. 4*ox0+19 of row 4*oy+ky (4 B aligned: W % 4 == 0)
                    const uint32_t* rowp = reinterpret_cast<const uint32_t*>(
                        s_img + (c * H + oy * kC1S + ky) * W + ox0 * kC1S);
                    float px[20];
#pragma unroll
                    for (int w = 0; w < 5; ++w) {
                        const uint32_t v = rowp[w];
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            px[4 * w + j] = __fmul_rn(static_cast<float>((v >> (8 * j)) & 0xffu), kInv255);
                    }
                    const float4* wp = reinterpret_cast<const float4*>(s_w + (c * 64 + ky * 8) * kC1Out);
#pragma unroll
                    for (int kx = 0; kx < kC1K; ++kx) {
                        float wv[kC1Out];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float4 t = wp[kx * 4 + q];
                            wv[4 * q] = t.x; wv[4 * q + 1] = t.y; wv[4 * q + 2] = t.z; wv[4 * q + 3] = t.w;
                        }
#pragma unroll
                        for (int p = 0; p < 4; ++p) {
                            const float x = px[4 * p + kx];
#pragma unroll
                            for (int oc = 0; oc < kC1Out; ++oc) acc[p][oc] = fmaf(x, wv[oc], acc[p][oc]);
                        }
                    }
                }
            }
            float* yb = Y + (static_cast<int64_t>(n) * kC1Out * OH + oy) * OW + ox0;
            const bool full = (OW % 4 == 0);
#pragma unroll
            for (int oc = 0; oc < kC1Out; ++oc) {
                float o[4];
#pragma unroll
                for (int p = 0; p < 4; ++p) o[p] = relu ? fmaxf(acc[p][oc], 0.0f) : acc[p][oc];
                float* yp = yb + static_cast<int64_t>(oc) * OH * OW;
                if (full) {
                    *reinterpret_cast<float4*>(yp) = make_float4(o[0], o[1], o[2], o[3]);
                } else {
#pragma unroll
                    for (int p = 0; p < 4; ++p)
                        if (ox0 + p < OW) yp[p] = o[p];
                }
            }
        }
    }
}

__global__ void __launch_bounds__(kWgradThreads)
conv1_wgrad_kernel(const uint8_t* __restrict__ obs, const int64_t* __restrict__ rows,
                   const float* __restrict__ Y, const float* __restrict__ dY, float* __restrict__ partial,
                   int N, int H, int W, int OH, int OW, int relu) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int P = OH * OW;
    float* s_g = reinterpret_cast<float*>(smem);                       // [P][16]
    uint8_t* s_img = smem + static_cast<size_t>(P) * kC1Out * sizeof(float);
    const int img_bytes = kC1In * H * W;
    const int t = threadIdx.x;                                          // taps t + 64*c, c = 0..3
    const int ky = (t >> 3) & 7, kx = t & 7;
    float acc[kC1In][kC1Out];
#pragma unroll
    for (int c = 0; c < kC1In; ++c)
#pragma unroll
        for (int oc = 0; oc < kC1Out; ++oc) acc[c][oc] = 0.0f;
    float accb = 0.0f;                                                  // threads 0..15: bias gradient

    for (int n = blockIdx.x; n < N; n += gridDim.x) {
        const int64_t r = rows != nullptr ? rows[n] : n;
        __syncthreads();
        stage_image(s_img, obs + r * img_bytes, img_bytes, threadIdx.x, kWgradThreads);
        const float* yn = Y + static_cast<int64_t>(n) * kC1Out * P;
        const float* gn = dY + static_cast<int64_t>(n) * kC1Out * P;
        // [oc][p] -> [p][oc]: consecutive threads take consecutive oc of one position so the
        // transposed store is conflict free (the strided global read is served from L2 sectors)
        for (int i = threadIdx.x; i < kC1Out * P; i += kWgradThreads) {
            const int p = i >> 4, oc = i & 15;
            const int gi = oc * P + p;
            const float g = gn[gi];
            s_g[i] = (!relu || yn[gi] > 0.0f) ? g : 0.0f;               // ReLU backward (threshold)
        }
        __syncthreads();
        const uint8_t* base = s_img + ky * W + kx;
        const int plane = H * W;
        for (int oy = 0; oy < OH; ++oy) {
            const uint8_t* rowp = base + oy * kC1S * W;
            const float4* gp = reinterpret_cast<const float4*>(s_g + static_cast<size_t>(oy) * OW * kC1Out);
#pragma unroll 2
            for (int ox = 0; ox < OW; ++ox) {
                float x[kC1In];
#pragma unroll
                for (int c = 0; c < kC1In; ++c)
                    x[c] = __fmul_rn(static_cast<float>(rowp[c * plane + ox * kC1S]), kInv255);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 g = gp[ox * 4 + q];
#pragma unroll
                    for (int c = 0; c < kC1In; ++c) {
                        acc[c][4 * q] = fmaf(x[c], g.x, acc[c][4 * q]);
                        acc[c][4 * q + 1] = fmaf(x[c], g.y, acc[c][4 * q + 1]);
                        acc[c][4 * q + 2] = fmaf(x[c], g.z, acc[c][4 * q + 2]);
                        acc[c][4 * q + 3] = fmaf(x[c], g.w, acc[c][4 * q + 3]);
                    }
                }
            }
        }
        if (threadIdx.x < kC1Out) {
            float s = 0.0f;
            for (int p = 0; p < P; ++p) s += s_g[p * kC1Out + threadIdx.x];
            accb += s;
        }
    }
    float* out = partial + static_cast<int64_t>(blockIdx.x) * (kC1Out * kC1Taps + kC1Out);
#pragma unroll
    for (int c = 0; c < kC1In; ++c)
#pragma unroll
        for (int oc = 0; oc < kC1Out; ++oc) out[oc * kC1Taps + c * 64 + t] = acc[c][oc];   // [oc][c][ky][kx]
    if (threadIdx.x < kC1Out) out[kC1Out * kC1Taps + threadIdx.x] = accb;
}

__global__ void conv1_wgrad_reduce_kernel(const float* __restrict__ partial, int nparts,
                                          float* __restrict__ dW, float* __restrict__ db) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    constexpr int kAll = kC1Out * kC1Taps + kC1Out;
    if (i >= kAll) return;
    float s = 0.0f;
    for (int b = 0; b < nparts; ++b) s += partial[static_cast<int64_t>(b) * kAll + i];  // fixed order
    if (i < kC1Out * kC1Taps) dW[i] = s;
    else db[i - kC1Out * kC1Taps] = s;
}

static int conv1_check(int N, int C, int H, int W, int* OH, int* OW) {
    RL_REQUIRE(N >= 0 && C == kC1In && H >= kC1K && W >= kC1K, RL_EINVAL,
               "conv1: supports C=4 frames, 8x8 stride-4 kernel (got C=%d H=%d W=%d)", C, H, W);
    RL_REQUIRE(W % 4 == 0 && (C * H * W) % 16 == 0, RL_EALIGN, "conv1: W %% 4 == 0 and C*H*W %% 16 == 0 required");
    *OH = (H - kC1K) / kC1S + 1;
    *OW = (W - kC1K) / kC1S + 1;
    return RL_OK;
}

}  // namespace rl

extern "C" {

int rl_conv1_u8_forward(const uint8_t* obs, const int64_t* rows, const float* weight, const float* bias,
                        float* out, int64_t N, int C, int H, int W, int relu, void* stream) {
    RL_REQUIRE(obs && weight && bias && out, RL_EINVAL, "rl_conv1_u8_forward: null pointer");
    int OH, OW;
    int rc = rl::conv1_check(static_cast<int>(N), C, H, W, &OH, &OW);
    if (rc != RL_OK) return rc;
    if (N == 0) return RL_OK;
    RL_REQUIRE(rl::aligned(obs, 16) && rl::aligned(out, 16), RL_EALIGN, "rl_conv1_u8_forward: 16B alignment");
    const size_t smem = rl::kC1Taps * rl::kC1Out * sizeof(float) + static_cast<size_t>(C) * H * W + 32;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(rl::conv1_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        attr_set = true;
    }
    RL_REQUIRE(smem <= 160 * 1024, RL_EINVAL, "rl_conv1_u8_forward: image too large for shared memory");
    int sms = rl::sm_count();
    if (sms <= 0) sms = 148;
    const int per_sm = static_cast<int>((200 * 1024) / smem) > 4 ? 4 : static_cast<int>((200 * 1024) / smem);
    int64_t grid = static_cast<int64_t>(sms) * (per_sm < 1 ? 1 : per_sm);
    if (grid > N) grid = N;
    rl::conv1_fwd_kernel<<<static_cast<unsigned>(grid), rl::kFwdThreads, smem, rl::as_stream(stream)>>>(
        obs, rows, weight, bias, out, static_cast<int>(N), H, W, OH, OW, relu);
    return rl::check_launch("conv1_fwd_kernel");
}

int64_t rl_conv1_u8_wgrad_scratch_bytes(void) {
    int sms = rl::sm_count();
    if (sms <= 0) sms = 148;
    return static_cast<int64_t>(sms) * 4 * (rl::kC1Out * rl::kC1Taps + rl::kC1Out) * sizeof(float);
}

int rl_conv1_u8_wgrad(const uint8_t* obs, const int64_t* rows, const float* out, const float* grad_out,
                      float* grad_weight, float* grad_bias, int64_t N, int C, int H, int W, int relu,
                      void* scratch, void* stream) {
    RL_REQUIRE(obs && out && grad_out && grad_weight && grad_bias && scratch, RL_EINVAL,
               "rl_conv1_u8_wgrad: null pointer");
    int OH, OW;
    int rc = rl::conv1_check(static_cast<int>(N), C, H, W, &OH, &OW);
    if (rc != RL_OK) return rc;
    RL_REQUIRE(N >= 1, RL_EINVAL, "rl_conv1_u8_wgrad: N=%lld", (long long)N);
    RL_REQUIRE(rl::aligned(obs, 16), RL_EALIGN, "rl_conv1_u8_wgrad: obs must be 16B aligned");
    const size_t smem = static_cast<size_t>(OH) * OW * rl::kC1Out * sizeof(float) + static_cast<size_t>(C) * H * W + 32;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(rl::conv1_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr_set = true;
    }
    RL_REQUIRE(smem <= 200 * 1024, RL_EINVAL, "rl_conv1_u8_wgrad: image too large for shared memory");
    int sms = rl::sm_count();
    if (sms <= 0) sms = 148;
    int64_t grid = static_cast<int64_t>(sms) * 4;  // 4 CTAs of 2 warps per SM (54 KB smem each)
    if (grid > N) grid = N;
    float* partial = static_cast<float*>(scratch);
    cudaStream_t st = rl::as_stream(stream);
    rl::conv1_wgrad_kernel<<<static_cast<unsigned>(grid), rl::kWgradThreads, smem, st>>>(
        obs, rows, out, grad_out, partial, static_cast<int>(N), H, W, OH, OW, relu);
    rc = rl::check_launch("conv1_wgrad_kernel");
    if (rc != RL_OK) return rc;
    constexpr int kAll = rl::kC1Out * rl::kC1Taps + rl::kC1Out;
    rl::conv1_wgrad_reduce_kernel<<<(kAll + 255) / 256, 256, 0, st>>>(partial, static_cast<int>(grid), grad_weight,
                                                                      grad_bias);
    return rl::check_launch("conv1_wgrad_reduce_kernel");
}

}  // extern "C"
