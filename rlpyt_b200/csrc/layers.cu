// Small HBM-bound helpers around the tensor-core layer kernels.
//
//   rl_relu_backward_f32 : dst = grad * (out > 0)     the ReLU backward of rlpyt/models/conv2d.py:41 and
//                          rlpyt/models/mlp.py:33 (torch.nn.ReLU after every conv / linear layer) in one
//                          pass (autograd's eager form is a compare kernel writing a bool tensor + a multiply)
//   rl_transpose_f32     : dst[c][r] = src[r][c]       operand re-layout for the "TN" GEMM of gemm_tf32x3.cu
//                          (weight gradient: both operands need the batch axis contiguous), 32x32 tiles
//                          through padded shared memory, both sides coalesced
// Both are pure data movement: bit-exact against torch.
#include "common.cuh"

namespace rl {

__global__ void __launch_bounds__(256)
relu_backward_kernel(const float* __restrict__ grad, const float* __restrict__ out, float* __restrict__ dst,
                     int64_t n) {
    const int64_t nv = n / 4;
    const float4* g4 = reinterpret_cast<const float4*>(grad);
    const float4* o4 = reinterpret_cast<const float4*>(out);
    float4* d4 = reinterpret_cast<float4*>(dst);
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nv; i += stride) {
        const float4 g = ldg_stream(g4 + i), o = ldg_stream(o4 + i);
        float4 d;
        d.x = o.x > 0.0f ? g.x : 0.0f;
        d.y = o.y > 0.0f ? g.y : 0.0f;
        d.z = o.z > 0.0f ? g.z : 0.0f;
        d.w = o.w > 0.0f ? g.w : 0.0f;
        d4[i] = d;
    }
    for (int64_t i = nv * 4 + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
        dst[i] = out[i] > 0.0f ? grad[i] : 0.0f;
}

__global__ void __launch_bounds__(256)
relu_backward_scalar_kernel(const float* __restrict__ grad, const float* __restrict__ out, float* __restrict__ dst,
                            int64_t n) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
        dst[i] = out[i] > 0.0f ? grad[i] : 0.0f;
}

// 32 x 32 tile per (32 x 8)-thread block; tile index over grid.x so the row count is not limited to 65535 blocks
__global__ void __launch_bounds__(256)
transpose_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t rows, int64_t cols, int64_t tiles_c) {
    __shared__ float tile[32][33];
    const int64_t t = blockIdx.x;
    const int64_t r0 = (t / tiles_c) * 32, c0 = (t % tiles_c) * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
        const int64_t r = r0 + ty + j, c = c0 + tx;
        if (r < rows && c < cols) tile[ty + j][tx] = src[r * cols + c];
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
        const int64_t c = c0 + ty + j, r = r0 + tx;
        if (r < rows && c < cols) dst[c * rows + r] = tile[tx][ty + j];
    }
}

}  // namespace rl

extern "C" {

int rl_relu_backward_f32(const float* grad, const float* out, float* dst, int64_t n, void* stream) {
    if (n == 0) return RL_OK;
    RL_REQUIRE(grad && out && dst && n > 0, RL_EINVAL, "rl_relu_backward_f32: null pointer or negative n");
    int sms = rl::sm_count();
    if (sms <= 0) sms = 148;
    int64_t blocks = (n / 4 + 255) / 256;
    if (blocks < 1) blocks = 1;
    if (blocks > 8LL * sms) blocks = 8LL * sms;
    if (rl::aligned(grad, 16) && rl::aligned(out, 16) && rl::aligned(dst, 16))
        rl::relu_backward_kernel<<<static_cast<unsigned>(blocks), 256, 0, rl::as_stream(stream)>>>(grad, out, dst, n);
    else
        rl::relu_backward_scalar_kernel<<<static_cast<unsigned>(blocks), 256, 0, rl::as_stream(stream)>>>(grad, out,
                                                                                                          dst, n);
    return rl::check_launch("relu_backward_kernel");
}

int rl_transpose_f32(const float* src, float* dst, int64_t rows, int64_t cols, void* stream) {
    if (rows == 0 || cols == 0) return RL_OK;
    RL_REQUIRE(src && dst && rows > 0 && cols > 0, RL_EINVAL, "rl_transpose_f32: null pointer or negative extent");
    const int64_t tiles_r = (rows + 31) / 32, tiles_c = (cols + 31) / 32;
    RL_REQUIRE(tiles_r * tiles_c < (int64_t(1) << 31), RL_EINVAL, "rl_transpose_f32: matrix too large");
    rl::transpose_kernel<<<static_cast<unsigned>(tiles_r * tiles_c), 256, 0, rl::as_stream(stream)>>>(src, dst, rows,
                                                                                                       cols, tiles_c);
    return rl::check_launch("transpose_kernel");
}

}  // extern "C"
