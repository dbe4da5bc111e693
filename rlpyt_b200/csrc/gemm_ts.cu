// C-ABI of the TMEM-operand 3xTF32 GEMM (gemm_ts.cuh) and of the two helpers that prepare its B operand.
#include "gemm_ts.cuh"

extern "C" {

int64_t rl_gemm_ts_workspace_bytes(int64_t M, int64_t N, int64_t K) {
    if (M < 1 || N < 1 || K < 1) return 0;
    int sms = rl::sm_count();
    if (sms <= 0) sms = 148;
    return rl::gts::workspace_bytes(M, N, K, sms);
}

int rl_gemm_ts_f32(const float* A, int a_mmajor, const float* B, const float* B_lo, const float* bias, float* C, int c_trans,
                   int64_t M, int64_t N, int64_t K, int relu, void* workspace, void* stream) {
    RL_REQUIRE(A && B && B_lo && C, RL_EINVAL, "rl_gemm_ts_f32: null pointer");
    RL_REQUIRE(M >= 1 && N >= 1 && K >= 1 && M < (1LL << 31) && N < (1LL << 31) && K < (1LL << 31), RL_EINVAL,
               "rl_gemm_ts_f32: M=%lld N=%lld K=%lld", (long long)M, (long long)N, (long long)K);
    RL_REQUIRE(K % 4 == 0 && (!a_mmajor || M % 4 == 0), RL_EALIGN,
               "rl_gemm_ts_f32: row pitches must be multiples of 16 bytes (K %% 4 == 0; M %% 4 == 0 for an M-major A)");
    RL_REQUIRE(rl::aligned(A, 16) && rl::aligned(B, 16) && rl::aligned(B_lo, 16) && rl::aligned(C, 16) &&
                   (workspace == nullptr || rl::aligned(workspace, 16)),
               RL_EALIGN, "rl_gemm_ts_f32: 16-byte aligned A, B, B_lo, C, workspace required");
    int sms = rl::sm_count();
    if (sms <= 0) sms = 148;
    const cudaError_t e = rl::gts::launch(A, a_mmajor ? 1 : 0, B, B_lo, bias, C, c_trans ? 1 : 0, M, N, K, relu,
                                          static_cast<float*>(workspace), sms, rl::as_stream(stream));
    RL_REQUIRE(e != cudaErrorInvalidValue, RL_EINVAL, "rl_gemm_ts_f32: cuTensorMapEncodeTiled unavailable or failed");
    return rl::check_launch("gemm_ts_kernel");
}

int rl_gemm_ts_masked_f32(const float* A, const float* B, const float* B_lo, const float* out_mask, float* C, int64_t M,
                          int64_t N, int64_t K, void* workspace, void* stream) {
    RL_REQUIRE(A && B && B_lo && C && out_mask, RL_EINVAL, "rl_gemm_ts_masked_f32: null pointer");
    RL_REQUIRE(M >= 1 && N >= 1 && K >= 1 && M < (1LL << 31) && N < (1LL << 31) && K < (1LL << 31), RL_EINVAL,
               "rl_gemm_ts_masked_f32: M=%lld N=%lld K=%lld", (long long)M, (long long)N, (long long)K);
    RL_REQUIRE(K % 4 == 0, RL_EALIGN, "rl_gemm_ts_masked_f32: row pitches must be multiples of 16 bytes (K %% 4 == 0)");
    RL_REQUIRE(rl::aligned(A, 16) && rl::aligned(B, 16) && rl::aligned(B_lo, 16) && rl::aligned(C, 16) && rl::aligned(out_mask, 16) &&
                   (workspace == nullptr || rl::aligned(workspace, 16)),
               RL_EALIGN, "rl_gemm_ts_masked_f32: 16-byte aligned A, B, B_lo, C, out_mask, workspace required");
    int sms = rl::sm_count();
    if (sms <= 0) sms = 148;
    const cudaError_t e = rl::gts::launch(A, 0, B, B_lo, nullptr, C, 0, M, N, K, 0, static_cast<float*>(workspace), sms,
                                          rl::as_stream(stream), out_mask);
    RL_REQUIRE(e != cudaErrorInvalidValue, RL_EINVAL, "rl_gemm_ts_masked_f32: cuTensorMapEncodeTiled unavailable or failed");
    return rl::check_launch("gemm_ts_kernel");
}

int rl_split_lo_f32(const float* src, float* lo, int64_t n, void* stream) {
    if (n == 0) return RL_OK;
    RL_REQUIRE(src && lo && n > 0, RL_EINVAL, "rl_split_lo_f32: null pointer or negative n");
    int sms = rl::sm_count();
    if (sms <= 0) sms = 148;
    int64_t blocks = (n + 255) / 256;
    if (blocks > 8LL * sms) blocks = 8LL * sms;
    rl::gts::split_lo_kernel<<<static_cast<unsigned>(blocks), 256, 0, rl::as_stream(stream)>>>(src, lo, n);
    return rl::check_launch("split_lo_kernel");
}

int rl_transpose_split_f32(const float* src, float* dst, float* dst_lo, int64_t rows, int64_t cols, void* stream) {
    if (rows == 0 || cols == 0) return RL_OK;
    RL_REQUIRE(src && dst && dst_lo && rows > 0 && cols > 0, RL_EINVAL, "rl_transpose_split_f32: null pointer or negative extent");
    const int64_t tiles_r = (rows + 31) / 32, tiles_c = (cols + 31) / 32;
    RL_REQUIRE(tiles_r * tiles_c < (int64_t(1) << 31), RL_EINVAL, "rl_transpose_split_f32: matrix too large");
    rl::gts::transpose_split_kernel<<<static_cast<unsigned>(tiles_r * tiles_c), 256, 0, rl::as_stream(stream)>>>(
        src, dst, dst_lo, rows, cols, tiles_c);
    return rl::check_launch("transpose_split_kernel");
}

}  // extern "C"
