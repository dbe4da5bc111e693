// C ABI of the kind::i8 first-layer kernels (conv1_i8.cuh): forward and weight/bias gradient of
// img.float().mul_(1/255) -> Conv2d(4->16, k8, s4, p0) -> ReLU on uint8 frames
// (rlpyt/models/pg/atari_ff_model.py:50-53, rlpyt/models/conv2d.py:36-44), with the minibatch row gather of
// rlpyt/algos/pg/ppo.py:99-100 applied per frame by the loader.
#include "conv1_i8.cuh"

using namespace rl::c1i8;

extern "C" {

int rl_conv1_u8_i8_supported(int C, int H, int W) {
    if (!geom_ok(C, H, W)) return 0;
    const Geom g = make_geom(1, H, W);
    return (smem_ok(g) && wg::smem_ok(g)) ? 1 : 0;
}

int rl_conv1_u8_forward_i8(const uint8_t* obs, const int64_t* rows, const float* weight, const float* bias, float* out,
                           int64_t N, int C, int H, int W, int relu, void* stream) {
    RL_REQUIRE(obs && weight && bias && out, RL_EINVAL, "rl_conv1_u8_forward_i8: null pointer");
    RL_REQUIRE(N >= 0 && N < (int64_t(1) << 31) && geom_ok(C, H, W), RL_EINVAL,
               "rl_conv1_u8_forward_i8: needs C=4, H %% 4 == 0, W %% 4 == 0, W <= 128 (got C=%d H=%d W=%d)", C, H, W);
    RL_REQUIRE(rl::aligned(obs, 16), RL_EALIGN, "rl_conv1_u8_forward_i8: frames must be 16-byte aligned (bulk copies)");
    if (N == 0) return RL_OK;
    const Geom g = make_geom(N, H, W);
    RL_REQUIRE(smem_ok(g), RL_EINVAL, "rl_conv1_u8_forward_i8: frame too large for the shared-memory ring (%dx%d)", H, W);
    int sms = rl::sm_count();
    if (sms <= 0) sms = 148;
    const cudaError_t e = launch_fwd(obs, rows, weight, bias, out, g, relu, sms, rl::as_stream(stream));
    if (e != cudaSuccess) {
        rl::set_error("rl_conv1_u8_forward_i8: %s", cudaGetErrorString(e));
        return static_cast<int>(e);
    }
    return RL_OK;
}

int rl_conv1_u8_forward_i8_stream(const uint8_t* obs_host_mapped, const float* weight, const float* bias, float* out,
                                  uint8_t* obs_copy, int64_t N, int C, int H, int W, int relu, void* stream) {
    RL_REQUIRE(obs_host_mapped && weight && bias && out && obs_copy, RL_EINVAL, "rl_conv1_u8_forward_i8_stream: null pointer");
    RL_REQUIRE(N >= 0 && N < (int64_t(1) << 31) && geom_ok(C, H, W), RL_EINVAL,
               "rl_conv1_u8_forward_i8_stream: needs C=4, H %% 4 == 0, W %% 4 == 0, W <= 128 (got C=%d H=%d W=%d)", C, H, W);
    RL_REQUIRE(rl::aligned(obs_host_mapped, 16) && rl::aligned(obs_copy, 16), RL_EALIGN,
               "rl_conv1_u8_forward_i8_stream: frames must be 16-byte aligned (bulk copies)");
    if (N == 0) return RL_OK;
    const Geom g = make_geom(N, H, W);
    RL_REQUIRE(smem_ok(g), RL_EINVAL, "rl_conv1_u8_forward_i8_stream: frame too large for the shared-memory ring (%dx%d)", H, W);
    int sms = rl::sm_count();
    if (sms <= 0) sms = 148;
    const cudaError_t e = launch_fwd(obs_host_mapped, nullptr, weight, bias, out, g, relu, sms, rl::as_stream(stream), obs_copy);
    if (e != cudaSuccess) {
        rl::set_error("rl_conv1_u8_forward_i8_stream: %s", cudaGetErrorString(e));
        return static_cast<int>(e);
    }
    return RL_OK;
}

int64_t rl_conv1_u8_wgrad_i8_scratch_bytes(void) {
    int sms = rl::sm_count();
    if (sms <= 0) sms = 148;
    return static_cast<int64_t>(wg::scratch_bytes(sms));
}

static int wgrad_i8_impl(const uint8_t* obs, const int64_t* rows, const float* out, const float* grad_out,
                         float* grad_weight, float* grad_bias, int64_t N, int C, int H, int W, void* scratch,
                         void* stream, const float* chan_absmax) {
    RL_REQUIRE(obs && grad_out && grad_weight && scratch, RL_EINVAL, "rl_conv1_u8_wgrad_i8: null pointer");
    RL_REQUIRE(N >= 0 && geom_ok(C, H, W), RL_EINVAL,
               "rl_conv1_u8_wgrad_i8: needs C=4, H %% 4 == 0, W %% 4 == 0, W <= 128 (got C=%d H=%d W=%d)", C, H, W);
    RL_REQUIRE(rl::aligned(obs, 16) && rl::aligned(grad_out, 16) && rl::aligned(scratch, 16), RL_EALIGN,
               "rl_conv1_u8_wgrad_i8: frames, grad_out and scratch must be 16-byte aligned");
    int sms = rl::sm_count();
    if (sms <= 0) sms = 148;
    RL_REQUIRE(N <= static_cast<int64_t>(sms) * wg::kMaxFramesPerCta, RL_EINVAL,
               "rl_conv1_u8_wgrad_i8: at most %d frames per call (int32 accumulators)", sms * wg::kMaxFramesPerCta);
    const cudaStream_t st = rl::as_stream(stream);
    if (N == 0) {
        cudaMemsetAsync(grad_weight, 0, 16 * 4 * 8 * 8 * sizeof(float), st);
        if (grad_bias) cudaMemsetAsync(grad_bias, 0, 16 * sizeof(float), st);
        return rl::check_launch("rl_conv1_u8_wgrad_i8");
    }
    const Geom g = make_geom(N, H, W);
    RL_REQUIRE(wg::smem_ok(g), RL_EINVAL, "rl_conv1_u8_wgrad_i8: frame too large for the shared-memory rings (%dx%d)", H, W);
    const cudaError_t e = wg::launch_wgrad(obs, rows, out, grad_out, grad_weight, grad_bias, g, sms, scratch, st, chan_absmax);
    if (e != cudaSuccess) {
        rl::set_error("rl_conv1_u8_wgrad_i8: %s", cudaGetErrorString(e));
        return static_cast<int>(e);
    }
    return RL_OK;
}

int rl_conv1_u8_wgrad_i8(const uint8_t* obs, const int64_t* rows, const float* out, const float* grad_out,
                         float* grad_weight, float* grad_bias, int64_t N, int C, int H, int W, void* scratch,
                         void* stream) {
    return wgrad_i8_impl(obs, rows, out, grad_out, grad_weight, grad_bias, N, C, H, W, scratch, stream, nullptr);
}

int rl_conv1_u8_wgrad_i8_scaled(const uint8_t* obs, const int64_t* rows, const float* out, const float* grad_out,
                                const float* chan_absmax, float* grad_weight, float* grad_bias, int64_t N, int C, int H, int W,
                                void* scratch, void* stream) {
    RL_REQUIRE(chan_absmax != nullptr, RL_EINVAL, "rl_conv1_u8_wgrad_i8_scaled: null chan_absmax");
    return wgrad_i8_impl(obs, rows, out, grad_out, grad_weight, grad_bias, N, C, H, W, scratch, stream, chan_absmax);
}

}  // extern "C"
