// C ABI of the "v2" second-layer kernels (conv2_s2d.cuh): Conv2d(16->32, k4, s2, p1)(+bias)(+ReLU) forward and its
// input gradient on fp32 NCHW activations (rlpyt/models/conv2d.py:36-44, rlpyt/models/pg/atari_ff_model.py:31-35).
#include "conv2_s2d.cuh"

using namespace rl::c2s;

extern "C" {

int rl_conv2_s2d_supported(int C, int IH, int IW) {
    if (!geom_ok(C, IH, IW)) return 0;
    const Geom g = make_geom(1, IH, IW);
    return (dg::smem_ok(g) && wg2::smem_ok(g)) ? 1 : 0;
}

int rl_conv2_forward_s2d(const float* x, const float* weight, const float* bias, float* out, int64_t N, int C, int IH,
                         int IW, int relu, void* stream) {
    RL_REQUIRE(x && weight && bias && out, RL_EINVAL, "rl_conv2_forward_s2d: null pointer");
    RL_REQUIRE(N >= 0 && N < (int64_t(1) << 31) && geom_ok(C, IH, IW), RL_EINVAL,
               "rl_conv2_forward_s2d: needs C=16, OW <= 14 (got C=%d %dx%d)", C, IH, IW);
    RL_REQUIRE(rl::aligned(x, 16) && rl::aligned(weight, 4), RL_EALIGN, "rl_conv2_forward_s2d: x must be 16-byte aligned");
    if (N == 0) return RL_OK;
    int sms = rl::sm_count();
    if (sms <= 0) sms = 148;
    const cudaError_t e = launch_fwd(x, weight, bias, out, make_geom(N, IH, IW), relu, sms, rl::as_stream(stream));
    if (e != cudaSuccess) {
        rl::set_error("rl_conv2_forward_s2d: %s", cudaGetErrorString(e));
        return static_cast<int>(e);
    }
    return RL_OK;
}

int rl_conv2_dgrad_s2d(const float* grad_out_masked, const float* weight, float* grad_x, int64_t N, int C, int IH, int IW,
                       void* stream) {
    RL_REQUIRE(grad_out_masked && weight && grad_x, RL_EINVAL, "rl_conv2_dgrad_s2d: null pointer");
    RL_REQUIRE(N >= 0 && N < (int64_t(1) << 31) && geom_ok(C, IH, IW), RL_EINVAL,
               "rl_conv2_dgrad_s2d: needs C=16, OW <= 14 (got C=%d %dx%d)", C, IH, IW);
    const Geom g = make_geom(N, IH, IW);
    RL_REQUIRE(dg::smem_ok(g), RL_EINVAL, "rl_conv2_dgrad_s2d: plane too large for the shared-memory stages (%dx%d)", IH, IW);
    RL_REQUIRE(rl::aligned(grad_out_masked, 16) && rl::aligned(grad_x, 16), RL_EALIGN,
               "rl_conv2_dgrad_s2d: gradients must be 16-byte aligned (bulk copies)");
    if (N == 0) return RL_OK;
    int sms = rl::sm_count();
    if (sms <= 0) sms = 148;
    const cudaError_t e = dg::launch_dgrad(grad_out_masked, weight, grad_x, g, sms, rl::as_stream(stream));
    if (e != cudaSuccess) {
        rl::set_error("rl_conv2_dgrad_s2d: %s", cudaGetErrorString(e));
        return static_cast<int>(e);
    }
    return RL_OK;
}

int rl_conv2_dgrad_s2d_absmax(const float* grad_out_masked, const float* weight, float* grad_x, float* chan_absmax, int64_t N,
                              int C, int IH, int IW, void* stream) {
    RL_REQUIRE(grad_out_masked && weight && grad_x && chan_absmax, RL_EINVAL, "rl_conv2_dgrad_s2d_absmax: null pointer");
    RL_REQUIRE(N >= 0 && N < (int64_t(1) << 31) && geom_ok(C, IH, IW), RL_EINVAL,
               "rl_conv2_dgrad_s2d_absmax: needs C=16, OW <= 14 (got C=%d %dx%d)", C, IH, IW);
    const Geom g = make_geom(N, IH, IW);
    RL_REQUIRE(dg::smem_ok(g), RL_EINVAL, "rl_conv2_dgrad_s2d_absmax: plane too large for the shared-memory stages (%dx%d)", IH, IW);
    RL_REQUIRE(rl::aligned(grad_out_masked, 16) && rl::aligned(grad_x, 16) && rl::aligned(chan_absmax, 4), RL_EALIGN,
               "rl_conv2_dgrad_s2d_absmax: gradients must be 16-byte aligned (bulk copies)");
    int sms = rl::sm_count();
    if (sms <= 0) sms = 148;
    if (N == 0) {
        cudaMemsetAsync(chan_absmax, 0, 16 * sizeof(float), rl::as_stream(stream));
        return rl::check_launch("rl_conv2_dgrad_s2d_absmax");
    }
    const cudaError_t e = dg::launch_dgrad(grad_out_masked, weight, grad_x, g, sms, rl::as_stream(stream), chan_absmax);
    if (e != cudaSuccess) {
        rl::set_error("rl_conv2_dgrad_s2d_absmax: %s", cudaGetErrorString(e));
        return static_cast<int>(e);
    }
    return RL_OK;
}

int64_t rl_conv2_wgrad_s2d_scratch_bytes(void) {
    int sms = rl::sm_count();
    if (sms <= 0) sms = 148;
    return static_cast<int64_t>(wg2::scratch_bytes(sms));
}

int rl_conv2_wgrad_s2d(const float* x, const float* grad_out_masked, float* grad_weight, float* grad_bias, int64_t N, int C,
                       int IH, int IW, void* scratch, void* stream) {
    RL_REQUIRE(x && grad_out_masked && grad_weight && scratch, RL_EINVAL, "rl_conv2_wgrad_s2d: null pointer");
    RL_REQUIRE(N >= 0 && N < (int64_t(1) << 31) && geom_ok(C, IH, IW), RL_EINVAL,
               "rl_conv2_wgrad_s2d: needs C=16, OW <= 14 (got C=%d %dx%d)", C, IH, IW);
    const Geom g = make_geom(N, IH, IW);
    RL_REQUIRE(wg2::smem_ok(g), RL_EINVAL, "rl_conv2_wgrad_s2d: plane too large for the shared-memory stages (%dx%d)", IH, IW);
    RL_REQUIRE(rl::aligned(x, 16) && rl::aligned(grad_out_masked, 16) && rl::aligned(scratch, 16), RL_EALIGN,
               "rl_conv2_wgrad_s2d: x, grad_out and scratch must be 16-byte aligned");
    const cudaStream_t st = rl::as_stream(stream);
    if (N == 0) {
        cudaMemsetAsync(grad_weight, 0, 32 * 16 * 16 * sizeof(float), st);
        if (grad_bias) cudaMemsetAsync(grad_bias, 0, 32 * sizeof(float), st);
        return rl::check_launch("rl_conv2_wgrad_s2d");
    }
    int sms = rl::sm_count();
    if (sms <= 0) sms = 148;
    const cudaError_t e = wg2::launch_wgrad(x, grad_out_masked, grad_weight, grad_bias, g, sms, scratch, st);
    if (e != cudaSuccess) {
        rl::set_error("rl_conv2_wgrad_s2d: %s", cudaGetErrorString(e));
        return static_cast<int>(e);
    }
    return RL_OK;
}

}  // extern "C"
