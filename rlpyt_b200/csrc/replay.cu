// K8: batch extraction from the frame-wise n-step replay buffer - frame-stack gather with
// done-blanking for the observation AND the n-step target observation, plus every scalar field,
// in one launch.
//
// Reference (restated in oracle/replay.py): rlpyt/replays/non_sequence/frame.py:14-30
// (extract_observation: stack frames[t:t+nf, b]; for f in 1..nf-1 zero obs[:nf-f] where
// done[t-f, b]) and rlpyt/replays/non_sequence/n_step.py:16-43 (extract_batch: prev_action /
// prev_reward at t-1 zeroed where done[t-1]; action, return_, done, done_n at t; target inputs at
// (t+n) % T; numpy negative indices wrap, hence the explicit mod below).
//
// Layout: frames [T+nf-1, B, H*W] u8 (frame.py:39-43: the observation of time t is frames[t:t+nf]);
// scalars [T,B].  Output observation [n, nf, H*W] u8.  One CTA per (sample, obs|target, frame)
// copies H*W bytes (7056 B = 441 uint4 at 84x84) with 16-byte accesses, or writes zeros when the
// frame is blanked - so blanked frames are never read.  HBM-bound: 2*n*nf*H*W read + the same
// written (28.9 MB + 28.9 MB at n=512, nf=4, 84x84); the scalar gathers add ~20 KB.
#include "common.cuh"

namespace rl {

struct ReplayView {
    const uint8_t* frames;      // [T+nf-1, B, hw]
    const int64_t* action;      // [T,B]
    const float* reward;        // [T,B]
    const uint8_t* done;        // [T,B]
    const float* return_;       // [T,B]
    const uint8_t* done_n;      // [T,B]
    int64_t T, B, hw;
    int nf, n_step;
};

struct ReplayOut {
    uint8_t* obs;               // [n, nf, hw]
    uint8_t* target_obs;        // [n, nf, hw]
    int64_t* prev_action; float* prev_reward;
    int64_t* action; float* return_; uint8_t* done; uint8_t* done_n;
    int64_t* target_prev_action; float* target_prev_reward;
};

constexpr int kExtractThreads = 128;

__global__ void __launch_bounds__(kExtractThreads)
replay_extract_kernel(ReplayView v, ReplayOut o, const int64_t* __restrict__ T_idx,
                      const int64_t* __restrict__ B_idx, int64_t n) {
    const int64_t blk = blockIdx.x;
    const int f = static_cast<int>(blk % v.nf);
    const int which = static_cast<int>((blk / v.nf) % 2);   // 0: observation, 1: target observation
    const int64_t i = blk / (2 * v.nf);
    const int64_t t0 = T_idx[i];
    const int64_t b = B_idx[i];
    const int64_t t = which == 0 ? t0 : (t0 + v.n_step) % v.T;                 // n_step.py:23

    // frame f (0 = oldest) is blank iff done[t-k] for some k in 1..nf-1-f   (frame.py:26-29)
    bool blank = false;
    for (int k = 1; k <= v.nf - 1 - f; ++k) {
        const int64_t tk = ((t - k) % v.T + v.T) % v.T;                         // numpy negative index
        blank = blank || (v.done[tk * v.B + b] != 0);
    }
    uint8_t* dst = (which == 0 ? o.obs : o.target_obs) + (i * v.nf + f) * v.hw;
    const uint8_t* src = v.frames + ((t + f) * v.B + b) * v.hw;
    const bool vec = (v.hw % 16 == 0) && aligned_dev(dst) && aligned_dev(src);
    if (vec) {
        const int64_t nv = v.hw / 16;
        const uint4 zero = make_uint4(0, 0, 0, 0);
        constexpr int kU = 4;  // independent 16 B loads in flight per thread (441 uint4 per 84x84 frame)
        for (int64_t j0 = threadIdx.x; j0 < nv; j0 += kExtractThreads * kU) {
            uint4 x[kU];
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                const int64_t j = j0 + static_cast<int64_t>(u) * kExtractThreads;
                x[u] = (blank || j >= nv) ? zero : ldg_stream(reinterpret_cast<const uint4*>(src) + j);
            }
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                const int64_t j = j0 + static_cast<int64_t>(u) * kExtractThreads;
                if (j < nv) stg_stream(reinterpret_cast<uint4*>(dst) + j, x[u]);
            }
        }
    } else {
        for (int64_t j = threadIdx.x; j < v.hw; j += kExtractThreads) dst[j] = blank ? uint8_t(0) : src[j];
    }

    if (f == 0 && threadIdx.x == 0) {
        const int64_t tm1 = ((t - 1) % v.T + v.T) % v.T;
        const int64_t at = t * v.B + b, am1 = tm1 * v.B + b;
        if (which == 0) {
            const bool is_new = v.done[am1] != 0;                               // n_step.py:40-42
            o.prev_action[i] = is_new ? 0 : v.action[am1];
            o.prev_reward[i] = is_new ? 0.0f : v.reward[am1];
            o.action[i] = v.action[at];
            o.return_[i] = v.return_[at];
            o.done[i] = v.done[at];
            o.done_n[i] = v.done_n[at];
        } else {
            o.target_prev_action[i] = v.action[am1];                            // n_step.py:36-37 (not zeroed)
            o.target_prev_reward[i] = v.reward[am1];
        }
    }
}

}  // namespace rl

extern "C" {

int rl_replay_extract(const uint8_t* frames, const int64_t* action, const float* reward, const uint8_t* done,
                      const float* return_, const uint8_t* done_n, int64_t T, int64_t B, int64_t frame_bytes,
                      int n_frames, int n_step, const int64_t* T_idx, const int64_t* B_idx, int64_t n,
                      uint8_t* out_obs, uint8_t* out_target_obs, int64_t* out_prev_action,
                      float* out_prev_reward, int64_t* out_action, float* out_return, uint8_t* out_done,
                      uint8_t* out_done_n, int64_t* out_target_prev_action, float* out_target_prev_reward,
                      void* stream) {
    RL_REQUIRE(frames && action && reward && done && return_ && done_n && T_idx && B_idx, RL_EINVAL,
               "rl_replay_extract: null input pointer");
    RL_REQUIRE(out_obs && out_target_obs && out_prev_action && out_prev_reward && out_action && out_return &&
                   out_done && out_done_n && out_target_prev_action && out_target_prev_reward,
               RL_EINVAL, "rl_replay_extract: null output pointer");
    RL_REQUIRE(T >= 1 && B >= 1 && frame_bytes >= 1 && n_frames >= 1 && n_step >= 1 && n >= 0, RL_EINVAL,
               "rl_replay_extract: bad extent");
    if (n == 0) return RL_OK;
    rl::ReplayView v{frames, action, reward, done, return_, done_n, T, B, frame_bytes, n_frames, n_step};
    rl::ReplayOut o{out_obs, out_target_obs, out_prev_action, out_prev_reward, out_action, out_return,
                    out_done, out_done_n, out_target_prev_action, out_target_prev_reward};
    const int64_t blocks = n * 2 * n_frames;
    RL_REQUIRE(blocks < (1LL << 31), RL_EINVAL, "rl_replay_extract: batch too large");
    rl::replay_extract_kernel<<<static_cast<unsigned>(blocks), rl::kExtractThreads, 0, rl::as_stream(stream)>>>(
        v, o, T_idx, B_idx, n);
    return rl::check_launch("replay_extract_kernel");
}

}  // extern "C"
