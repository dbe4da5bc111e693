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
// scalars [T,B].  Output observation [n, nf, H*W] u8.  HBM-bound: 2*n*nf*H*W read + the same written
// (28.9 MB + 28.9 MB at n=512, nf=4, 84x84); the scalar gathers add ~20 KB.
//
// Two implementations of the frame movement:
// * bulk (default when H*W is a multiple of 16 and the stack fits shared memory): one CTA per (sample, obs|target)
//   brings the nf frames of a stack into shared memory with nf concurrent cp.async.bulk copies (28 KB in flight per
//   CTA, 7 CTAs per SM: the whole n=512 batch is ONE wave of 1024 CTAs with ~200 KB in flight per SM), zero-fills
//   the blanked frames there, and writes the contiguous [nf, H*W] stack back with ONE bulk store.  No per-byte
//   instructions at all: the copy engine moves the data, one thread per CTA issues it.
// * vector (fallback: odd frame sizes): one CTA per (sample, obs|target, frame) with 16-byte ld/st when aligned,
//   bytes otherwise.
// A blanked frame is never read in either.
#include <stdlib.h>

#include "tc_common.cuh"

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

// One CTA moves one frame (hw bytes), or writes zeros when it is blanked - a blanked frame is never read.
__device__ __forceinline__ void copy_frame(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src, int64_t hw, bool blank) {
    const bool vec = (hw % 16 == 0) && aligned_dev(dst) && aligned_dev(src);
    if (vec) {
        const int64_t nv = hw / 16;
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
        for (int64_t j = threadIdx.x; j < hw; j += kExtractThreads) dst[j] = blank ? uint8_t(0) : src[j];
    }
}

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
    copy_frame(dst, src, v.hw, blank);

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


// ---- bulk-copy frame movement ----------------------------------------------------------------------------------
constexpr int kBulkThreads = 128;

__device__ __forceinline__ void bulk_store_g(void* dst_global, uint32_t src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(reinterpret_cast<uint64_t>(dst_global)),
                 "r"(src_smem), "r"(bytes) : "memory");
}

// The nf frames of the observation at ring time t, column b -> dst[nf*hw] (contiguous).  Called by every thread of a
// kBulkThreads CTA; `stack` = nf*hw bytes of 128-byte aligned shared memory, `bar` / `mask_slot` in static shared memory.
__device__ __forceinline__ void move_stack_bulk(uint8_t* __restrict__ dst, const ReplayView& v, int64_t t, int64_t b,
                                                uint8_t* stack, uint64_t* bar, uint32_t* mask_slot) {
    using namespace tc;
    const int tid = threadIdx.x, lane = tid & 31;
    const uint32_t hw = static_cast<uint32_t>(v.hw);
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid < 32) {
        // frame f (0 = oldest) is blank iff done[t-k] for some k in 1..nf-1-f   (frame.py:26-29)
        bool blank = false;
        if (lane < v.nf)
            for (int k = 1; k <= v.nf - 1 - lane; ++k) {
                const int64_t tk = ((t - k) % v.T + v.T) % v.T;                 // numpy negative index
                blank = blank || (v.done[tk * v.B + b] != 0);
            }
        const uint32_t mask = __ballot_sync(0xffffffffu, blank);
        if (lane == 0) {
            *mask_slot = mask;
            mbar_expect_tx(bar, static_cast<uint32_t>(v.nf - __popc(mask)) * hw);
        }
        __syncwarp();
        if (lane < v.nf && !blank)
            bulk_load(smem_u32(stack) + lane * hw, v.frames + ((t + lane) * v.B + b) * v.hw, hw, bar);
    }
    __syncthreads();
    const uint32_t mask = *mask_slot;
    if (mask != 0) {
        const uint4 zero = make_uint4(0, 0, 0, 0);
        for (int f = 0; f < v.nf; ++f)
            if ((mask >> f) & 1u)
                for (uint32_t j = tid; j < hw / 16; j += kBulkThreads) reinterpret_cast<uint4*>(stack + f * hw)[j] = zero;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");           // generic-proxy zeros -> visible to the bulk store
    }
    __syncthreads();
    if (tid == 0) {
        mbar_wait(bar, 0);
        bulk_store_g(dst, smem_u32(stack), static_cast<uint32_t>(v.nf) * hw);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");          // shared memory must outlive the read
    }
}

__global__ void __launch_bounds__(kBulkThreads)
replay_extract_bulk_kernel(ReplayView v, ReplayOut o, const int64_t* __restrict__ T_idx,
                           const int64_t* __restrict__ B_idx, int64_t n) {
    extern __shared__ __align__(128) uint8_t stack[];
    __shared__ uint64_t bar;
    __shared__ uint32_t mask_slot;
    const int64_t blk = blockIdx.x;
    const int which = static_cast<int>(blk & 1);                               // 0: observation, 1: target observation
    const int64_t i = blk >> 1;
    const int64_t t0 = T_idx[i];
    const int64_t b = B_idx[i];
    const int64_t t = which == 0 ? t0 : (t0 + v.n_step) % v.T;                 // n_step.py:23
    if (threadIdx.x == 64) {                                                   // scalar fields: a thread of an otherwise idle warp
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
    move_stack_bulk((which == 0 ? o.obs : o.target_obs) + i * v.nf * v.hw, v, t, b, stack, &bar, &mask_slot);
}

// ---- sequence extraction (R2D1 replay) ---------------------------------------------------------------------------
// Reference (restated in oracle/replay_sequence.py): rlpyt/replays/sequence/n_step.py:68-101 (extract_batch),
// rlpyt/replays/sequence/frame.py:18-50 (observation sequences from single frames; frames of the previous episode
// zeroed) and rlpyt/utils/misc.py:38-56 (extract_sequences).  Sample i starts at ring time t = T_idx[i], column
// b = B_idx[i]; with L = seq_T + n_step:
//   all_observation [L, n, nf, hw]   position j = the frame stack of ring time (t + j) % T
//   all_action, all_reward [L, n]    sequences started at t - 1 (they begin with prev_action / prev_reward)
//   return_, done, done_n [seq_T, n] sequences started at t
// extract_sequences' handling of a NEGATIVE start (t - 1 = -1) is kept as it is: the first L + start positions read
// rows 0.., the last -start positions read the ring's last rows (misc.py:49-51) - i.e. the wrapped row lands at the
// end of the sequence, not at its beginning.
__device__ __forceinline__ int64_t sequence_row(int64_t start, int64_t j, int64_t len, int64_t ring) {
    if (start + len > ring) return ((start + j) % ring + ring) % ring;     // wrap at the end
    if (start < 0) return j < len + start ? j : ring - len + j;            // misc.py:49-51
    return start + j;
}

__global__ void __launch_bounds__(kExtractThreads)
replay_extract_seq_frames_kernel(ReplayView v, uint8_t* __restrict__ out, const int64_t* __restrict__ T_idx,
                                 const int64_t* __restrict__ B_idx, int64_t n, int64_t L) {
    const int64_t blk = blockIdx.x;
    const int c = static_cast<int>(blk % v.nf);
    const int64_t i = (blk / v.nf) % n;
    const int64_t j = blk / (v.nf * n);
    const int64_t b = B_idx[i];
    const int64_t t = (T_idx[i] + j) % v.T;
    bool blank = false;                                                     // frame.py:39-48
    for (int k = 1; k <= v.nf - 1 - c; ++k) {
        const int64_t tk = ((t - k) % v.T + v.T) % v.T;
        blank = blank || (v.done[tk * v.B + b] != 0);
    }
    copy_frame(out + ((j * n + i) * v.nf + c) * v.hw, v.frames + ((t + c) * v.B + b) * v.hw, v.hw, blank);
}

__global__ void __launch_bounds__(kBulkThreads)
replay_extract_seq_frames_bulk_kernel(ReplayView v, uint8_t* __restrict__ out, const int64_t* __restrict__ T_idx,
                                      const int64_t* __restrict__ B_idx, int64_t n, int64_t L) {
    extern __shared__ __align__(128) uint8_t stack[];
    __shared__ uint64_t bar;
    __shared__ uint32_t mask_slot;
    const int64_t blk = blockIdx.x;
    const int64_t i = blk % n, j = blk / n;
    const int64_t t = (T_idx[i] + j) % v.T;
    move_stack_bulk(out + (j * n + i) * v.nf * v.hw, v, t, B_idx[i], stack, &bar, &mask_slot);
}

__global__ void __launch_bounds__(256)
replay_extract_seq_scalars_kernel(ReplayView v, const int64_t* __restrict__ T_idx, const int64_t* __restrict__ B_idx,
                                  int64_t n, int64_t seq_T, int64_t L, int64_t* __restrict__ all_action,
                                  float* __restrict__ all_reward, float* __restrict__ return_, uint8_t* __restrict__ done,
                                  uint8_t* __restrict__ done_n) {
    const int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= L * n) return;
    const int64_t j = e / n, i = e % n;
    const int64_t t = T_idx[i], b = B_idx[i];
    const int64_t rp = sequence_row(t - 1, j, L, v.T) * v.B + b;
    all_action[e] = v.action[rp];
    all_reward[e] = v.reward[rp];
    if (j < seq_T) {
        const int64_t r = sequence_row(t, j, seq_T, v.T) * v.B + b;
        return_[e] = v.return_[r];
        done[e] = v.done[r];
        done_n[e] = v.done_n[r];
    }
}

// Bulk copies need 16-byte aligned addresses and sizes; the stack of one observation must fit shared memory.
static bool bulk_eligible(const void* frames, const void* out_a, const void* out_b, int64_t frame_bytes, int n_frames) {
    if (getenv("RLPYT_B200_REPLAY_VECTOR") != nullptr) return false;          // cross-check switch for the tests
    const auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    return frame_bytes % 16 == 0 && n_frames >= 1 && n_frames <= 32 && al(frames) && al(out_a) && al(out_b) &&
           static_cast<int64_t>(n_frames) * frame_bytes <= 200 * 1024;
}

static int bulk_smem_attr(const void* kernel, size_t smem) {
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        RL_REQUIRE(e == cudaSuccess, static_cast<int>(e), "cudaFuncSetAttribute(shared memory %zu): %s", smem, cudaGetErrorString(e));
    }
    return RL_OK;
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
    if (rl::bulk_eligible(frames, out_obs, out_target_obs, frame_bytes, n_frames)) {
        const size_t smem = static_cast<size_t>(n_frames) * static_cast<size_t>(frame_bytes);
        int rc = rl::bulk_smem_attr(reinterpret_cast<const void*>(rl::replay_extract_bulk_kernel), smem);
        if (rc != RL_OK) return rc;
        rl::replay_extract_bulk_kernel<<<static_cast<unsigned>(2 * n), rl::kBulkThreads, smem, rl::as_stream(stream)>>>(
            v, o, T_idx, B_idx, n);
        return rl::check_launch("replay_extract_bulk_kernel");
    }
    rl::replay_extract_kernel<<<static_cast<unsigned>(blocks), rl::kExtractThreads, 0, rl::as_stream(stream)>>>(
        v, o, T_idx, B_idx, n);
    return rl::check_launch("replay_extract_kernel");
}

int rl_replay_extract_sequences(const uint8_t* frames, const int64_t* action, const float* reward, const uint8_t* done,
                                const float* return_, const uint8_t* done_n, int64_t T, int64_t B, int64_t frame_bytes,
                                int n_frames, int n_step, const int64_t* T_idx, const int64_t* B_idx, int64_t n,
                                int64_t seq_T, uint8_t* out_all_obs, int64_t* out_all_action, float* out_all_reward,
                                float* out_return, uint8_t* out_done, uint8_t* out_done_n, void* stream) {
    RL_REQUIRE(frames && action && reward && done && return_ && done_n && T_idx && B_idx, RL_EINVAL,
               "rl_replay_extract_sequences: null input pointer");
    RL_REQUIRE(out_all_obs && out_all_action && out_all_reward && out_return && out_done && out_done_n, RL_EINVAL,
               "rl_replay_extract_sequences: null output pointer");
    RL_REQUIRE(T >= 1 && B >= 1 && frame_bytes >= 1 && n_frames >= 1 && n_step >= 1 && n >= 0 && seq_T >= 1 &&
                   seq_T + n_step <= T,
               RL_EINVAL, "rl_replay_extract_sequences: bad extent (sequences must be shorter than the ring)");
    if (n == 0) return RL_OK;
    const int64_t L = seq_T + n_step;
    rl::ReplayView v{frames, action, reward, done, return_, done_n, T, B, frame_bytes, n_frames, n_step};
    const int64_t blocks = L * n * n_frames;
    RL_REQUIRE(blocks < (1LL << 31), RL_EINVAL, "rl_replay_extract_sequences: batch too large");
    int rc;
    if (rl::bulk_eligible(frames, out_all_obs, out_all_obs, frame_bytes, n_frames)) {
        const size_t smem = static_cast<size_t>(n_frames) * static_cast<size_t>(frame_bytes);
        rc = rl::bulk_smem_attr(reinterpret_cast<const void*>(rl::replay_extract_seq_frames_bulk_kernel), smem);
        if (rc != RL_OK) return rc;
        rl::replay_extract_seq_frames_bulk_kernel<<<static_cast<unsigned>(L * n), rl::kBulkThreads, smem, rl::as_stream(stream)>>>(
            v, out_all_obs, T_idx, B_idx, n, L);
        rc = rl::check_launch("replay_extract_seq_frames_bulk_kernel");
    } else {
        rl::replay_extract_seq_frames_kernel<<<static_cast<unsigned>(blocks), rl::kExtractThreads, 0, rl::as_stream(stream)>>>(
            v, out_all_obs, T_idx, B_idx, n, L);
        rc = rl::check_launch("replay_extract_seq_frames_kernel");
    }
    if (rc != RL_OK) return rc;
    rl::replay_extract_seq_scalars_kernel<<<static_cast<unsigned>((L * n + 255) / 256), 256, 0, rl::as_stream(stream)>>>(
        v, T_idx, B_idx, n, seq_T, L, out_all_action, out_all_reward, out_return, out_done, out_done_n);
    return rl::check_launch("replay_extract_seq_scalars_kernel");
}

}  // extern "C"
