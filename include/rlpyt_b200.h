/*
 * rlpyt_b200.h - C ABI of librlpyt_b200.so: the B200 (sm_100a) kernels behind rlpyt's
 * data-parallel inner loop.
 *
 * rlpyt has no FFI layer: its "plugin API" is a set of duck-typed Python classes handed
 * to the runner (rlpyt/runners/minibatch_rl.py:33-46).  Every entry point below replaces
 * the ARITHMETIC of one reference function / method (cited per function, paths relative
 * to the reference checkout); the Python classes in rlpyt_b200/ keep the reference
 * signatures and call these through ctypes (INTEGRATION.md shows the binding).
 *
 * Conventions (all functions):
 *   - plain C: raw DEVICE pointers + extents + a stream handle (cudaStream_t passed as
 *     void*; NULL = legacy default stream).  No torch / C++ types cross the boundary.
 *   - buffers are caller-owned; nothing is allocated or freed behind the caller's back;
 *     scratch space is passed explicitly (sizes given by rl_*_scratch_bytes helpers).
 *   - stream-ordered, no host synchronisation inside, re-entrant, no global state other
 *     than the per-thread last-error string.
 *   - returns 0 on success, a negative RL_E* code on bad arguments, or a positive
 *     cudaError_t value if a launch failed.  rl_b200_last_error() returns the text.
 *   - layouts are the reference's: [T,B] time-major C-contiguous (element (t,b) at
 *     t*B+b), bool/done as 1 byte 0/1, actions and indices int64.
 */
#ifndef RLPYT_B200_H
#define RLPYT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RL_OK 0
#define RL_EINVAL (-1)   /* bad extent / null pointer / unsupported combination */
#define RL_EALIGN (-2)   /* pointer alignment requirement violated */

/* ABI version (bumped on any signature change) and last error text (thread-local). */
int rl_b200_abi_version(void);
const char* rl_b200_last_error(void);
/* Number of SMs of the current device (148 on B200); <0 on error.  Host-side helper. */
int rl_b200_sm_count(void);

/* HOST helper for the env workers (rlpyt/samplers/parallel/gpu/collectors.py:38 writes
 * step.observation[b] = o): copy nbytes from src to the page-locked step buffer with
 * non-temporal stores so the GPU's DMA does not have to snoop the worker cores' caches.
 * Host pointers; no CUDA call inside; safe in forked worker processes. */
int rl_host_stream_copy(void* dst, const void* src, int64_t nbytes);

/* Host (page-locked) -> device copy of one env worker's observation rows, stream-ordered (cudaMemcpyAsync): the
 * master half of rlpyt/samplers/parallel/gpu/action_server.py:46-48, issued per worker as each one signals
 * obs_ready instead of once for the whole step buffer. */
int rl_upload_async(void* dst_device, const void* src_host, int64_t nbytes, void* stream);

/* ------------------------------------------------------------------ returns (K1-K4)
 * algo: 0 = auto, 1 = streaming column kernel (thread per 1/4 columns, sequential in t,
 *       reference operation order => bit-identical to the reference),
 *       2 = T-parallel warp-segmented scan (small B; reassociated => <=1e-5 rel).
 */

/* generalized_advantage_estimation - rlpyt/algos/utils.py:24-40.
 * advantage[t] = delta_t + (discount*lambda)*nd_t*advantage[t+1], return_ = advantage+value.
 * gamma_lambda must be (float)((double)discount*(double)gae_lambda) (utils.py:38).
 * done: [T,B] uint8 (0/1); bootstrap_value: [B]; outputs [T,B]. */
int rl_gae_f32(const float* reward, const float* value, const uint8_t* done,
               const float* bootstrap_value, float* advantage, float* return_,
               int T, int64_t B, float discount, float gamma_lambda, int algo, void* stream);

/* discount_return - rlpyt/algos/utils.py:8-21.  If value!=NULL also writes
 * advantage = return_ - value (rlpyt/algos/pg/base.py:54-55); else advantage may be NULL. */
int rl_discount_return_f32(const float* reward, const uint8_t* done,
                           const float* bootstrap_value, const float* value,
                           float* return_, float* advantage,
                           int T, int64_t B, float discount, int algo, void* stream);

/* discount_return_n_step - rlpyt/algos/utils.py:67-101.  reward/done: [T_in,B];
 * outputs [rlen,B] with rlen = do_truncated ? T_in : T_in-(n_step-1).
 * discount_pow[k] = (float)pow((double)discount,k), k=0..n_step-1 (host-computed so the
 * python double pow of utils.py:97 is reproduced), DEVICE pointer. */
int rl_nstep_return_f32(const float* reward, const uint8_t* done, const float* discount_pow,
                        float* return_, uint8_t* done_n,
                        int T_in, int64_t B, int n_step, int do_truncated, void* stream);

/* valid_from_done - rlpyt/algos/utils.py:104-112.  valid: [T,B] float32. */
int rl_valid_from_done_f32(const uint8_t* done, float* valid, int T, int64_t B, void* stream);

/* Advantage normalisation of process_returns - rlpyt/algos/pg/base.py:65-73:
 * adv = (adv - mean)/max(std,1e-6), unbiased std, statistics over valid>0 (valid may be
 * NULL = all).  In place.  scratch: rl_adv_normalize_scratch_bytes(n) bytes, 8B aligned.
 * stats_out (nullable): 3 floats {mean, std, count} for logging/tests. */
int64_t rl_adv_normalize_scratch_bytes(int64_t n);
int rl_adv_normalize_f32(float* advantage, const float* valid, int64_t n,
                         void* scratch, float* stats_out, void* stream);

/* ------------------------------------------------------------------ policy-gradient loss (K5)
 * Fused forward + gradient of the PPO / A2C loss w.r.t. the network outputs.
 *   PPO.loss  - rlpyt/algos/pg/ppo.py:136-153, A2C.loss - rlpyt/algos/pg/a2c.py:88-100,
 *   Categorical.likelihood_ratio/log_likelihood/entropy - rlpyt/distributions/categorical.py:32-43,
 *   valid_mean - rlpyt/utils/tensor.py:39-46, mean_entropy/mean_perplexity - distributions/base.py:57-68.
 * prob_*: [N,A] f32 row-major, action: [N] int64, value/return_/advantage: [N] f32,
 * valid: [N] f32 or NULL (plain mean).  out_scalars: 8 floats
 *   {loss, entropy, perplexity, pi_loss, value_loss, sum(valid) or N, 0, 0}.
 * grad_prob [N,A] / grad_value [N] (each nullable) receive dLoss/dprob_new, dLoss/dvalue
 * for an upstream gradient of 1.  scratch: rl_pg_loss_scratch_bytes(N) bytes, 8B aligned. */
int64_t rl_pg_loss_scratch_bytes(int64_t N);
int rl_ppo_loss_f32(const float* prob_new, const float* value, const float* prob_old,
                    const int64_t* action, const float* return_, const float* advantage,
                    const float* valid, int64_t N, int A, float ratio_clip,
                    float value_loss_coeff, float entropy_loss_coeff, float* out_scalars,
                    float* grad_prob, float* grad_value, void* scratch, void* stream);
/* rl_ppo_loss_f32 with the ratio clip read from DEVICE memory (one float): the linear schedule of
 * rlpyt/algos/pg/ppo.py:110-113 changes it every iteration, and a CUDA graph of the minibatch update captured once
 * must follow it without being re-captured. */
int rl_ppo_loss_devclip_f32(const float* prob_new, const float* value, const float* prob_old,
                            const int64_t* action, const float* return_, const float* advantage,
                            const float* valid, int64_t N, int A, const float* ratio_clip_dev,
                            float value_loss_coeff, float entropy_loss_coeff, float* out_scalars,
                            float* grad_prob, float* grad_value, void* scratch, void* stream);
int rl_a2c_loss_f32(const float* prob, const float* value, const int64_t* action,
                    const float* return_, const float* advantage, const float* valid,
                    int64_t N, int A, float value_loss_coeff, float entropy_loss_coeff,
                    float* out_scalars, float* grad_prob, float* grad_value, void* scratch,
                    void* stream);

/* ------------------------------------------------------------------ indexed row gathers
 * dst[i,:] = src[idx[i],:], rows of row_bytes bytes, idx int64 in [0, src_rows).
 * The PPO minibatch former: LossInputs[T_idxs, B_idxs] of rlpyt/algos/pg/ppo.py:94-100 with
 * idx = T_idx*B + B_idx into the [T*B, ...] view (observations: 28224 B rows), and the plain
 * "x[T_idxs, B_idxs]" field reads of rlpyt/replays/non_sequence/n_step.py:24-37. */
int rl_gather_rows(const void* src, const int64_t* idx, void* dst, int64_t n, int64_t row_bytes,
                   void* stream);
/* Up to RL_GATHER_MAX_FIELDS small fields (rows multiple of 4 B) with one index vector in one
 * launch.  src/dst/row_bytes are HOST arrays of n_fields entries (device pointers inside). */
#define RL_GATHER_MAX_FIELDS 8
int rl_gather_rows_multi(int n_fields, const void* const* src, void* const* dst,
                         const int64_t* row_bytes, const int64_t* idx, int64_t n, void* stream);

/* ------------------------------------------------------------------ optimizer step
 * Global-norm clip + Adam on one flat fp32 buffer: torch.nn.utils.clip_grad_norm_ +
 * torch.optim.Adam.step() as used at rlpyt/algos/pg/ppo.py:101-104 / a2c.py:50-53.
 * grad_scale multiplies the gradient first (1/world_size after a SUM all-reduce: the DDP
 * average of rlpyt/agents/base.py:118-136).  max_norm <= 0 disables clipping.  step >= 1 is
 * the Adam step count AFTER this update.  norm_out (nullable): pre-clip total norm (the
 * reference's gradNorm).  scratch: rl_clip_adam_scratch_bytes(n) bytes. */
int64_t rl_clip_adam_scratch_bytes(int64_t n);
int rl_clip_adam_f32(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                     float lr, float beta1, float beta2, float eps, float weight_decay, int64_t step,
                     float max_norm, float grad_scale, float* norm_out, void* scratch, void* stream);

/* ------------------------------------------------------------------ prioritized replay: sum tree (K6/K7)
 * fp64 implicit binary tree (rlpyt/replays/sum_tree.py:39-51): `levels` = tree_levels,
 * 2^levels - 1 nodes, leaves from 2^(levels-1)-1.  Caller-owned device array.  Bit-exact. */

/* SumTree.find + the index split of SumTree.sample - rlpyt/replays/sum_tree.py:211-222, :124-127.
 * uniforms: n doubles in [0,1] (np.random.rand on the host in the reference, :107).
 * Outputs (tree_idx required, others nullable): tree_idx, T_idx, B_idx = divmod(leaf, B),
 * priority = tree[tree_idx], scaled = tree[0]*u. */
int rl_sumtree_find_f64(const double* tree, int levels, const double* uniforms, int64_t n, int64_t B,
                        int64_t* tree_idx, int64_t* T_idx, int64_t* B_idx, double* priority,
                        double* scaled, void* stream);

/* SumTree.reconstruct / one segment of reconstruct_advance - sum_tree.py:150-153, :155-209:
 * writes new leaf values and adds (new - old) to every ancestor in the reference's (sequential,
 * array) order.  ONE ASCENDING segment of leaves per call: leaf_idx (device, ascending tree
 * indices, duplicates keep the first value) or, if NULL, the contiguous range
 * [leaf_base, leaf_base+n).  values (device doubles) or, if NULL, value_scalar for every leaf.
 * scratch_diffs: n doubles.  Segments of one reference call must be issued in reference order. */
int rl_sumtree_update_f64(double* tree, int levels, const int64_t* leaf_idx, int64_t leaf_base,
                          const double* values, double value_scalar, int64_t n, double* scratch_diffs,
                          void* stream);

/* update_batch_priorities of one sampled batch in one call (rlpyt/replays/sum_tree.py:130-138 with
 * prioritized.py:73-79 folded in): leaf_idx = the n (<= 2048) tree indices of the last sample(), in ANY order,
 * duplicates allowed (the first occurrence wins, np.unique(return_index=True)); new leaf values are either
 * (double)(float)pow(priorities_f32, alpha) - numpy's float32 `**` - or values_f64 as given (exactly one of the two
 * non-NULL).  The batch is sorted on the device; ancestors accumulate the differences in array order (np.add.at).
 * scratch_sorted_idx: n int64, scratch_diffs: n doubles. */
int rl_sumtree_update_batch(double* tree, int levels, const int64_t* leaf_idx, const float* priorities_f32, float alpha,
                            const double* values_f64, int64_t n, int64_t* scratch_sorted_idx, double* scratch_diffs,
                            void* stream);

/* priorities ** alpha as numpy float32 pow, widened to the tree's fp64
 * (rlpyt/replays/non_sequence/prioritized.py:73-79). */
int rl_pow_f32_to_f64(const float* x, float exponent, double* out, int64_t n, void* stream);

/* is_weights = (1/(p+1e-6))**beta / max -> float32 (prioritized.py:68-70).  n <= 2^20. */
int rl_is_weights_f32(const double* priority, double beta, float* out, int n, void* stream);
/* the same with an explicit epsilon: rlpyt/replays/sequence/prioritized.py:111-113 uses (1/p)**beta, eps = 0 */
int rl_is_weights_eps_f32(const double* priority, double beta, double eps, float* out, int n, void* stream);

/* ------------------------------------------------------------------ replay batch extraction (K8)
 * NStepReturnBuffer.extract_batch + NStepFrameBuffer.extract_observation -
 * rlpyt/replays/non_sequence/n_step.py:16-43, rlpyt/replays/non_sequence/frame.py:14-30.
 * frames: [T+n_frames-1, B, frame_bytes] u8 (rlpyt/replays/frame.py:39-43); action i64, reward f32,
 * done u8, return_ f32, done_n u8: [T,B].  T_idx/B_idx: n int64.  Outputs: observations
 * [n, n_frames, frame_bytes] u8 (oldest->newest, frames after an episode end blanked), scalars [n]. */
int rl_replay_extract(const uint8_t* frames, const int64_t* action, const float* reward,
                      const uint8_t* done, const float* return_, const uint8_t* done_n,
                      int64_t T, int64_t B, int64_t frame_bytes, int n_frames, int n_step,
                      const int64_t* T_idx, const int64_t* B_idx, int64_t n,
                      uint8_t* out_obs, uint8_t* out_target_obs, int64_t* out_prev_action,
                      float* out_prev_reward, int64_t* out_action, float* out_return,
                      uint8_t* out_done, uint8_t* out_done_n, int64_t* out_target_prev_action,
                      float* out_target_prev_reward, void* stream);

/* Sequence extraction for recurrent replay (R2D1): SequenceNStepReturnBuffer.extract_batch with
 * SequenceNStepFrameBuffer.extract_observation - rlpyt/replays/sequence/n_step.py:68-101,
 * rlpyt/replays/sequence/frame.py:18-50, rlpyt/utils/misc.py:38-56 (extract_sequences, incl. its placement of a
 * negative start index).  Same storage as rl_replay_extract.  With L = seq_T + n_step <= T, outputs (time-major):
 * all_obs [L, n, n_frames, frame_bytes] u8; all_action i64 / all_reward f32 [L, n] (start one step before T_idx);
 * return_ f32 / done u8 / done_n u8 [seq_T, n]. */
int rl_replay_extract_sequences(const uint8_t* frames, const int64_t* action, const float* reward, const uint8_t* done,
                                const float* return_, const uint8_t* done_n, int64_t T, int64_t B, int64_t frame_bytes,
                                int n_frames, int n_step, const int64_t* T_idx, const int64_t* B_idx, int64_t n,
                                int64_t seq_T, uint8_t* out_all_obs, int64_t* out_all_action, float* out_all_reward,
                                float* out_return, uint8_t* out_done, uint8_t* out_done_n, void* stream);

/* ------------------------------------------------------------------ AtariFf first layer on uint8 frames
 * img.float().mul_(1/255) -> Conv2d(4->16, k8, s4, p0) -> ReLU : rlpyt/models/pg/atari_ff_model.py:50-53,
 * rlpyt/models/conv2d.py:36-44, fused with the minibatch row gather of rlpyt/algos/pg/ppo.py:99-100.
 * obs: [R, 4, H, W] u8 (W % 4 == 0); rows: N int64 indices into R, or NULL (= 0..N-1);
 * weight [16,4,8,8], bias [16] f32; out [N,16,OH,OW] f32 with OH=(H-8)/4+1, OW=(W-8)/4+1.
 * relu != 0 applies max(.,0). */
int rl_conv1_u8_forward(const uint8_t* obs, const int64_t* rows, const float* weight, const float* bias,
                        float* out, int64_t N, int C, int H, int W, int relu, void* stream);
/* Weight / bias gradient of the same layer (what autograd's ConvolutionBackward + ThresholdBackward
 * produce for it): g = grad_out * (out > 0) if relu; grad_weight[16,4,8,8] = sum g (x) x/255,
 * grad_bias[16] = sum g.  Deterministic (fixed reduction order).  scratch:
 * rl_conv1_u8_wgrad_scratch_bytes() bytes. */
int64_t rl_conv1_u8_wgrad_scratch_bytes(void);
int rl_conv1_u8_wgrad(const uint8_t* obs, const int64_t* rows, const float* out, const float* grad_out,
                      float* grad_weight, float* grad_bias, int64_t N, int C, int H, int W, int relu,
                      void* scratch, void* stream);

/* ------------------------------------------------------------------ fp32-accurate tensor-core GEMM
 * C[M,N] = A[M,K] * B[N,K]^T (+ bias[N]) (+ ReLU): torch.nn.Linear as used by the AtariFf fully
 * connected layer (rlpyt/models/mlp.py:30-36 via rlpyt/models/pg/atari_ff_model.py:24-35) and its
 * dgrad / wgrad (with transposed operands).  tcgen05.mma kind::tf32 with a 3-term hi/lo split
 * (fp32-level accuracy), TMA-staged, accumulators in TMEM.  All row-major fp32, K %% 4 == 0,
 * 16-byte aligned bases; bias nullable. */
/* workspace: rl_gemm_tf32x3_workspace_bytes(M,N,K) bytes (0 = none needed; nullable): when the
 * tile grid cannot fill the SMs (agent.step: M = 256) K is split over up to 16 CTAs per tile and
 * the partial tiles are summed in order by a second kernel. */
int64_t rl_gemm_tf32x3_workspace_bytes(int64_t M, int64_t N, int64_t K);
int rl_gemm_tf32x3_f32(const float* A, const float* B, const float* bias, float* C, int64_t M, int64_t N,
                       int64_t K, int relu, void* workspace, void* stream);

/* Second version of the same product (rlpyt/models/mlp.py:30-36 forward, and the autograd backward of that
 * torch.nn.Linear: grad_input = grad_out W, grad_weight = grad_out^T x): the A operand is fed to tcgen05.mma from
 * tensor memory, one persistent CTA per SM.
 *   A      [M,K] row-major, or - a_mmajor != 0 - the [K,M] row-major matrix A^T (M % 4 == 0);
 *   B,B_lo [N,K] row-major: the operand and B - trunc_tf32(B) (rl_split_lo_f32 / rl_transpose_split_f32 write it);
 *   C      [M,N] row-major, or - c_trans != 0 - C^T as an [N,M] row-major matrix;  bias [N], nullable.
 * workspace: rl_gemm_ts_workspace_bytes(M,N,K) bytes (0 = none; nullable = never split K). */
int64_t rl_gemm_ts_workspace_bytes(int64_t M, int64_t N, int64_t K);
int rl_gemm_ts_f32(const float* A, int a_mmajor, const float* B, const float* B_lo, const float* bias, float* C,
                   int c_trans, int64_t M, int64_t N, int64_t K, int relu, void* workspace, void* stream);
/* The input gradient of a Linear layer whose INPUT is the output of a ReLU (rlpyt/models/conv2d.py:41 feeding
 * rlpyt/models/mlp.py:30-36): C[M,N] = (A[M,K] @ B[N,K]^T) where out_mask[M,N] > 0, else 0 - the ReLU backward of the
 * preceding layer folded into the GEMM's epilogue (out_mask = that layer's output, same layout as C). */
int rl_gemm_ts_masked_f32(const float* A, const float* B, const float* B_lo, const float* out_mask, float* C, int64_t M,
                          int64_t N, int64_t K, void* workspace, void* stream);
/* lo[i] = src[i] - trunc_tf32(src[i]);  dst = src^T ([cols,rows]) and dst_lo = dst - trunc_tf32(dst) */
int rl_split_lo_f32(const float* src, float* lo, int64_t n, void* stream);
int rl_transpose_split_f32(const float* src, float* dst, float* dst_lo, int64_t rows, int64_t cols, void* stream);

/* ------------------------------------------------------------------ conv layers as tcgen05 implicit GEMM
 * Same contracts as rl_conv1_u8_forward (uint8 frames, optional row gather, *1/255, k8 s4, 4->16)
 * and as Conv2d(16->32, k4, s2, p1)(+bias)(+ReLU) on fp32 NCHW activations
 * (rlpyt/models/conv2d.py:36-44, rlpyt/models/pg/atari_ff_model.py:31-35, :50-53), computed on the
 * tensor cores: tcgen05.mma kind::tf32 with the fp32-accurate hi/lo operand split, im2col gathered on
 * the fly into the swizzled shared-memory operand tiles, accumulators in TMEM. */
int rl_conv1_u8_forward_tc(const uint8_t* obs, const int64_t* rows, const float* weight, const float* bias,
                           float* out, int64_t N, int C, int H, int W, int relu, void* stream);
int rl_conv2_forward_tc(const float* x, const float* weight, const float* bias, float* out, int64_t N,
                        int C, int IH, int IW, int relu, void* stream);
/* Input gradient of the same layer (ConvolutionBackward wrt input): grad_out_masked [N,32,OH,OW] is
 * grad_out * (out > 0); grad_x [N,16,IH,IW].  Four parity-class GEMMs (K = 32 channels x 2x2 taps) on
 * the same tcgen05 kernel.  scratch: rl_conv2_dgrad_tc_scratch_bytes() bytes, 16B aligned. */
int64_t rl_conv2_dgrad_tc_scratch_bytes(void);
int rl_conv2_dgrad_tc(const float* grad_out_masked, const float* weight, float* grad_x, int64_t N, int C,
                      int IH, int IW, void* scratch, void* stream);
/* Weight and bias gradients of both layers (ConvolutionBackward wrt weight/bias, with the ReLU mask of
 * models/conv2d.py:41 folded in): one tcgen05 GEMM over all output positions,
 *     grad_weight[oc][tap] = scale * sum_m im2col(x)[m][tap] * g[m][oc],   g = grad_out * (out > 0),
 * taps on the MMA M axis (2 x 128), positions on the K axis, per-CTA partial sums promoted to fp32
 * registers every 4 k-blocks and reduced in CTA order (deterministic).  out may be NULL when grad_out
 * is already masked; grad_bias may be NULL.  grad_weight [16,4,8,8] / [32,16,4,4], grad_bias [16] / [32].
 * scratch: rl_conv_wgrad_tc_scratch_bytes() bytes, 16B aligned.  Requires N*OH*OW < 2^31. */
int64_t rl_conv_wgrad_tc_scratch_bytes(void);
int rl_conv1_u8_wgrad_tc(const uint8_t* obs, const int64_t* rows, const float* out, const float* grad_out,
                         float* grad_weight, float* grad_bias, int64_t N, int C, int H, int W, void* scratch,
                         void* stream);
int rl_conv2_wgrad_tc(const float* x, const float* out, const float* grad_out, float* grad_weight,
                      float* grad_bias, int64_t N, int C, int IH, int IW, void* scratch, void* stream);

/* ------------------------------------------------------------------ first conv layer on the INTEGER tensor cores
 * Same contracts as rl_conv1_u8_forward / rl_conv1_u8_wgrad_tc (rlpyt/models/pg/atari_ff_model.py:50-53,
 * rlpyt/models/conv2d.py:36-44; row gather of rlpyt/algos/pg/ppo.py:99-100), computed with tcgen05.mma kind::i8:
 * the uint8 frames are exact MMA operands (space-to-depth "cell" rows, no im2col expansion, frames streamed by
 * cp.async.bulk), the fp32 operand (filter bank / output gradient) is split into four base-128 digits against a
 * power-of-two scale per output channel, the int32 accumulators are exact and recombined in fp32 (forward) / fp64
 * (gradient).  Error vs exact arithmetic <= 2^-28 of the channel's largest |weight| (|gradient|) per term.
 * Requirements: C == 4, H % 4 == 0, W % 4 == 0, W <= 128, frames 16-byte aligned;
 * rl_conv1_u8_i8_supported() tells whether a geometry fits (else use the *_tc entry points).
 * wgrad: out may be NULL (grad_out already masked), grad_bias may be NULL; N <= 256 * SM count;
 * scratch: rl_conv1_u8_wgrad_i8_scratch_bytes() bytes, 16B aligned.  Deterministic. */
int rl_conv1_u8_i8_supported(int C, int H, int W);
int rl_conv1_u8_forward_i8(const uint8_t* obs, const int64_t* rows, const float* weight, const float* bias,
                           float* out, int64_t N, int C, int H, int W, int relu, void* stream);
/* The sampler's variant of the forward: the frames are read straight out of PAGE-LOCKED, device-mapped HOST memory
 * (the step buffer the env workers write, rlpyt/samplers/parallel/gpu/collectors.py:38) by the kernel's bulk copies and
 * are ALSO written to obs_copy ([N,C,H,W] u8 in HBM = observation[t] of the [T,B] batch) by one bulk store per frame, so
 * the observations cross PCIe once and no host->device copy stands in front of agent.step
 * (rlpyt/samplers/parallel/gpu/action_server.py:46-55 copies, then steps). */
int rl_conv1_u8_forward_i8_stream(const uint8_t* obs_host_mapped, const float* weight, const float* bias, float* out,
                                  uint8_t* obs_copy, int64_t N, int C, int H, int W, int relu, void* stream);
int64_t rl_conv1_u8_wgrad_i8_scratch_bytes(void);
int rl_conv1_u8_wgrad_i8(const uint8_t* obs, const int64_t* rows, const float* out, const float* grad_out,
                         float* grad_weight, float* grad_bias, int64_t N, int C, int H, int W, void* scratch,
                         void* stream);
/* The same with the per-channel bound max |grad_out[:, oc]| (16 floats in device memory, an upper bound is enough)
 * supplied by the caller - rl_conv2_dgrad_s2d_absmax produces it in its epilogue - instead of a separate pass over
 * grad_out. */
int rl_conv1_u8_wgrad_i8_scaled(const uint8_t* obs, const int64_t* rows, const float* out, const float* grad_out,
                                const float* chan_absmax, float* grad_weight, float* grad_bias, int64_t N, int C, int H, int W,
                                void* scratch, void* stream);

/* ------------------------------------------------------------------ policy + value heads in training
 * rlpyt/models/pg/atari_ff_model.py:56-61: pi = softmax(h W_pi^T + b_pi), v = h w_v^T + b_v for h [N,F], W_pi [A,F],
 * w_v [F] (A <= 32) - forward in one kernel, backward (softmax backward, both input gradients summed, both weight and
 * bias gradients; grad_prob / grad_value nullable = zero) in one kernel + a fixed-order fp64 reduction over CTAs
 * (F <= 1024; scratch: rl_pg_heads_backward_scratch_bytes(N, F, A) bytes).  Replaces the two torch.nn.Linear + softmax
 * (five SIMT GEMM / GEMV launches, split-K and bias reductions: the heads are 6 and 1 columns wide). */
int rl_pg_heads_forward_f32(const float* h, const float* w_pi, const float* b_pi, const float* w_v, const float* b_v,
                            float* prob, float* value, int64_t N, int F, int A, void* stream);
int64_t rl_pg_heads_backward_scratch_bytes(int64_t N, int F, int A);
int rl_pg_heads_backward_f32(const float* h, const float* prob, const float* grad_prob, const float* grad_value,
                             const float* w_pi, const float* w_v, float* grad_h, float* grad_w_pi, float* grad_b_pi,
                             float* grad_w_v, float* grad_b_v, int64_t N, int F, int A, void* scratch, void* stream);

/* ------------------------------------------------------------------ second conv layer without im2col expansion
 * Same contracts as rl_conv2_forward_tc / rl_conv2_dgrad_tc (Conv2d(16->32, k4, s2, p1), fp32 NCHW;
 * rlpyt/models/conv2d.py:36-44, rlpyt/models/pg/atari_ff_model.py:31-35): space-to-depth "cell" rows, the four 2x2
 * taps as row-shifted tcgen05 descriptors over one shared-memory tile, images streamed by cp.async.bulk, the input
 * gradient assembled in shared memory and written by one bulk store; fp32-accurate 3-term TF32.
 * Requirements: C == 16, OW <= 14, x / gradients 16-byte aligned; rl_conv2_s2d_supported() tells
 * whether a geometry fits (else use the *_tc entry points).  No scratch. */
int rl_conv2_s2d_supported(int C, int IH, int IW);
int rl_conv2_forward_s2d(const float* x, const float* weight, const float* bias, float* out, int64_t N, int C,
                         int IH, int IW, int relu, void* stream);
int rl_conv2_dgrad_s2d(const float* grad_out_masked, const float* weight, float* grad_x, int64_t N, int C, int IH,
                       int IW, void* stream);
/* rl_conv2_dgrad_s2d that also writes chan_absmax[16] = max |grad_x[:, c]| per input channel (device floats), for the
 * first layer's rl_conv1_u8_wgrad_i8_scaled. */
int rl_conv2_dgrad_s2d_absmax(const float* grad_out_masked, const float* weight, float* grad_x, float* chan_absmax, int64_t N,
                              int C, int IH, int IW, void* stream);
/* Weight / bias gradient in the same cell space (grad_out already ReLU-masked): both operands MN-major with
 * K = cells, M = 64 accumulators per tap promoted to fp32 registers per image, per-CTA partials reduced in a
 * fixed order.  grad_weight [32,16,4,4], grad_bias [32] (nullable); scratch: rl_conv2_wgrad_s2d_scratch_bytes(). */
int64_t rl_conv2_wgrad_s2d_scratch_bytes(void);
int rl_conv2_wgrad_s2d(const float* x, const float* grad_out_masked, float* grad_weight, float* grad_bias, int64_t N,
                       int C, int IH, int IW, void* scratch, void* stream);

/* ------------------------------------------------------------------ Categorical.sample (agent.step)
 * rlpyt/distributions/categorical.py:25-30 (torch.multinomial over the trailing dim) as an inverse-CDF draw:
 * action[i] = #{k < A-1 : u_i >= cumsum_fp32(prob[i])[k]} (oracle/pg_loss.py:sample_categorical).  prob [N,A] f32.
 * uniform [N] f32 in [0,1): injected draws (bit-exact parity with the oracle) - or NULL, then u_i is the first
 * Philox4x32-10 word of counter (i, rng_state[1]) under key rng_state[0], top 24 bits; rng_state is int64[2] =
 * {seed, call counter} in DEVICE memory and the kernel increments the counter, so a captured launch draws new
 * numbers at every graph replay.  uniform_out [N] (nullable) receives the uniforms used.  One 256-thread block. */
int rl_categorical_sample_f32(const float* prob, const float* uniform, int64_t* rng_state, int64_t* action,
                              float* uniform_out, int64_t N, int A, void* stream);
/* The policy head of agent.step in one launch (rlpyt/models/pg/atari_ff_model.py:56-58, rlpyt/agents/pg/
 * categorical.py:37-39): prob = softmax(h w_pi^T + b_pi) [B,A], value = h w_v + b_v [B], action ~ prob (rule and
 * Philox stream of rl_categorical_sample_f32; row i uses counter (i, rng_state[1])).  h [B,F], w_pi [A,F], b_pi [A],
 * w_v [F], b_v [1] f32, A <= 32.  rng_state: int64[3] in device memory = {seed, call counter, 0 (block ticket)} -
 * the last block increments the counter - or NULL with injected `uniform` [B]. */
int rl_pg_head_sample_f32(const float* h, const float* w_pi, const float* b_pi, const float* w_v, const float* b_v,
                          const float* uniform, int64_t* rng_state, float* prob, float* value, int64_t* action,
                          int64_t B, int F, int A, void* stream);

/* ------------------------------------------------------------------ DQN loss
 * rlpyt/algos/dqn/dqn.py:230-263 `DQN.loss` after the two network forwards: Q(s,a) selection, (double-)DQN
 * target, y = return_ + (1 - done_n) * disc_n * target_q, Huber loss with threshold delta_clip (delta_clip < 0:
 * plain 0.5*delta^2), optional importance weights, loss = mean, td_abs_errors = clamp(|delta|, 0, delta_clip)
 * (bit-identical per-sample fp32 arithmetic - these are the new priorities), and grad_qs = dloss/dqs [N,A].
 * qs/target_qs/next_qs [N,A] f32 (next_qs NULL: no double-DQN), action [N] i64, return_ [N] f32, done_n [N] u8,
 * is_weights [N] f32 or NULL, disc_n = (float)(discount ** n_step_return); out_scalars[2] = {loss, N};
 * grad_qs may be NULL.  scratch: rl_dqn_loss_scratch_bytes(N) bytes, 8B aligned.  A <= 64. */
int64_t rl_dqn_loss_scratch_bytes(int64_t N);
int rl_dqn_loss_f32(const float* qs, const float* target_qs, const float* next_qs, const int64_t* action,
                    const float* return_, const uint8_t* done_n, const float* is_weights, int64_t N, int A,
                    float disc_n, float delta_clip, float* out_scalars, float* td_abs_errors, float* grad_qs,
                    void* scratch, void* stream);

/* ------------------------------------------------------------------ layer helpers (HBM-bound)
 * rl_relu_backward_f32: dst[i] = out[i] > 0 ? grad[i] : 0 - the backward of the torch.nn.ReLU that
 *   follows every conv / linear layer (rlpyt/models/conv2d.py:41, rlpyt/models/mlp.py:33) in one pass.
 * rl_transpose_f32: dst[cols,rows] = src[rows,cols]^T - operand re-layout for the weight-gradient GEMM
 *   (rl_gemm_tf32x3_f32 wants the reduction axis contiguous in both operands). */
int rl_relu_backward_f32(const float* grad, const float* out, float* dst, int64_t n, void* stream);
int rl_transpose_f32(const float* src, float* dst, int64_t rows, int64_t cols, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RLPYT_B200_H */
