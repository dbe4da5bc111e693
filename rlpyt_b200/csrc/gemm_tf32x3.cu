// fp32-accurate GEMM on the 5th-gen tensor cores: C[M,N] = A[M,K] * B[N,K]^T (+ bias[N]) (+ ReLU),
// A and B row-major with K contiguous ("TN"), i.e. exactly torch.nn.Linear: y = x W^T + b.
//
// It is the one dense-GEMM-shaped piece of the path: the AtariFf fully connected layer
// (rlpyt/models/pg/atari_ff_model.py:24-35 -> rlpyt/models/mlp.py:30-36, 3200 -> 512 + ReLU;
// M = B = 256 in agent.step, M = 8192 per PPO minibatch, forward + dgrad + wgrad).
//
// Why 3xTF32: the reference computes in fp32 and parity is 1e-5, so a plain TF32/BF16 MMA (10 / 7
// mantissa bits) is not admissible, and cuBLAS's fp32 path runs on the SIMT pipe (~0.5 ms for
// 8192x512x3200).  Each operand is split exactly as x = hi + lo with hi = x truncated to TF32 and
// lo = x - hi (13 significant bits), and  A*B ~= Ahi*Bhi + Ahi*Blo + Alo*Bhi  is accumulated in fp32
// in TMEM: relative error ~2^-20 per product, at one third of the TF32 tensor rate.
//
// The tensor core's fp32 accumulator truncates on every accumulate step (measured: one TMEM
// accumulation over K=3200 = 1200 MMA steps drifts by ~4e-6 of sum|a||b|, i.e. 1.5e-4 on O(1)
// outputs), so the K loop is PROMOTED: every kChunk k-blocks (K=128, 48 MMA steps) the partial
// tile is drained from TMEM and added to fp32 registers with round-to-nearest, using two TMEM
// accumulators so the drain overlaps the next chunk's MMAs.
//
// Structure (one CTA per 128x128 output tile, 384 threads, warp-specialised):
//   warp 0   TMA producer: cp.async.bulk.tensor 2D loads of the raw fp32 A/B k-blocks (128 rows x
//            32 floats = one 128-byte swizzle atom per row, SWIZZLE_128B) into a 3-stage ring;
//   warps 4-7 "split" warps: rewrite each landed tile in place to hi and write lo to a twin tile
//            (elementwise, so the swizzled layout is preserved), fence.proxy.async, signal;
//   warp 1   MMA issuer: one elected lane issues 4 k-slices x 3 tcgen05.mma.kind::tf32
//            (M=128, N=128, K=8) per stage, tcgen05.commit frees the stage / publishes the tile;
//   warp 2   allocates / frees 256 TMEM columns (2 x 128 lanes x 128 fp32 accumulators);
//   warps 8-11 drain + epilogue: per chunk tcgen05.ld 32x32b.x32 -> += 128 fp32 registers; at the end
//            + bias -> ReLU -> 16-byte global stores.
// Shared memory: 3 stages x (A, A_lo, B, B_lo) x 16 KiB = 192 KiB.
#include "tc_common.cuh"

namespace rl {
namespace gemm {

using namespace tc;

constexpr int BM = 128, BN = 128, BK = 32;       // BK fp32 = 128 bytes = one swizzle-128B row
constexpr int kStages = 3;
constexpr int kTileBytes = BM * BK * 4;          // 16 KiB (BM == BN)
constexpr int kStageBytes = 4 * kTileBytes;      // A_hi, A_lo, B_hi, B_lo
constexpr int kThreads = 384;
constexpr int kSplitThreads = 128;               // warps 4..7
constexpr int kDrainThreads = 128;               // warps 8..11
constexpr int kChunk = 4;                        // k-blocks accumulated in TMEM before promotion
constexpr int kTmemCols = 256;                   // two 128-column accumulators
constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
static_assert(3 * kStages + 4 + 1 <= 32, "barrier block");

constexpr uint32_t kIdesc = tc::make_idesc_tf32(BM, BN);

__device__ __forceinline__ void split4(float4& v, float4& lo) {
    float4 hi;
    tc::split_tf32(v.x, hi.x, lo.x);
    tc::split_tf32(v.y, hi.y, lo.y);
    tc::split_tf32(v.z, hi.z, lo.z);
    tc::split_tf32(v.w, hi.w, lo.w);
    v = hi;
}

// kRawHi: leave the TMA-loaded fp32 tile in place as the "hi" operand (the tensor core ignores the low
// 13 mantissa bits of a TF32 operand, i.e. it sees trunc(x) - exactly the hi term split_tf32 would
// write) and only write the lo tile: one third less split-warp shared-memory traffic.
template <bool kRawHi>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                   float* __restrict__ C, const float* __restrict__ bias, int M, int N, int K, int relu,
                   int kb_per_split) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
    uint64_t* full_tma = bars;                 // [kStages] TMA landed
    uint64_t* full_mma = bars + kStages;       // [kStages] hi/lo written
    uint64_t* empty = bars + 2 * kStages;      // [kStages] MMAs of the stage retired
    uint64_t* tmem_full = bars + 3 * kStages;  // [2] chunk accumulated (MMA -> drain)
    uint64_t* tmem_empty = tmem_full + 2;      // [2] accumulator drained (drain -> MMA)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    // split-K (small M): blockIdx.z owns k-blocks [kb0, kb0 + num_kb) and writes its partial tile to
    // slice z of the workspace C[z][M][N]; splitk_reduce_kernel sums the slices in order.
    const int total_kb = (K + BK - 1) / BK;
    const int kb0 = blockIdx.z * kb_per_split;
    const int num_kb = min(kb_per_split, total_kb - kb0);
    C += static_cast<int64_t>(blockIdx.z) * M * N;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full_tma[s], 1);
            mbar_init(&full_mma[s], kSplitThreads / 32);   // one elected arrive per split warp
            mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tmem_full[b], 1);
            mbar_init(&tmem_empty[b], kDrainThreads / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = uniform_u32(*tmem_slot);

    if (warp == 0) {
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a)) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_b)) : "memory");
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % kStages;
                const uint32_t ph = (kb / kStages) & 1;
                mbar_wait(&empty[s], ph ^ 1);
                uint8_t* st = smem + s * kStageBytes;
                mbar_expect_tx(&full_tma[s], 2 * kTileBytes);
                tma_load_2d(st, &map_a, (kb0 + kb) * BK, m0, &full_tma[s]);
                tma_load_2d(st + 2 * kTileBytes, &map_b, (kb0 + kb) * BK, n0, &full_tma[s]);
            }
        }
    } else if (warp == 1) {
        const int num_chunks = (num_kb + kChunk - 1) / kChunk;
        for (int c = 0; c < num_chunks; ++c) {
            const int buf = c & 1;
            mbar_wait(&tmem_empty[buf], ((c >> 1) & 1) ^ 1);   // drained (passes at once the first two times)
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tmem_acc = tmem_base + static_cast<uint32_t>(buf * BN);
            const int kb_end = min(num_kb, (c + 1) * kChunk);
            for (int kb = c * kChunk; kb < kb_end; ++kb) {
                const int s = kb % kStages;
                const uint32_t ph = (kb / kStages) & 1;
                mbar_wait(&full_mma[s], ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (elect_one()) {
                    uint8_t* st = smem + s * kStageBytes;
                    const uint64_t a_hi = make_desc(st), a_lo = make_desc(st + kTileBytes);
                    const uint64_t b_hi = make_desc(st + 2 * kTileBytes), b_lo = make_desc(st + 3 * kTileBytes);
#pragma unroll
                    for (int k = 0; k < BK / 8; ++k) {          // UMMA_K = 8 tf32 = 32 bytes = 2 x 16 B
                        const uint64_t adv = static_cast<uint64_t>(k * 2);
                        umma_tf32(tmem_acc, a_hi + adv, b_hi + adv, kIdesc, (kb > c * kChunk || k > 0) ? 1u : 0u);
                        umma_tf32(tmem_acc, a_hi + adv, b_lo + adv, kIdesc, 1u);
                        umma_tf32(tmem_acc, a_lo + adv, b_hi + adv, kIdesc, 1u);
                    }
                    umma_commit(&empty[s]);
                    if (kb == kb_end - 1) umma_commit(&tmem_full[buf]);
                }
                __syncwarp();
            }
        }
    } else if (warp >= 4 && warp < 8) {
        const int t = threadIdx.x - 128;
        for (int kb = 0; kb < num_kb; ++kb) {
            const int s = kb % kStages;
            const uint32_t ph = (kb / kStages) & 1;
            mbar_wait(&full_tma[s], ph);
            const uint32_t st = smem_u32(smem + s * kStageBytes) + static_cast<uint32_t>(t) * 16u;
#pragma unroll
            for (int i = 0; i < kTileBytes / 16 / kSplitThreads; ++i) {   // 8 float4 per thread per tile
                const uint32_t a = st + static_cast<uint32_t>(i * kSplitThreads) * 16u;
                float4 va = lds128(a), la, vb = lds128(a + 2 * kTileBytes), lb;
                split4(va, la);
                split4(vb, lb);
                if (!kRawHi) { sts128(a, va); sts128(a + 2 * kTileBytes, vb); }
                sts128(a + kTileBytes, la); sts128(a + 3 * kTileBytes, lb);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> async (tensor) proxy
            __syncwarp();
            if (lane == 0) mbar_arrive(&full_mma[s]);
        }
    } else if (warp >= 8) {
        // ---- drain: promote each chunk's TMEM partial into fp32 registers (round-to-nearest adds)
        const int q = warp - 8;                                  // TMEM lane quarter of this warp
        const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
        float acc[BN];
#pragma unroll
        for (int j = 0; j < BN; ++j) acc[j] = 0.0f;
        const int num_chunks = (num_kb + kChunk - 1) / kChunk;
        for (int c = 0; c < num_chunks; ++c) {
            const int buf = c & 1;
            mbar_wait(&tmem_full[buf], (c >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int g = 0; g < BN / 32; ++g) {
                uint32_t r[32];
                tmem_ld32(tmem_base + lane_base + static_cast<uint32_t>(buf * BN + g * 32), r);
#pragma unroll
                for (int j = 0; j < 32; ++j) acc[g * 32 + j] += __uint_as_float(r[j]);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[buf]);
        }
        // ---- epilogue: (+bias, ReLU) -> global
        const int row = m0 + q * 32 + lane;
        const bool vec_ok = (N % 4 == 0);
        if (row < M) {
            float* out = C + static_cast<int64_t>(row) * N + n0;
#pragma unroll
            for (int j = 0; j < BN; j += 4) {
                const int n = n0 + j;
                float v[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float x = acc[j + e];
                    if (bias != nullptr && n + e < N) x += bias[n + e];
                    v[e] = relu ? fmaxf(x, 0.0f) : x;
                }
                if (vec_ok && n + 3 < N) {
                    *reinterpret_cast<float4*>(out + j) = make_float4(v[0], v[1], v[2], v[3]);
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (n + e < N) out[j + e] = v[e];
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols));
    }
}

// out[m][n] = sum_z ws[z][m][n] (+ bias[n]) (+ ReLU), z in ascending order (deterministic)
__global__ void splitk_reduce_kernel(const float* __restrict__ ws, int splits, const float* __restrict__ bias,
                                     float* __restrict__ out, int64_t MN, int N, int relu) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= MN) return;
    float s = 0.0f;
    for (int z = 0; z < splits; ++z) s += ws[static_cast<int64_t>(z) * MN + i];
    if (bias != nullptr) s += bias[i % N];
    out[i] = relu ? fmaxf(s, 0.0f) : s;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// rows x K fp32 row-major matrix, box = 128 rows x 32 floats, 128-byte swizzle, OOB -> 0
static int make_map(CUtensorMap* map, const float* base, int64_t rows, int64_t K) {
    EncodeTiledFn fn = encode_fn();
    RL_REQUIRE(fn != nullptr, RL_EINVAL, "gemm_tf32x3: cuTensorMapEncodeTiled unavailable");
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(rows)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(K) * 4};
    cuuint32_t box[2] = {BK, BM};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    RL_REQUIRE(r == CUDA_SUCCESS, RL_EINVAL, "gemm_tf32x3: cuTensorMapEncodeTiled failed (%d)", static_cast<int>(r));
    return RL_OK;
}

}  // namespace gemm
}  // namespace rl

extern "C" {

int64_t rl_gemm_tf32x3_workspace_bytes(int64_t M, int64_t N, int64_t K) {
    const int64_t tiles = ((M + rl::gemm::BM - 1) / rl::gemm::BM) * ((N + rl::gemm::BN - 1) / rl::gemm::BN);
    const int64_t kb = (K + rl::gemm::BK - 1) / rl::gemm::BK;
    int sms = rl::sm_count();
    if (sms <= 0) sms = 148;
    int64_t splits = 1;
    if (tiles * 2 <= sms) {                       // fewer than half a wave of tiles: split K
        splits = sms / tiles;
        if (splits > kb / 4) splits = kb / 4;     // keep >= 4 k-blocks (one promotion chunk) per split
        if (splits > 16) splits = 16;
        if (splits < 1) splits = 1;
    }
    return splits > 1 ? splits * M * N * static_cast<int64_t>(sizeof(float)) : 0;
}

int rl_gemm_tf32x3_f32(const float* A, const float* B, const float* bias, float* C, int64_t M, int64_t N,
                       int64_t K, int relu, void* workspace, void* stream) {
    RL_REQUIRE(A && B && C, RL_EINVAL, "rl_gemm_tf32x3_f32: null pointer");
    RL_REQUIRE(M >= 1 && N >= 1 && K >= 1, RL_EINVAL, "rl_gemm_tf32x3_f32: M=%lld N=%lld K=%lld", (long long)M,
               (long long)N, (long long)K);
    RL_REQUIRE(K % 4 == 0 && rl::aligned(A, 16) && rl::aligned(B, 16) && rl::aligned(C, 16), RL_EALIGN,
               "rl_gemm_tf32x3_f32: K %% 4 == 0 and 16B-aligned A, B, C required (TMA row pitch)");
    CUtensorMap ma, mb;
    int rc = rl::gemm::make_map(&ma, A, M, K);
    if (rc != RL_OK) return rc;
    rc = rl::gemm::make_map(&mb, B, N, K);
    if (rc != RL_OK) return rc;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(rl::gemm::gemm_tf32x3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             rl::gemm::kSmemBytes);
        cudaFuncSetAttribute(rl::gemm::gemm_tf32x3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             rl::gemm::kSmemBytes);
        attr_set = true;
    }
    static int raw_hi = -1;
    if (raw_hi < 0) {
        const char* e = getenv("RLPYT_B200_GEMM_RAW_HI");
        raw_hi = e ? atoi(e) : 1;
    }
    auto kern = raw_hi ? rl::gemm::gemm_tf32x3_kernel<true> : rl::gemm::gemm_tf32x3_kernel<false>;
    const int64_t ws_bytes = rl_gemm_tf32x3_workspace_bytes(M, N, K);
    const int total_kb = static_cast<int>((K + rl::gemm::BK - 1) / rl::gemm::BK);
    int splits = static_cast<int>(ws_bytes / (M * N * static_cast<int64_t>(sizeof(float))));
    if (splits < 1 || workspace == nullptr) splits = 1;
    int kb_per_split = (total_kb + splits - 1) / splits;
    splits = (total_kb + kb_per_split - 1) / kb_per_split;   // no empty split
    cudaStream_t st = rl::as_stream(stream);
    dim3 grid(static_cast<unsigned>((N + rl::gemm::BN - 1) / rl::gemm::BN),
              static_cast<unsigned>((M + rl::gemm::BM - 1) / rl::gemm::BM), static_cast<unsigned>(splits));
    if (splits == 1) {
        kern<<<grid, rl::gemm::kThreads, rl::gemm::kSmemBytes, st>>>(
            ma, mb, C, bias, static_cast<int>(M), static_cast<int>(N), static_cast<int>(K), relu, total_kb);
        return rl::check_launch("gemm_tf32x3_kernel");
    }
    RL_REQUIRE(rl::aligned(workspace, 16), RL_EALIGN, "rl_gemm_tf32x3_f32: workspace must be 16B aligned");
    float* ws = static_cast<float*>(workspace);
    kern<<<grid, rl::gemm::kThreads, rl::gemm::kSmemBytes, st>>>(
        ma, mb, ws, nullptr, static_cast<int>(M), static_cast<int>(N), static_cast<int>(K), 0, kb_per_split);
    rc = rl::check_launch("gemm_tf32x3_kernel");
    if (rc != RL_OK) return rc;
    const int64_t MN = M * N;
    rl::gemm::splitk_reduce_kernel<<<static_cast<unsigned>((MN + 255) / 256), 256, 0, st>>>(
        ws, splits, bias, C, MN, static_cast<int>(N), relu);
    return rl::check_launch("splitk_reduce_kernel");
}

}  // extern "C"
