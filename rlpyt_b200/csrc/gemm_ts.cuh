// fp32-accurate GEMM on the 5th-gen tensor cores, second version: C[M,N] = A[M,K] * B[N,K]^T (+ bias[N]) (+ ReLU)
// with the A operand fed from TENSOR MEMORY ("TS" form of tcgen05.mma) and a persistent tile loop.
//
// It replaces gemm_tf32x3.cu on the fully connected layer of the AtariFf network
// (rlpyt/models/pg/atari_ff_model.py:24-35 -> rlpyt/models/mlp.py:30-36, 3200 -> 512 + ReLU; M = 8192 per PPO
// minibatch: forward, input gradient, weight gradient; M = 256 in agent.step).
//
// Why.  The first kernel (both operands in shared memory) is bound by shared-memory bandwidth, not by the tensor
// pipe: per 32-wide k-block of a 128x128 tile it moves 192 KB through shared memory (TMA writes 32, split warps
// read 32 / write 32, the 12 MMAs read 96) = 1536 cycles at 128 B/clk against 768 cycles of MMA - measured 161 us
// at 8192x512x3200 = exactly that model.  Here
//   * A (the large operand: activations) is loaded raw by TMA, read ONCE by the split warps, and written to TMEM
//     as two 32-column operands (raw = "hi": the tensor core ignores the 13 low mantissa bits, and
//     lo = x - trunc(x)); the MMAs read it from TMEM, not from shared memory;
//   * B (the small operand: weights / transposed gradient) arrives already split: B and B_lo = B - trunc(B) are
//     two tensors in HBM (rl_split_lo_f32 / rl_transpose_split_f32), so nobody rewrites B in shared memory;
//   * per k-block: A 16 (TMA) + 16 (split read), B 32 (TMA), MMA reads of B 48  = 112 KB = 896 cycles.
// The A source may also be given "M-major" (a [K,M] row-major matrix, i.e. A^T as it lies in memory): the split
// warps read a [32 k][128 m] box column-wise (conflict-free LDS.32) - the weight gradient needs no transpose of
// the activations.  C may be written transposed (C^T [N,M]) for the same reason.
//
// K loop promotion as before (the TMEM fp32 accumulator truncates): every kChunk k-blocks the partial tile is
// drained into fp32 registers with round-to-nearest adds, two accumulators so the drain overlaps the MMAs.
//
// One persistent CTA per SM walks units (tile, k-split) round-robin; 384 threads, warp-specialised:
//   warp 0     TMA producer: A ring (4 x 16 KB), B ring (4 x (hi 16 KB + lo 16 KB))
//   warp 1     MMA issuer: per k-block 4 k-slices x 3 tcgen05.mma.kind::tf32 (A from TMEM, B from smem)
//   warp 2     TMEM allocation (512 columns: 2 x 128 accumulator + 3 x 64 A stages)
//   warps 4-7  split warps: smem A tile -> registers -> tcgen05.st (raw, lo) into the TMEM A ring
//   warps 8-11 drain + epilogue
#pragma once
#include "tc_common.cuh"

namespace rl {
namespace gts {

using namespace tc;

constexpr int BM = 128, BN = 128, BK = 32;
constexpr int kAS = 4, kBS = 4, kTS = 3;
constexpr int kATile = BM * BK * 4;              // 16 KiB
constexpr int kBHalf = BN * BK * 4;              // 16 KiB
constexpr int kBTile = 2 * kBHalf;               // hi rows then lo rows
constexpr int kThreads = 384;
constexpr int kChunk = 4;
constexpr int kTCol0 = 2 * BN;                   // A stages start after the two accumulators
constexpr int kTStageCols = 2 * BK;              // raw 32 + lo 32
constexpr int kTmemCols = 512;
static_assert(kTCol0 + kTS * kTStageCols <= kTmemCols, "TMEM budget");
constexpr int kBarBytes = 512;
constexpr int kSmemBytes = kAS * kATile + kBS * kBTile + 1024 /*align slack*/ + kBarBytes;

constexpr uint32_t kIdesc = tc::make_idesc_tf32(BM, BN);

struct Params {
    float* C;              // [M,N] (or [N,M] when c_trans); with splits > 1: workspace [splits][...] in the same layout
    const float* bias;     // [N] or nullptr (applied here only when splits == 1)
    const float* mask;     // [M,N] or nullptr: result element kept where mask > 0, else 0 (the ReLU backward of the layer
                           // that produced this GEMM's gradient input, fused into the epilogue; not with c_trans)
    int M, N, K;
    int relu;
    int c_trans;
    int splits, kb_per_split;
    int tiles_m, tiles_n;
};

__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
          "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
          "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}

// kAMajorM: A is given as a [K,M] row-major matrix (box = 32 k-rows x 128 m floats, no swizzle);
// otherwise [M,K] row-major (box = 128 rows x 32 floats, SWIZZLE_128B).
template <bool kAMajorM>
__global__ void __launch_bounds__(kThreads, 1)
gemm_ts_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const __grid_constant__ CUtensorMap map_blo, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + kAS * kATile;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kAS * kATile + kBS * kBTile);
    uint64_t* a_full = bars;                 // [kAS] TMA -> split
    uint64_t* a_empty = a_full + kAS;        // [kAS] split -> TMA
    uint64_t* b_full = a_empty + kAS;        // [kBS] TMA -> MMA
    uint64_t* b_empty = b_full + kBS;        // [kBS] MMA -> TMA
    uint64_t* t_full = b_empty + kBS;        // [kTS] split -> MMA (A in TMEM)
    uint64_t* t_empty = t_full + kTS;        // [kTS] MMA -> split
    uint64_t* acc_full = t_empty + kTS;      // [2]
    uint64_t* acc_empty = acc_full + 2;      // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    const int total_kb = (p.K + BK - 1) / BK;
    const int units = p.tiles_m * p.tiles_n * p.splits;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kAS; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 4); }
        for (int s = 0; s < kBS; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
        for (int s = 0; s < kTS; ++s) { mbar_init(&t_full[s], 4); mbar_init(&t_empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 4); }
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
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a)) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_b)) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_blo)) : "memory");
            uint32_t it = 0;
            for (int u = blockIdx.x; u < units; u += gridDim.x) {
                const int z = u % p.splits, tile = u / p.splits;
                const int m0 = (tile / p.tiles_n) * BM, n0 = (tile % p.tiles_n) * BN;
                const int kb0 = z * p.kb_per_split, nkb = min(p.kb_per_split, total_kb - kb0);
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const uint32_t as = it % kAS, bs = it % kBS;
                    const int k0 = (kb0 + kb) * BK;
                    mbar_wait(&a_empty[as], ((it / kAS) & 1) ^ 1);
                    mbar_expect_tx(&a_full[as], kATile);
                    if (kAMajorM) tma_load_2d(smem_a + as * kATile, &map_a, m0, k0, &a_full[as]);
                    else tma_load_2d(smem_a + as * kATile, &map_a, k0, m0, &a_full[as]);
                    mbar_wait(&b_empty[bs], ((it / kBS) & 1) ^ 1);
                    mbar_expect_tx(&b_full[bs], kBTile);
                    tma_load_2d(smem_b + bs * kBTile, &map_b, k0, n0, &b_full[bs]);
                    tma_load_2d(smem_b + bs * kBTile + kBHalf, &map_blo, k0, n0, &b_full[bs]);
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        uint32_t it = 0, ch = 0;
        for (int u = blockIdx.x; u < units; u += gridDim.x) {
            const int z = u % p.splits;
            const int kb0 = z * p.kb_per_split, nkb = min(p.kb_per_split, total_kb - kb0);
            for (int c0 = 0; c0 < nkb; c0 += kChunk, ++ch) {
                const uint32_t buf = ch & 1;
                mbar_wait(&acc_empty[buf], ((ch >> 1) & 1) ^ 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tmem_acc = tmem_base + buf * BN;
                const int c1 = min(nkb, c0 + kChunk);
                for (int kb = c0; kb < c1; ++kb, ++it) {
                    const uint32_t ts = it % kTS, bs = it % kBS;
                    mbar_wait(&t_full[ts], (it / kTS) & 1);
                    mbar_wait(&b_full[bs], (it / kBS) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (elect_one()) {
                        const uint32_t a_raw = tmem_base + kTCol0 + ts * kTStageCols, a_lo = a_raw + BK;
                        const uint64_t b_hi = make_desc(smem_b + bs * kBTile), b_lo = make_desc(smem_b + bs * kBTile + kBHalf);
#pragma unroll
                        for (int k = 0; k < BK / 8; ++k) {
                            const uint64_t adv = static_cast<uint64_t>(k * 2);
                            umma_tf32_ts(tmem_acc, a_raw + k * 8, b_hi + adv, kIdesc, (kb > c0 || k > 0) ? 1u : 0u);
                            umma_tf32_ts(tmem_acc, a_raw + k * 8, b_lo + adv, kIdesc, 1u);
                            umma_tf32_ts(tmem_acc, a_lo + k * 8, b_hi + adv, kIdesc, 1u);
                        }
                        umma_commit(&t_empty[ts]);
                        umma_commit(&b_empty[bs]);
                        if (kb == c1 - 1) umma_commit(&acc_full[buf]);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp >= 4 && warp < 8) {
        // ------------------------------------------------------------------ split warps: smem A -> TMEM (raw, lo)
        const int q = warp - 4, r = q * 32 + lane;
        const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
        uint32_t it = 0;
        for (int u = blockIdx.x; u < units; u += gridDim.x) {
            const int z = u % p.splits;
            const int kb0 = z * p.kb_per_split, nkb = min(p.kb_per_split, total_kb - kb0);
            for (int kb = 0; kb < nkb; ++kb, ++it) {
                const uint32_t as = it % kAS, ts = it % kTS;
                mbar_wait(&a_full[as], (it / kAS) & 1);
                const uint32_t tile = smem_u32(smem_a + as * kATile);
                uint32_t v[32], lo[32];
                if (kAMajorM) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = lds32(tile + j * (BM * 4) + r * 4);
                } else {
                    const uint32_t row = tile + r * 128;
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const float4 f = lds128(row + ((c ^ (r & 7)) << 4));
                        v[4 * c + 0] = __float_as_uint(f.x); v[4 * c + 1] = __float_as_uint(f.y);
                        v[4 * c + 2] = __float_as_uint(f.z); v[4 * c + 3] = __float_as_uint(f.w);
                    }
                }
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    lo[j] = __float_as_uint(__uint_as_float(v[j]) - __uint_as_float(v[j] & 0xFFFFE000u));
                __syncwarp();
                if (lane == 0) mbar_arrive(&a_empty[as]);         // the tile is in registers: the slot may be refilled
                mbar_wait(&t_empty[ts], ((it / kTS) & 1) ^ 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t taddr = tmem_base + lane_base + kTCol0 + ts * kTStageCols;
                tmem_st32(taddr, v);
                tmem_st32(taddr + BK, lo);
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&t_full[ts]);
            }
        }
    } else if (warp >= 8) {
        // ------------------------------------------------------------------ drain + epilogue
        const int q = warp - 8;
        const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
        float acc[BN];
#pragma unroll
        for (int j = 0; j < BN; ++j) acc[j] = 0.0f;
        uint32_t ch = 0;
        const int64_t MN = static_cast<int64_t>(p.M) * p.N;
        for (int u = blockIdx.x; u < units; u += gridDim.x) {
            const int z = u % p.splits, tile = u / p.splits;
            const int m0 = (tile / p.tiles_n) * BM, n0 = (tile % p.tiles_n) * BN;
            const int kb0 = z * p.kb_per_split, nkb = min(p.kb_per_split, total_kb - kb0);
            for (int c0 = 0; c0 < nkb; c0 += kChunk, ++ch) {
                const uint32_t buf = ch & 1;
                mbar_wait(&acc_full[buf], (ch >> 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                for (int g = 0; g < BN / 32; ++g) {
                    uint32_t t[32];
                    tmem_ld32(tmem_base + lane_base + buf * BN + g * 32, t);
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc[g * 32 + j] += __uint_as_float(t[j]);
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[buf]);
            }
            const int row = m0 + q * 32 + lane;
            float* dst = p.C + (p.splits > 1 ? static_cast<int64_t>(z) * MN : 0);
            const float* bias = p.splits > 1 ? nullptr : p.bias;
            const bool relu = p.splits == 1 && p.relu;
            const float* mask = p.splits > 1 ? nullptr : p.mask;
            if (p.c_trans) {
#pragma unroll
                for (int j = 0; j < BN; ++j) {
                    const int n = n0 + j;
                    float x = acc[j];
                    if (n < p.N && row < p.M) {
                        if (bias != nullptr) x += bias[n];
                        dst[static_cast<int64_t>(n) * p.M + row] = relu ? fmaxf(x, 0.0f) : x;
                    }
                }
            } else if (row < p.M) {
                float* out = dst + static_cast<int64_t>(row) * p.N + n0;
                const float* mrow = mask != nullptr ? mask + static_cast<int64_t>(row) * p.N + n0 : nullptr;
                const bool vec_ok = (p.N % 4 == 0);
#pragma unroll
                for (int j = 0; j < BN; j += 4) {
                    const int n = n0 + j;
                    float o[4];
                    float mk[4] = {1.0f, 1.0f, 1.0f, 1.0f};
                    if (mrow != nullptr) {
                        if (vec_ok && n + 3 < p.N) {
                            const float4 m4 = *reinterpret_cast<const float4*>(mrow + j);
                            mk[0] = m4.x; mk[1] = m4.y; mk[2] = m4.z; mk[3] = m4.w;
                        } else {
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                if (n + e < p.N) mk[e] = mrow[j + e];
                        }
                    }
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float x = acc[j + e];
                        if (bias != nullptr && n + e < p.N) x += bias[n + e];
                        x = relu ? fmaxf(x, 0.0f) : x;
                        o[e] = mk[e] > 0.0f ? x : 0.0f;
                    }
                    if (vec_ok && n + 3 < p.N) {
                        *reinterpret_cast<float4*>(out + j) = make_float4(o[0], o[1], o[2], o[3]);
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            if (n + e < p.N) out[j + e] = o[e];
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < BN; ++j) acc[j] = 0.0f;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols));
    }
}

// out[i] = sum_z ws[z][i] (+ bias) (+ ReLU), z ascending (deterministic); the layout of `out` is the slices' layout:
// [M,N] (bias index i % N) or, c_trans, [N,M] (bias index i / M).
__global__ void __launch_bounds__(256)
ts_splitk_reduce_kernel(const float* __restrict__ ws, int splits, const float* __restrict__ bias, float* __restrict__ out,
                        int64_t MN, int M, int N, int relu, int c_trans, const float* __restrict__ mask) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= MN) return;
    float s = 0.0f;
    for (int z = 0; z < splits; ++z) s += ws[static_cast<int64_t>(z) * MN + i];
    if (bias != nullptr) s += bias[c_trans ? i / M : i % N];
    s = relu ? fmaxf(s, 0.0f) : s;
    out[i] = (mask == nullptr || mask[i] > 0.0f) ? s : 0.0f;
}

// lo = x - trunc_tf32(x), elementwise (the B operand's second term)
__global__ void __launch_bounds__(256)
split_lo_kernel(const float* __restrict__ src, float* __restrict__ lo, int64_t n) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float x = src[i];
        lo[i] = x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
    }
}

// dst[c][r] = src[r][c], dst_lo[c][r] = src[r][c] - trunc_tf32(src[r][c]); 32x32 tiles through padded shared memory
__global__ void __launch_bounds__(256)
transpose_split_kernel(const float* __restrict__ src, float* __restrict__ dst, float* __restrict__ dst_lo, int64_t rows,
                       int64_t cols, int64_t tiles_c) {
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
        if (r < rows && c < cols) {
            const float x = tile[tx][ty + j];
            dst[c * rows + r] = x;
            dst_lo[c * rows + r] = x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
        }
    }
}

// ---- host side ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline EncodeTiledFn encode_fn() {
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

// [rows, inner] fp32 row-major; box = box_rows x box_inner; swizzle 128B when the box's inner extent is 128 bytes
static inline bool make_map(CUtensorMap* map, const float* base, int64_t rows, int64_t inner, int box_rows, int box_inner) {
    EncodeTiledFn fn = encode_fn();
    if (fn == nullptr) return false;
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(inner), static_cast<cuuint64_t>(rows)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(inner) * 4};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(box_inner), static_cast<cuuint32_t>(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    const CUtensorMapSwizzle sw = box_inner * 4 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE;
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

struct Plan { int tiles_m, tiles_n, total_kb, splits, kb_per_split, grid; };

// k-split choice: minimise rounds x (k-blocks per unit + per-unit overhead) + the cost of the reduction pass
static inline Plan make_plan(int64_t M, int64_t N, int64_t K, int sms, bool allow_split) {
    Plan pl;
    pl.tiles_m = static_cast<int>((M + BM - 1) / BM);
    pl.tiles_n = static_cast<int>((N + BN - 1) / BN);
    pl.total_kb = static_cast<int>((K + BK - 1) / BK);
    const int64_t tiles = static_cast<int64_t>(pl.tiles_m) * pl.tiles_n;
    int best_s = 1;
    int64_t best_cost = INT64_MAX;
    for (int s = 1; s <= 16 && allow_split; ++s) {
        const int kbs = (pl.total_kb + s - 1) / s;
        if (s > 1 && kbs < kChunk) break;
        const int real_s = (pl.total_kb + kbs - 1) / kbs;
        if (real_s != s) continue;
        const int64_t rounds = (tiles * s + sms - 1) / sms;
        const int64_t cost = rounds * (kbs + 4) + (s > 1 ? 10 + 2 * s : 0);
        if (cost < best_cost) { best_cost = cost; best_s = s; }
    }
    pl.splits = best_s;
    pl.kb_per_split = (pl.total_kb + best_s - 1) / best_s;
    const int64_t units = tiles * best_s;
    pl.grid = static_cast<int>(units < sms ? units : sms);
    return pl;
}

static inline int64_t workspace_bytes(int64_t M, int64_t N, int64_t K, int sms) {
    const Plan pl = make_plan(M, N, K, sms, true);
    return pl.splits > 1 ? static_cast<int64_t>(pl.splits) * M * N * 4 : 0;
}

// Returns cudaSuccess / the launch error; cudaErrorInvalidValue when the tensor maps cannot be built.
static inline cudaError_t launch(const float* A, int a_mmajor, const float* B, const float* B_lo, const float* bias, float* C,
                                 int c_trans, int64_t M, int64_t N, int64_t K, int relu, float* workspace, int sms,
                                 cudaStream_t st, const float* mask = nullptr) {
    CUtensorMap ma, mb, mlo;
    const bool ok_a = a_mmajor ? make_map(&ma, A, K, M, BK, BM) : make_map(&ma, A, M, K, BM, BK);
    if (!ok_a || !make_map(&mb, B, N, K, BN, BK) || !make_map(&mlo, B_lo, N, K, BN, BK)) return cudaErrorInvalidValue;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(gemm_ts_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        cudaFuncSetAttribute(gemm_ts_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        attr_set = true;
    }
    const Plan pl = make_plan(M, N, K, sms, workspace != nullptr);
    Params p;
    p.C = pl.splits > 1 ? workspace : C;
    p.mask = mask;
    p.bias = bias; p.M = static_cast<int>(M); p.N = static_cast<int>(N); p.K = static_cast<int>(K);
    p.relu = relu; p.c_trans = c_trans; p.splits = pl.splits; p.kb_per_split = pl.kb_per_split;
    p.tiles_m = pl.tiles_m; p.tiles_n = pl.tiles_n;
    if (a_mmajor) gemm_ts_kernel<true><<<pl.grid, kThreads, kSmemBytes, st>>>(ma, mb, mlo, p);
    else gemm_ts_kernel<false><<<pl.grid, kThreads, kSmemBytes, st>>>(ma, mb, mlo, p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess || pl.splits == 1) return e;
    const int64_t MN = M * N;
    ts_splitk_reduce_kernel<<<static_cast<unsigned>((MN + 255) / 256), 256, 0, st>>>(
        workspace, pl.splits, bias, C, MN, static_cast<int>(M), static_cast<int>(N), relu, c_trans, mask);
    return cudaGetLastError();
}

}  // namespace gts
}  // namespace rl
