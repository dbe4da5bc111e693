// First conv layer of the AtariFf / AtariDqn networks on the INTEGER tensor cores (tcgen05.mma kind::i8),
// "v2" of the layer-1 kernels of conv_tc.cu.
//
// Reference: rlpyt/models/conv2d.py:36-44 with rlpyt/models/pg/atari_ff_model.py:31-35,50-53:
//     img.float().mul_(1/255) -> Conv2d(4->16, k8, s4, p0) -> ReLU          (uint8 frames [N,4,H,W])
//
// Why integers.  The frames are uint8: exact int8-MMA operands as they lie in HBM - no conversion,
// a quarter of the shared-memory bytes of a TF32 operand.  The fp32 filter bank is rewritten once per
// CTA as a 4-digit base-128 fixed-point number per output channel,
//     w = s_oc/64 * (q0 + q1/128 + q2/128^2 + q3/128^3),   q_i in [-64,64] (int8),  s_oc = 2^e > max|w[oc]|,
// (exact for every weight within 1/16 of its channel's largest, otherwise |error| <= 2^-28 s_oc), the
// four digits ride side by side on the MMA N axis (N = 4 x 16), the int32 accumulators are EXACT, and the
// epilogue recombines them in fp32: y = relu(s_oc/(64*255) * (((a3/128 + a2)/128 + a1)/128 + a0) + b).
// The result is closer to the exact real-number convolution than an fp32 FMA chain (tests: <= 3e-6 of
// sum|x||w| against fp64, measured ~1e-7).
//
// Why no im2col.  Space-to-depth by the stride turns k8s4 into a 2x2 stride-1 convolution over "cells"
// (cell (Y,X) = the 4x4 pixel block x[:, 4Y..4Y+3, 4X..4X+3], 64 bytes (c, ky', kx')):
//     out[oy, ox] = sum_{by,bx} cell(oy+by, ox+bx) . W4[by,bx]
// One shared-memory A row = the cell PAIR (r, r+1) of an image (128 bytes = bx 0|1 on the K axis), rows in
// cell order r = Y*GW + X; the by = 1 tap is the same tile read through a descriptor whose start address
// is advanced by GW rows (tools/probes/tcgen05_shift_probe.cu: a K-major SWIZZLE_128B descriptor may
// start at any row, base_offset 0 - the swizzle is a function of the absolute address bits).  Every
// input byte is written to the operand tile twice (im2col: four times, as 4-byte floats).
// Output rows with Y = GH-1 or X = GW-1 (9 %) are computed and dropped.
//
// Why bulk copies.  The v1 kernels were latency bound: register-staged 4-byte gathers kept ~10 KB per SM
// in flight (~0.8 TB/s).  Here a loader thread streams whole frames (28 KB, contiguous, row-gather
// index applied per frame) into a 4-deep shared-memory ring with cp.async.bulk (113 KB in flight per
// SM), producer warps re-lay them out with LDS.32 -> 2 x STS.32 (bank-conflict free), one warp issues
// 8 MMAs (M=128, N=64, K=32) per 128-row tile, four epilogue warps drain TMEM.
#pragma once
#include "tc_common.cuh"

namespace rl {
namespace c1i8 {

using namespace tc;

constexpr int kRows = 128;                   // cells (GEMM rows) per tile
constexpr int kSlotRows = 160;               // staged rows per tile: 128 + GW (by = 1 halo), GW <= 32
constexpr int kSlotBytes = kSlotRows * 128;  // 20 KiB
constexpr int kASlots = 3;
constexpr int kRaw = 4;                      // frames in flight per CTA
constexpr int kDigits = 4;
constexpr int kOC = 16;
constexpr int kN = kDigits * kOC;            // MMA N: digit-major, n = digit*16 + oc
constexpr int kBTile = kN * 128;             // one by tap: [64 rows x 128 B (bx, c, ky', kx')]
constexpr int kThreads = 448;                // warps 0-7 re-layout, 8-11 epilogue, 12 MMA + TMEM, 13 loader
constexpr int kProducerWarps = 8, kEpiWarp0 = 8, kMmaWarp = 12, kLoadWarp = 13;
constexpr int kTmemCols = 128;               // 2 accumulator buffers x 64 int32 columns

struct Geom {
    int n_img, H, W, GH, GW, OH, OW;
    int n_cells, n_tiles;                    // GH*GW, ceil(n_cells / 128)
    uint32_t img_bytes, raw_stage_bytes;     // 4*H*W, rounded up to 128
    uint32_t div_magic;                      // (cell * div_magic) >> 16 == cell / GW for every cell < 1024 (checked on the host)
};

__host__ __device__ constexpr uint32_t make_idesc_i8(int M, int N) {   // A u8 (0), B s8 (1), D s32 (2), K-major
    return (2u << 4) | (0u << 7) | (1u << 10) | (static_cast<uint32_t>(N >> 3) << 17) |
           (static_cast<uint32_t>(M >> 4) << 24);
}

__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void sts8(uint32_t addr, int v) {
    asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

struct SmemLayout {
    uint32_t b_off, a_off, raw_off, scale_off, bar_off, total;
    __host__ __device__ explicit SmemLayout(uint32_t raw_stage_bytes) {
        b_off = 0;
        a_off = 2 * kBTile;
        raw_off = a_off + kASlots * kSlotBytes;
        scale_off = raw_off + kRaw * raw_stage_bytes;
        bar_off = scale_off + 128;
        total = bar_off + 256 + 1024;        // + slack for the 1024-byte alignment of the base
    }
};

// Filter bank -> base-128 digits in the K-major SWIZZLE_128B B tiles (see the header); scale[oc] = s_oc/(64*255).
__device__ __forceinline__ void quantize_filters(const float* __restrict__ Wg, const float* __restrict__ bias,
                                                 uint32_t b_u32, float* scale_bias) {
    // s_oc: power of two strictly above the channel's largest |w| (0 -> 1)
    if (threadIdx.x < 256) {
        const int oc = threadIdx.x >> 4, part = threadIdx.x & 15;
        float m = 0.0f;
#pragma unroll
        for (int i = 0; i < 16; ++i) m = fmaxf(m, fabsf(Wg[oc * 256 + part * 16 + i]));
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (part == 0) {
            int e = 0;
            if (m > 0.0f) frexpf(m, &e);         // m = f * 2^e, f in [0.5, 1)  ->  2^e > m
            const float s = ldexpf(1.0f, e);
            scale_bias[oc] = s;
            scale_bias[16 + oc] = bias[oc];
        }
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < kOC * 256; idx += kThreads) {
        const int oc = idx >> 8, c = (idx >> 6) & 3, ky = (idx >> 3) & 7, kx = idx & 7;
        const float s = scale_bias[oc];
        float r = Wg[idx] / s * 64.0f;           // exact: s is a power of two
        const int by = ky >> 2, kk = (kx >> 2) * 64 + c * 16 + (ky & 3) * 4 + (kx & 3);
#pragma unroll
        for (int d = 0; d < kDigits; ++d) {
            const float q = rintf(r);
            r = (r - q) * 128.0f;                // exact (the difference has at most 24 significant bits)
            const int n = d * kOC + oc;
            sts8(b_u32 + static_cast<uint32_t>(by * kBTile + n * 128 + (((kk >> 4) ^ (n & 7)) << 4) + (kk & 15)),
                 static_cast<int>(q));
        }
    }
}

// Re-layout of one 128-cell tile of a frame (raw_base = the frame's channel c in the shared-memory ring) into a
// cell-pair operand slot: group u of this warp = cells 32*(g0 + gstep*u) + lane of the tile.  A thread reads the
// 4 pixel rows of its cell's channel (4 x LDS.32, lanes = consecutive words) - exactly one 16-byte chunk of an
// operand row - and writes it twice with STS.128: as the left half of row q and the right half of row q-1 (8
// consecutive rows cover the 8 swizzled chunk positions: conflict free).  All loads are issued before the wait
// on the slot and the stores.
template <int NU>
__device__ __forceinline__ void stage_tile(const Geom& g, uint32_t raw_base, uint32_t slot_base, int t, int c, int g0, int gstep,
                                           int lane, uint64_t* slot_empty, uint32_t empty_parity) {
    const int q_end = kRows + g.GW;                          // cells [0, q_end] of the tile are needed
    const uint32_t W4 = static_cast<uint32_t>(g.W);
    uint32_t w[NU][4];
    bool ok[NU];
#pragma unroll
    for (int u = 0; u < NU; ++u) {
        const int q = 32 * (g0 + gstep * u) + lane;
        const int cell = t * kRows + q;
        ok[u] = q <= q_end && cell < g.n_cells;
        const int Yc = static_cast<int>((static_cast<uint32_t>(cell) * g.div_magic) >> 16);
        const int Xc = cell - Yc * g.GW;
        const uint32_t src = raw_base + static_cast<uint32_t>(4 * Yc) * W4 + static_cast<uint32_t>(4 * Xc);
#pragma unroll
        for (int k = 0; k < 4; ++k) w[u][k] = ok[u] ? lds32(src + static_cast<uint32_t>(k) * W4) : 0u;
    }
    mbar_wait(slot_empty, empty_parity);
#pragma unroll
    for (int u = 0; u < NU; ++u) {
        const int q = 32 * (g0 + gstep * u) + lane;
        if (ok[u] && q < q_end) sts128u(slot_base + static_cast<uint32_t>(q * 128 + ((c ^ (q & 7)) << 4)), w[u]);
        if (ok[u] && q >= 1)
            sts128u(slot_base + static_cast<uint32_t>((q - 1) * 128 + (((4 + c) ^ ((q - 1) & 7)) << 4)), w[u]);
    }
}

__global__ void __launch_bounds__(kThreads, 1)
conv1_i8_fwd_kernel(const uint8_t* __restrict__ X, const int64_t* __restrict__ rows, const float* __restrict__ Wg,
                    const float* __restrict__ bias, float* __restrict__ Y, Geom g, int relu, uint8_t* __restrict__ copy_out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const SmemLayout L(g.raw_stage_bytes);
    const uint32_t smem_u = smem_u32(smem);
    const uint32_t b_u32 = smem_u + L.b_off, a_u32 = smem_u + L.a_off, raw_u32 = smem_u + L.raw_off;
    float* scale_bias = reinterpret_cast<float*>(smem + L.scale_off);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bar_off);
    uint64_t* raw_full = bars;                       // [kRaw]  loader (tx bytes) -> re-layout warps
    uint64_t* raw_empty = raw_full + kRaw;           // [kRaw]  re-layout warps -> loader
    uint64_t* a_full = raw_empty + kRaw;             // [kASlots] re-layout warps -> MMA
    uint64_t* a_empty = a_full + kASlots;            // [kASlots] MMA (commit) -> re-layout warps
    uint64_t* acc_full = a_empty + kASlots;          // [2] MMA -> epilogue
    uint64_t* acc_empty = acc_full + 2;              // [2] epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    constexpr uint32_t kIdesc = make_idesc_i8(kRows, kN);

    if (threadIdx.x == 0) {
        for (int s = 0; s < kRaw; ++s) { mbar_init(&raw_full[s], 1); mbar_init(&raw_empty[s], kProducerWarps); }
        for (int s = 0; s < kASlots; ++s) { mbar_init(&a_full[s], kProducerWarps); mbar_init(&a_empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kMmaWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    quantize_filters(Wg, bias, b_u32, scale_bias);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = uniform_u32(*tmem_slot);
    const int n_local = (g.n_img - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);

    if (warp == kLoadWarp) {
        // ================================================================ frame loader
        if (elect_one()) {
            for (int i = 0; i < n_local; ++i) {
                const int rs = i % kRaw;
                mbar_wait(&raw_empty[rs], ((i / kRaw) & 1) ^ 1);
                const int64_t n = static_cast<int64_t>(blockIdx.x) + static_cast<int64_t>(i) * gridDim.x;
                const int64_t img = rows != nullptr ? rows[n] : n;
                mbar_expect_tx(&raw_full[rs], g.img_bytes);
                bulk_load(raw_u32 + static_cast<uint32_t>(rs) * g.raw_stage_bytes, X + img * g.img_bytes,
                          g.img_bytes, &raw_full[rs]);
            }
        }
        __syncwarp();
    } else if (warp < kProducerWarps) {
        // ================================================================ re-layout: frame -> cell-pair rows
        // warp = (32-cell group parity | channel c), see stage_tile
        const int c = warp & 3, gpar = warp >> 2;
        const uint32_t src_ch = static_cast<uint32_t>(c * g.H * g.W);
        uint32_t it = 0;                                     // global tile counter (slot ring)
        for (int i = 0; i < n_local; ++i) {
            const int rs = i % kRaw;
            mbar_wait(&raw_full[rs], (i / kRaw) & 1);
            const uint32_t raw_base = raw_u32 + static_cast<uint32_t>(rs) * g.raw_stage_bytes + src_ch;
            // copy_out: the frame that just landed in shared memory also goes to HBM (frame n of copy_out) by one bulk
            // store - the sampler's step: frames stream in from the page-locked step buffer over PCIe ONCE, feeding
            // both agent.step's first layer and observation[t] of the resident [T,B] batch (no separate H2D in front
            // of the network).
            const bool storer = copy_out != nullptr && warp == 0 && lane == 0;
            if (storer) {
                const int64_t n = static_cast<int64_t>(blockIdx.x) + static_cast<int64_t>(i) * gridDim.x;
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(reinterpret_cast<uint64_t>(copy_out + n * g.img_bytes)),
                             "r"(raw_u32 + static_cast<uint32_t>(rs) * g.raw_stage_bytes), "r"(g.img_bytes) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            for (int t = 0; t < g.n_tiles; ++t, ++it) {
                const int slot = it % kASlots;
                stage_tile<3>(g, raw_base, a_u32 + static_cast<uint32_t>(slot * kSlotBytes), t, c, gpar, 2, lane,
                              &a_empty[slot], ((it / kASlots) & 1) ^ 1);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&a_full[slot]);
            }
            if (storer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the stage may be refilled once the store has read it
            __syncwarp();
            if (lane == 0) mbar_arrive(&raw_empty[rs]);
        }
    } else if (warp == kMmaWarp) {
        // ================================================================ MMA issuer
        uint32_t it = 0;
        for (int i = 0; i < n_local; ++i) {
            for (int t = 0; t < g.n_tiles; ++t, ++it) {
                const int slot = it % kASlots, buf = it & 1;
                mbar_wait(&acc_empty[buf], ((it >> 1) & 1) ^ 1);
                mbar_wait(&a_full[slot], (it / kASlots) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (elect_one()) {
                    const uint32_t acc = tmem_base + static_cast<uint32_t>(buf * kN);
#pragma unroll
                    for (int by = 0; by < 2; ++by) {
                        const uint64_t da = make_desc(smem + L.a_off + slot * kSlotBytes + by * g.GW * 128);
                        const uint64_t db = make_desc(smem + L.b_off + by * kBTile);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_i8(acc, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), kIdesc,
                                    (by | k) ? 1u : 0u);
                    }
                    umma_commit(&a_empty[slot]);
                    umma_commit(&acc_full[buf]);
                }
                __syncwarp();
            }
        }
    } else {
        // ================================================================ epilogue (warps 8..11)
        const int qw = warp - kEpiWarp0;
        const uint32_t lane_base = static_cast<uint32_t>(qw * 32) << 16;
        const int P = g.OH * g.OW;
        float mul[kOC], add[kOC];                            // s_oc / (64 * 255 * 128), bias: registers for the whole kernel
#pragma unroll
        for (int oc = 0; oc < kOC; ++oc) {
            mul[oc] = scale_bias[oc] * (0.015625f * 0.0078125f) * (1.0f / 255.0f);
            add[oc] = scale_bias[16 + oc];
        }
        uint32_t it = 0;
        for (int i = 0; i < n_local; ++i) {
            const int64_t n = static_cast<int64_t>(blockIdx.x) + static_cast<int64_t>(i) * gridDim.x;
            for (int t = 0; t < g.n_tiles; ++t, ++it) {
                const int buf = it & 1;
                mbar_wait(&acc_full[buf], (it >> 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const int cell = t * kRows + qw * 32 + lane;
                const bool warp_live = t * kRows + qw * 32 < g.n_cells;       // warp-uniform
                uint32_t r0[32], r1[32];
                if (warp_live) {
                    tmem_ld32(tmem_base + lane_base + static_cast<uint32_t>(buf * kN), r0);        // digits 0, 1
                    tmem_ld32(tmem_base + lane_base + static_cast<uint32_t>(buf * kN + 32), r1);   // digits 2, 3
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[buf]);
                if (!warp_live) continue;
                const int Yc = static_cast<int>((static_cast<uint32_t>(cell) * g.div_magic) >> 16);
                const int Xc = cell - Yc * g.GW;
                if (cell < g.n_cells && Yc < g.OH && Xc < g.OW) {
                    float* yo = Y + n * (static_cast<int64_t>(kOC) * P) + Yc * g.OW + Xc;
#pragma unroll
                    for (int oc = 0; oc < kOC; ++oc) {
                        // digit pairs combined in int32 (|a_d| < 2^23, so a*128 + a' < 2^31), then two conversions
                        const int t01 = static_cast<int>(r0[oc]) * 128 + static_cast<int>(r0[16 + oc]);
                        const int t23 = static_cast<int>(r1[oc]) * 128 + static_cast<int>(r1[16 + oc]);
                        float v = fmaf(static_cast<float>(t23), 1.0f / 16384.0f, static_cast<float>(t01));   // 128 x the digit sum
                        v = fmaf(v, mul[oc], add[oc]);
                        if (relu) v = fmaxf(v, 0.0f);
                        yo[static_cast<int64_t>(oc) * P] = v;
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == kMmaWarp)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols));
}

inline bool geom_ok(int C, int H, int W) {
    if (!(C == 4 && H >= 8 && W >= 8 && H % 4 == 0 && W % 4 == 0 && W / 4 <= 32 && W / 4 >= 2)) return false;
    const uint32_t GW = W / 4, n_cells = (H / 4) * GW;
    if (n_cells + 192 >= 1024) return false;                 // 16-bit reciprocal division below, rows per frame
    const uint32_t magic = (65536u + GW - 1u) / GW;
    for (uint32_t cell = 0; cell < 1024; ++cell)
        if (((cell * magic) >> 16) != cell / GW) return false;
    return true;
}

inline Geom make_geom(int64_t N, int H, int W) {
    Geom g;
    g.n_img = static_cast<int>(N); g.H = H; g.W = W; g.GH = H / 4; g.GW = W / 4; g.OH = g.GH - 1; g.OW = g.GW - 1;
    g.n_cells = g.GH * g.GW; g.n_tiles = (g.n_cells + kRows - 1) / kRows;
    g.img_bytes = static_cast<uint32_t>(4 * H * W);
    g.raw_stage_bytes = (g.img_bytes + 127u) & ~127u;
    g.div_magic = (65536u + static_cast<uint32_t>(g.GW) - 1u) / static_cast<uint32_t>(g.GW);
    return g;
}

// max frame size for which the 4-deep raw ring fits next to the operand tiles (227 KB per CTA)
inline bool smem_ok(const Geom& g) { return SmemLayout(g.raw_stage_bytes).total <= 232448u; }

inline cudaError_t launch_fwd(const uint8_t* X, const int64_t* rows, const float* W, const float* bias, float* Y,
                              const Geom& g, int relu, int sms, cudaStream_t st, uint8_t* copy_out = nullptr) {
    const SmemLayout L(g.raw_stage_bytes);
    static uint32_t attr_bytes = 0;
    if (L.total > attr_bytes) {
        cudaError_t e = cudaFuncSetAttribute(conv1_i8_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(L.total));
        if (e != cudaSuccess) return e;
        attr_bytes = L.total;
    }
    const int grid = g.n_img < sms ? g.n_img : sms;
    conv1_i8_fwd_kernel<<<static_cast<unsigned>(grid), kThreads, L.total, st>>>(X, rows, W, bias, Y, g, relu, copy_out);
    return cudaGetLastError();
}


// ====================================================================================================
// Weight + bias gradient of the same layer on the integer tensor cores.
//     dW[oc, c, 4by+ky', 4bx+kx'] = 1/255 * sum_{n, cells r} g[n, oc, r] * A_n[r + by*GW][(bx, c, ky', kx')]
// with g = grad_out * (out > 0) at the output position of cell r (0 for the dropped cells) and A_n the SAME
// cell-pair rows the forward kernel stages - read here as an MN-major operand (M = the 128 bytes of a row,
// K = cells; tools/probes/conv1_i8_probe.cu checks MN-major SWIZZLE_128B with a K-direction start shift).
// g is rewritten as 4 base-128 digits against one power-of-two scale per output channel,
//     g = S_oc/64 * (q0 + q1/128 + q2/128^2 + q3/128^3),  S_oc = 2^e > max |g[:, oc]|   (absmax_kernel),
// i.e. 28 bits below the channel's largest gradient of the minibatch (an fp32 accumulation of the same sum
// carries 24 bits below its running value), the digits ride on the MMA N axis (N = 4 x 16) as a K-major B
// tile [64 x 128 cells], and the int32 accumulators - EXACT, 2 x (128 lanes x 64 columns) of TMEM - run over
// all frames of the CTA (|acc| <= 255 * 64 * cells: a CTA may own 256 frames).  Each CTA dumps its integers;
// wgrad_i8_reduce_kernel adds them over CTAs and digits in fp64 in a fixed order and rounds once.
// Frames and gradients arrive by cp.async.bulk (2 + 2 images in flight per CTA); only the ReLU mask source
// `out` is read with ordinary loads, prefetched one tile ahead.
namespace wg {

constexpr int kXRaw = 2, kGRaw = 2, kSlots = 3;
constexpr int kBSlotBytes = kN * 128;                // [64 rows (digit, oc)] x [128 cells], K-major SWIZZLE_128B
constexpr int kAWarps = 4, kBWarp0 = 4, kBWarps = 8; // warps 0-3 frames -> A (+ final drain), 4-11 g -> digits, 12 MMA, 13 loader
constexpr int kMaxFramesPerCta = 256;
constexpr float kRintMagic = 12582912.0f;            // 1.5 * 2^23: (t + magic) - magic = rint(t), low byte of the sum = int8(rint(t))

struct SmemLayout {
    uint32_t a_off, b_off, x_off, g_off, bar_off, total, g_stage_bytes;
    __host__ __device__ SmemLayout(uint32_t raw_stage_bytes, uint32_t g_bytes) {
        g_stage_bytes = (g_bytes + 127u) & ~127u;
        a_off = 0;
        b_off = a_off + kSlots * kSlotBytes;
        x_off = b_off + kSlots * kBSlotBytes;
        g_off = x_off + kXRaw * raw_stage_bytes;
        bar_off = g_off + kGRaw * g_stage_bytes;
        total = bar_off + 256 + 1024;
    }
};

// max |g[:, oc]| as the bit pattern of a non-negative float (atomicMax on ints orders them correctly)
__global__ void __launch_bounds__(256)
absmax_kernel(const float* __restrict__ G, int64_t n_rows, int P, unsigned int* __restrict__ gmax_bits) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;      // multiple of 16: one channel per warp
    float m = 0.0f;
    for (int64_t r = warp0; r < n_rows; r += n_warps) {
        const float* row = G + r * P;
        for (int i = lane; i < P; i += 32) m = fmaxf(m, fabsf(ldg_stream(row + i)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0 && warp0 < n_rows) atomicMax(gmax_bits + (warp0 & 15), __float_as_uint(m));
}

__global__ void __launch_bounds__(kThreads, 1)
conv1_i8_wgrad_kernel(const uint8_t* __restrict__ X, const int64_t* __restrict__ rows, const float* __restrict__ Out,
                      const float* __restrict__ G, const float* __restrict__ gmax, int* __restrict__ partial,
                      float* __restrict__ partial_bias, Geom g) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int P = g.OH * g.OW;
    const uint32_t g_bytes = static_cast<uint32_t>(kOC * P * 4);
    const SmemLayout L(g.raw_stage_bytes, g_bytes);
    const uint32_t smem_u = smem_u32(smem);
    const uint32_t a_u32 = smem_u + L.a_off, b_u32 = smem_u + L.b_off, x_u32 = smem_u + L.x_off, g_u32 = smem_u + L.g_off;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bar_off);
    uint64_t* x_full = bars;                         // [kXRaw]
    uint64_t* x_empty = x_full + kXRaw;
    uint64_t* g_full = x_empty + kXRaw;              // [kGRaw]
    uint64_t* g_empty = g_full + kGRaw;
    uint64_t* a_full = g_empty + kGRaw;              // [kSlots]
    uint64_t* a_empty = a_full + kSlots;
    uint64_t* b_full = a_empty + kSlots;             // [kSlots]
    uint64_t* b_empty = b_full + kSlots;
    uint64_t* acc_full = b_empty + kSlots;           // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);

    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    constexpr uint32_t kIdesc = make_idesc_i8(kRows, kN) | (1u << 15);      // A MN-major, B K-major

    if (threadIdx.x == 0) {
        for (int s = 0; s < kXRaw; ++s) { mbar_init(&x_full[s], 1); mbar_init(&x_empty[s], kAWarps); }
        for (int s = 0; s < kGRaw; ++s) { mbar_init(&g_full[s], 1); mbar_init(&g_empty[s], kBWarps); }
        for (int s = 0; s < kSlots; ++s) {
            mbar_init(&a_full[s], kAWarps); mbar_init(&a_empty[s], 1);
            mbar_init(&b_full[s], kBWarps); mbar_init(&b_empty[s], 1);
        }
        mbar_init(acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kMmaWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = uniform_u32(*tmem_slot);
    const int n_local = (g.n_img - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);

    if (warp == kLoadWarp) {
        // ================================================================ loader: frames and gradients
        if (elect_one()) {
            for (int i = 0; i < n_local; ++i) {
                const int s = i & 1;
                const uint32_t ph = ((i >> 1) & 1) ^ 1;
                const int64_t n = static_cast<int64_t>(blockIdx.x) + static_cast<int64_t>(i) * gridDim.x;
                const int64_t img = rows != nullptr ? rows[n] : n;
                mbar_wait(&x_empty[s], ph);
                mbar_expect_tx(&x_full[s], g.img_bytes);
                bulk_load(x_u32 + static_cast<uint32_t>(s) * g.raw_stage_bytes, X + img * g.img_bytes, g.img_bytes, &x_full[s]);
                mbar_wait(&g_empty[s], ph);
                mbar_expect_tx(&g_full[s], g_bytes);
                bulk_load(g_u32 + static_cast<uint32_t>(s) * L.g_stage_bytes, G + n * (static_cast<int64_t>(kOC) * P), g_bytes,
                          &g_full[s]);
            }
        }
        __syncwarp();
    } else if (warp < kAWarps) {
        // ================================================================ frames -> cell-pair rows (MN-major A)
        const int c = warp;
        const uint32_t src_ch = static_cast<uint32_t>(c * g.H * g.W);
        uint32_t it = 0;
        for (int i = 0; i < n_local; ++i) {
            const int s = i & 1;
            mbar_wait(&x_full[s], (i >> 1) & 1);
            const uint32_t raw_base = x_u32 + static_cast<uint32_t>(s) * g.raw_stage_bytes + src_ch;
            for (int t = 0; t < g.n_tiles; ++t, ++it) {
                const int slot = it % kSlots;
                stage_tile<5>(g, raw_base, a_u32 + static_cast<uint32_t>(slot * kSlotBytes), t, c, 0, 1, lane, &a_empty[slot],
                              ((it / kSlots) & 1) ^ 1);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&a_full[slot]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&x_empty[s]);
        }
        // ---- final drain: this CTA's exact integer sums, [by][m = TMEM lane][n = digit*16 + oc]
        mbar_wait(acc_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
        int* dst = partial + (static_cast<int64_t>(blockIdx.x) * 2 * kRows + warp * 32 + lane) * kN;
#pragma unroll
        for (int by = 0; by < 2; ++by) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t r[32];
                if (n_local > 0) {
                    tmem_ld32(tmem_base + lane_base + static_cast<uint32_t>(by * kN + h * 32), r);
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) r[j] = 0u;
                }
                int4* d4 = reinterpret_cast<int4*>(dst + static_cast<int64_t>(by) * kRows * kN + h * 32);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    d4[j] = make_int4(static_cast<int>(r[4 * j]), static_cast<int>(r[4 * j + 1]), static_cast<int>(r[4 * j + 2]),
                                      static_cast<int>(r[4 * j + 3]));
            }
        }
    } else if (warp < kBWarp0 + kBWarps) {
        // ================================================================ gradients -> base-128 digits (K-major B)
        // warp wb owns channels wb and wb+8; lane = cell quad of the tile (cells 4*lane .. 4*lane+3): its four
        // digit bytes per cell quad are one 32-bit word of B row (digit, oc); a warp writes whole 128-byte rows.
        const int wb = warp - kBWarp0;
        float mult[2], bias_acc[2] = {0.0f, 0.0f};
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const float m = gmax[wb + 8 * k];
            int e = 0;
            if (m > 0.0f) frexpf(m, &e);
            mult[k] = ldexpf(64.0f, -e);                     // 64 / S_oc, S_oc = 2^e > max |g|
        }
        const bool has_mask = Out != nullptr;
        float ovn[2][4];
        auto fetch_out = [&](int i, int t, float (&o)[2][4]) {
            const int64_t n = static_cast<int64_t>(blockIdx.x) + static_cast<int64_t>(i) * gridDim.x;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int cell = t * kRows + 4 * lane + j;
                const int Yc = static_cast<int>((static_cast<uint32_t>(cell) * g.div_magic) >> 16);
                const int Xc = cell - Yc * g.GW;
                const bool valid = has_mask && cell < g.n_cells && Yc < g.OH && Xc < g.OW;
                const int pos = Yc * g.OW + Xc;
#pragma unroll
                for (int k = 0; k < 2; ++k)
                    o[k][j] = valid ? ldg_stream(Out + (n * kOC + wb + 8 * k) * P + pos) : 1.0f;
            }
        };
        if (n_local > 0) fetch_out(0, 0, ovn);
        uint32_t it = 0;
        for (int i = 0; i < n_local; ++i) {
            const int s = i & 1;
            mbar_wait(&g_full[s], (i >> 1) & 1);
            const uint32_t g_base = g_u32 + static_cast<uint32_t>(s) * L.g_stage_bytes;
            for (int t = 0; t < g.n_tiles; ++t, ++it) {
                const int slot = it % kSlots;
                float ov[2][4];
#pragma unroll
                for (int k = 0; k < 2; ++k)
#pragma unroll
                    for (int j = 0; j < 4; ++j) ov[k][j] = ovn[k][j];
                {   // prefetch the mask source of the next tile (possibly of the next frame)
                    int i2 = i, t2 = t + 1;
                    if (t2 == g.n_tiles) { t2 = 0; ++i2; }
                    if (i2 < n_local) fetch_out(i2, t2, ovn);
                }
                float gv[2][4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int cell = t * kRows + 4 * lane + j;
                    const int Yc = static_cast<int>((static_cast<uint32_t>(cell) * g.div_magic) >> 16);
                    const int Xc = cell - Yc * g.GW;
                    const bool valid = cell < g.n_cells && Yc < g.OH && Xc < g.OW;
                    const uint32_t pos4 = static_cast<uint32_t>(Yc * g.OW + Xc) * 4u;
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        const float v = valid ? __uint_as_float(lds32(g_base + static_cast<uint32_t>((wb + 8 * k) * P) * 4u + pos4)) : 0.0f;
                        gv[k][j] = ov[k][j] > 0.0f ? v : 0.0f;
                    }
                }
                mbar_wait(&b_empty[slot], ((it / kSlots) & 1) ^ 1);
                const uint32_t bslot = b_u32 + static_cast<uint32_t>(slot * kBSlotBytes);
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const int oc = wb + 8 * k;
                    float tq[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        bias_acc[k] += gv[k][j];
                        tq[j] = gv[k][j] * mult[k];
                    }
#pragma unroll
                    for (int d = 0; d < kDigits; ++d) {
                        uint32_t u[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float uf = tq[j] + kRintMagic;            // low byte = int8(rint(t))
                            const float q = uf - kRintMagic;
                            tq[j] = (tq[j] - q) * 128.0f;                   // exact
                            u[j] = __float_as_uint(uf);
                        }
                        const uint32_t word = __byte_perm(__byte_perm(u[0], u[1], 0x0040), __byte_perm(u[2], u[3], 0x0040), 0x5410);
                        const int nrow = d * kOC + oc;
                        sts32u(bslot + static_cast<uint32_t>(nrow * 128 + (((lane >> 2) ^ (nrow & 7)) << 4) + (lane & 3) * 4), word);
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&b_full[slot]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&g_empty[s]);
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const float sum = warp_sum(bias_acc[k]);
            if (lane == 0) partial_bias[static_cast<int64_t>(blockIdx.x) * kOC + wb + 8 * k] = sum;
        }
    } else if (warp == kMmaWarp) {
        // ================================================================ MMA issuer
        uint32_t it = 0;
        for (int i = 0; i < n_local; ++i) {
            for (int t = 0; t < g.n_tiles; ++t, ++it) {
                const int slot = it % kSlots;
                const uint32_t ph = (it / kSlots) & 1;
                mbar_wait(&a_full[slot], ph);
                mbar_wait(&b_full[slot], ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (elect_one()) {
#pragma unroll
                    for (int by = 0; by < 2; ++by) {
                        const uint32_t acc = tmem_base + static_cast<uint32_t>(by * kN);
                        const uint64_t da = make_desc(smem + L.a_off + slot * kSlotBytes + by * g.GW * 128);
                        const uint64_t db = make_desc(smem + L.b_off + slot * kBSlotBytes);
#pragma unroll
                        for (int k = 0; k < 4; ++k)        // K = 32 cells: A advances 32 rows (4096 B), B 32 bytes
                            umma_i8(acc, da + static_cast<uint64_t>(256 * k), db + static_cast<uint64_t>(2 * k), kIdesc,
                                    (it > 0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit(&a_empty[slot]);
                    umma_commit(&b_empty[slot]);
                }
                __syncwarp();
            }
        }
        if (elect_one()) umma_commit(acc_full);
        __syncwarp();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == kMmaWarp)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols));
}

// dW[oc][c][ky][kx] = S_oc / (64 * 255) * sum_cta sum_d 128^-d * partial[cta][by][m][d*16 + oc]  (fp64, fixed order),
// db[oc] = sum_cta partial_bias[cta][oc].  One block per (by, m) row of 64 ints: thread = (CTA slice p of 8, oc);
// every thread sums its CTAs (p, p+8, ...) in order, the 8 slices are added in order through shared memory.
__global__ void __launch_bounds__(128)
wgrad_i8_reduce_kernel(const int* __restrict__ partial, const float* __restrict__ partial_bias, const float* __restrict__ gmax,
                       int n_cta, float* __restrict__ dW, float* __restrict__ db) {
    __shared__ double part[8][kOC];
    const int row = blockIdx.x;                                  // by * 128 + m
    const int oc = threadIdx.x & 15, p = threadIdx.x >> 4;
    if (row < 2 * kRows) {
        double acc = 0.0;
        for (int cta = p; cta < n_cta; cta += 8) {
            const int* q = partial + (static_cast<int64_t>(cta) * 2 * kRows + row) * kN + oc;
            acc += static_cast<double>(q[0]) + static_cast<double>(q[16]) * (1.0 / 128.0) + static_cast<double>(q[32]) * (1.0 / 16384.0) +
                   static_cast<double>(q[48]) * (1.0 / 2097152.0);
        }
        part[p][oc] = acc;
        __syncthreads();
        if (p == 0) {
            double t = 0.0;
#pragma unroll
            for (int k = 0; k < 8; ++k) t += part[k][oc];
            int e = 0;
            const float gm = gmax[oc];
            if (gm > 0.0f) frexpf(gm, &e);
            const double scale = ldexp(1.0, e) / (64.0 * 255.0);
            const int by = row >> 7, m = row & 127;
            const int bx = m >> 6, c = (m >> 4) & 3, kyp = (m >> 2) & 3, kxp = m & 3;
            dW[((oc * 4 + c) * 8 + 4 * by + kyp) * 8 + 4 * bx + kxp] = static_cast<float>(t * scale);
        }
    } else if (db != nullptr && threadIdx.x < kOC) {             // the extra block: bias
        double bsum = 0.0;
        for (int cta = 0; cta < n_cta; ++cta) bsum += static_cast<double>(partial_bias[cta * kOC + threadIdx.x]);
        db[threadIdx.x] = static_cast<float>(bsum);
    }
}

inline size_t scratch_bytes(int sms) {
    return 64 + static_cast<size_t>(sms) * (2 * kRows * kN * sizeof(int) + kOC * sizeof(float));
}
inline bool smem_ok(const Geom& g) {
    return SmemLayout(g.raw_stage_bytes, static_cast<uint32_t>(kOC * g.OH * g.OW * 4)).total <= 232448u;
}

// scratch: [16 floats gmax | partial ints | partial bias]; out may be null (grad_out already masked)
inline cudaError_t launch_wgrad(const uint8_t* X, const int64_t* rows, const float* Out, const float* G, float* dW, float* db,
                                const Geom& g, int sms, void* scratch, cudaStream_t st, const float* chan_absmax = nullptr) {
    const int P = g.OH * g.OW;
    const SmemLayout L(g.raw_stage_bytes, static_cast<uint32_t>(kOC * P * 4));
    static uint32_t attr_bytes = 0;
    if (L.total > attr_bytes) {
        cudaError_t e = cudaFuncSetAttribute(conv1_i8_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(L.total));
        if (e != cudaSuccess) return e;
        attr_bytes = L.total;
    }
    float* gmax = static_cast<float*>(scratch);
    int* partial = reinterpret_cast<int*>(static_cast<uint8_t*>(scratch) + 64);
    const int grid = g.n_img < sms ? g.n_img : sms;
    float* partial_bias = reinterpret_cast<float*>(partial + static_cast<size_t>(grid) * 2 * kRows * kN);
    if (chan_absmax != nullptr) {    // the producer of G already knows max |G[:, oc]| (conv2's input-gradient epilogue)
        gmax = const_cast<float*>(chan_absmax);
    } else {
        cudaError_t e = cudaMemsetAsync(gmax, 0, 64, st);
        if (e != cudaSuccess) return e;
        absmax_kernel<<<sms * 8, 256, 0, st>>>(G, static_cast<int64_t>(g.n_img) * kOC, P, reinterpret_cast<unsigned int*>(gmax));
    }
    conv1_i8_wgrad_kernel<<<static_cast<unsigned>(grid), kThreads, L.total, st>>>(X, rows, Out, G, gmax, partial, partial_bias, g);
    wgrad_i8_reduce_kernel<<<2 * kRows + 1, 128, 0, st>>>(partial, partial_bias, gmax, grid, dW, db);
    return cudaGetLastError();
}

}  // namespace wg

}  // namespace c1i8
}  // namespace rl
