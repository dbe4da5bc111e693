// Second conv layer of the AtariFf network, "v2": Conv2d(16->32, k4, s2, p1) (+ReLU) on fp32 NCHW activations
// without im2col expansion (rlpyt/models/conv2d.py:36-44 with rlpyt/models/pg/atari_ff_model.py:31-35).
//
// Space-to-depth by the stride turns k4s2p1 into a 2x2 stride-1 convolution over "cells": cell (Y,X) of an image
// = the 2x2 pixel block x[:, 2Y-1 .. 2Y, 2X-1 .. 2X] of the zero-padded plane, 64 floats (c, dy, dx),
//     out[oy, ox] = sum_{by,bx} cell(oy+by, ox+bx) . W4[by,bx],   W4[by,bx][oc][(c,dy,dx)] = w[oc, c, 2by+dy, 2bx+dx]
// One GEMM row per cell (r = Y*GW + X, GH = OH+1, GW = OW+1; a 20x20 plane has 121 cells = ONE 128-row tile per
// image), K = 64 floats = two K-major SWIZZLE_128B k-block planes (channels 0-7 | 8-15); the four taps are four
// ROW-SHIFTED descriptors into the same planes (shift = by*GW + bx rows, tools/probes/tcgen05_shift_probe.cu).
// Every activation is read from HBM once (whole images stream in by cp.async.bulk, 2 in flight per CTA), converted
// once - hi = the raw fp32 (the tensor core truncates it to TF32), lo = x - trunc(x) - and written once per term;
// the v1 kernel (conv_tc.cu) gathered every element four times with 4-byte register-staged loads and was
// latency-bound at ~0.15 of its HBM floor.  fp32-accurate 3-term TF32 split as in gemm_tf32x3.cu; the two k-block
// planes accumulate in separate TMEM columns (halves the accumulator truncation) and double as the pipeline
// stages: the MMAs of plane 0 run while the producers write plane 1.
#pragma once
#include "tc_common.cuh"

namespace rl {
namespace c2s {

using namespace tc;

constexpr int kRows = 128;
constexpr int kSlotRows = 144;                 // 128 + GW + 1 halo rows (GW <= 15)
constexpr int kPlaneBytes = kSlotRows * 128;   // one k-block plane of one term: 18 KiB
constexpr int kOC = 32, kC = 16;
constexpr int kBTile = kOC * 128;              // [32 oc x 32 floats]
constexpr int kBBytes = 4 * 2 * kBTile;        // 4 taps x 2 planes, one term
constexpr int kRaw = 2;
constexpr int kThreads = 448;                  // warps 0-7 re-layout, 8-11 epilogue, 12 MMA + TMEM, 13 loader
constexpr int kProducerWarps = 8, kEpiWarp0 = 8, kMmaWarp = 12, kLoadWarp = 13;
constexpr int kTmemCols = 128;                 // 2 buffers x 2 planes x 32 columns

struct Geom {
    int n_img, IH, IW, OH, OW, GH, GW, n_cells, n_tiles;
    uint32_t img_bytes, raw_stage_bytes, div_magic;
};

struct SmemLayout {
    uint32_t b_off, a_off, raw_off, bar_off, total;
    __host__ __device__ explicit SmemLayout(uint32_t raw_stage_bytes) {
        b_off = 0;                                   // [hi | lo] x [tap][plane] tiles
        a_off = 2 * kBBytes;                         // [plane][hi | lo] x 144 rows
        raw_off = a_off + 4 * kPlaneBytes;
        bar_off = raw_off + kRaw * raw_stage_bytes;
        total = bar_off + 256 + 1024;
    }
};

__device__ __forceinline__ void split_store(uint32_t hi_addr, uint32_t lo_addr, const float (&v)[4]) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        h[k] = __float_as_uint(v[k]);
        l[k] = __float_as_uint(v[k] - __uint_as_float(h[k] & 0xFFFFE000u));
    }
    sts128u(hi_addr, h);
    sts128u(lo_addr, l);
}

__global__ void __launch_bounds__(kThreads, 1)
conv2_s2d_fwd_kernel(const float* __restrict__ X, const float* __restrict__ Wg, const float* __restrict__ bias,
                     float* __restrict__ Y, Geom g, int relu) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const SmemLayout L(g.raw_stage_bytes);
    const uint32_t smem_u = smem_u32(smem);
    const uint32_t a_u32 = smem_u + L.a_off, raw_u32 = smem_u + L.raw_off;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bar_off);
    uint64_t* raw_full = bars;                       // [kRaw]
    uint64_t* raw_empty = raw_full + kRaw;           // [kRaw]
    uint64_t* p_full = raw_empty + kRaw;             // [2] plane written -> MMA
    uint64_t* p_empty = p_full + 2;                  // [2] plane consumed (commit) -> producers
    uint64_t* acc_full = p_empty + 2;                // [2]
    uint64_t* acc_empty = acc_full + 2;              // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    constexpr uint32_t kIdesc = make_idesc_tf32(kRows, kOC);

    if (threadIdx.x == 0) {
        for (int s = 0; s < kRaw; ++s) { mbar_init(&raw_full[s], 1); mbar_init(&raw_empty[s], kProducerWarps); }
        for (int p = 0; p < 2; ++p) { mbar_init(&p_full[p], kProducerWarps); mbar_init(&p_empty[p], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kMmaWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    // filter bank -> W4 tiles [tap][plane][oc][32 floats], hi and lo, K-major SWIZZLE_128B; kk = c*4 + dy*2 + dx
    for (int idx = threadIdx.x; idx < 4 * kOC * kC; idx += kThreads) {      // one 16-byte chunk (one channel's 2x2 taps) each
        const int c = idx & 15, oc = (idx >> 4) & 31, tap = idx >> 9;
        const int by = tap >> 1, bx = tap & 1;
        const float* wp = Wg + ((oc * kC + c) * 4 + 2 * by) * 4 + 2 * bx;     // w[oc][c][2by+dy][2bx+dx]
        const float v[4] = {wp[0], wp[1], wp[4], wp[5]};
        const uint32_t off = static_cast<uint32_t>((tap * 2 + (c >> 3)) * kBTile + oc * 128 + (((c & 7) ^ (oc & 7)) << 4));
        split_store(smem_u + L.b_off + off, smem_u + L.b_off + kBBytes + off, v);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = uniform_u32(*tmem_slot);
    const int n_local = (g.n_img - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);

    if (warp == kLoadWarp) {
        // ================================================================ image loader
        if (elect_one()) {
            for (int i = 0; i < n_local; ++i) {
                const int rs = i % kRaw;
                mbar_wait(&raw_empty[rs], ((i / kRaw) & 1) ^ 1);
                const int64_t n = static_cast<int64_t>(blockIdx.x) + static_cast<int64_t>(i) * gridDim.x;
                mbar_expect_tx(&raw_full[rs], g.img_bytes);
                bulk_load(raw_u32 + static_cast<uint32_t>(rs) * g.raw_stage_bytes,
                          reinterpret_cast<const uint8_t*>(X) + n * g.img_bytes, g.img_bytes, &raw_full[rs]);
            }
        }
        __syncwarp();
    } else if (warp < kProducerWarps) {
        // ================================================================ re-layout: NCHW plane -> cell rows (hi, lo)
        // warp w writes channel w of plane 0, then channel w of plane 1 (c = w + 8); lane = cell within a 32-cell
        // group.  A thread reads the 2x2 pixel block of its cell's channel (zero outside the plane) = one 16-byte
        // chunk of the cell's row, and stores it as hi and lo (8 consecutive rows cover the 8 swizzled chunk
        // positions: conflict free).
        const int q_end = kRows + g.GW + 1;                  // rows [0, q_end) of the tile are read by some tap
        uint32_t it = 0;                                     // tile counter
        for (int i = 0; i < n_local; ++i) {
            const int rs = i % kRaw;
            mbar_wait(&raw_full[rs], (i / kRaw) & 1);
            const uint32_t raw_base = raw_u32 + static_cast<uint32_t>(rs) * g.raw_stage_bytes;
            for (int t = 0; t < g.n_tiles; ++t, ++it) {
#pragma unroll
                for (int plane = 0; plane < 2; ++plane) {
                    const int c = warp + 8 * plane;
                    const uint32_t ch_base = raw_base + static_cast<uint32_t>(c * g.IH * g.IW) * 4u;
                    float v[5][4];
#pragma unroll
                    for (int u = 0; u < 5; ++u) {
                        const int q = 32 * u + lane;
                        const int cell = t * kRows + q;
                        const int Yc = static_cast<int>((static_cast<uint32_t>(cell) * g.div_magic) >> 16);
                        const int Xc = cell - Yc * g.GW;
                        const bool in = q < q_end && cell < g.n_cells;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const int iy = 2 * Yc - 1 + (k >> 1), ix = 2 * Xc - 1 + (k & 1);
                            const bool ok = in && iy >= 0 && iy < g.IH && ix >= 0 && ix < g.IW;
                            v[u][k] = ok ? __uint_as_float(lds32(ch_base + static_cast<uint32_t>(iy * g.IW + ix) * 4u)) : 0.0f;
                        }
                    }
                    mbar_wait(&p_empty[plane], (it & 1) ^ 1);
                    const uint32_t hi_base = a_u32 + static_cast<uint32_t>(plane * 2 * kPlaneBytes);
#pragma unroll
                    for (int u = 0; u < 5; ++u) {
                        const int q = 32 * u + lane;
                        if (q < q_end && q < kSlotRows) {
                            const uint32_t off = static_cast<uint32_t>(q * 128 + (((warp & 7) ^ (q & 7)) << 4));
                            split_store(hi_base + off, hi_base + kPlaneBytes + off, v[u]);
                        }
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&p_full[plane]);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&raw_empty[rs]);
        }
    } else if (warp == kMmaWarp) {
        // ================================================================ MMA issuer
        uint32_t it = 0;
        for (int i = 0; i < n_local; ++i) {
            for (int t = 0; t < g.n_tiles; ++t, ++it) {
                const int buf = it & 1;
                mbar_wait(&acc_empty[buf], ((it >> 1) & 1) ^ 1);
#pragma unroll
                for (int plane = 0; plane < 2; ++plane) {
                    mbar_wait(&p_full[plane], it & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (elect_one()) {
                        const uint32_t acc = tmem_base + static_cast<uint32_t>(buf * 2 * kOC + plane * kOC);
                        const uint8_t* a_hi = smem + L.a_off + plane * 2 * kPlaneBytes;
#pragma unroll
                        for (int tap = 0; tap < 4; ++tap) {
                            const int shift = (tap >> 1) * g.GW + (tap & 1);
                            const uint64_t dah = make_desc(a_hi + shift * 128), dal = make_desc(a_hi + kPlaneBytes + shift * 128);
                            const uint64_t dbh = make_desc(smem + L.b_off + (tap * 2 + plane) * kBTile);
                            const uint64_t dbl = make_desc(smem + L.b_off + kBBytes + (tap * 2 + plane) * kBTile);
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const uint64_t adv = static_cast<uint64_t>(2 * k);
                                umma_tf32(acc, dah + adv, dbh + adv, kIdesc, (tap | k) ? 1u : 0u);
                                umma_tf32(acc, dah + adv, dbl + adv, kIdesc, 1u);
                                umma_tf32(acc, dal + adv, dbh + adv, kIdesc, 1u);
                            }
                        }
                        umma_commit(&p_empty[plane]);
                        if (plane == 1) umma_commit(&acc_full[buf]);
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        // ================================================================ epilogue (warps 8..11)
        const int qw = warp - kEpiWarp0;
        const uint32_t lane_base = static_cast<uint32_t>(qw * 32) << 16;
        const int P = g.OH * g.OW;
        float bs[kOC];
#pragma unroll
        for (int oc = 0; oc < kOC; ++oc) bs[oc] = bias[oc];
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
                    tmem_ld32(tmem_base + lane_base + static_cast<uint32_t>(buf * 2 * kOC), r0);
                    tmem_ld32(tmem_base + lane_base + static_cast<uint32_t>(buf * 2 * kOC + kOC), r1);
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
                        float v = __uint_as_float(r0[oc]) + __uint_as_float(r1[oc]) + bs[oc];
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

inline Geom make_geom(int64_t N, int IH, int IW) {
    Geom g;
    g.n_img = static_cast<int>(N); g.IH = IH; g.IW = IW;
    g.OH = (IH + 2 - 4) / 2 + 1; g.OW = (IW + 2 - 4) / 2 + 1;
    g.GH = g.OH + 1; g.GW = g.OW + 1;
    g.n_cells = g.GH * g.GW; g.n_tiles = (g.n_cells + kRows - 1) / kRows;
    g.img_bytes = static_cast<uint32_t>(kC * IH * IW * 4);
    g.raw_stage_bytes = (g.img_bytes + 127u) & ~127u;
    g.div_magic = (65536u + static_cast<uint32_t>(g.GW) - 1u) / static_cast<uint32_t>(g.GW);
    return g;
}

inline bool geom_ok(int C, int IH, int IW) {
    if (!(C == kC && IH >= 2 && IW >= 2)) return false;      // images are 64*IH*IW bytes: always 16-byte multiples
    const Geom g = make_geom(1, IH, IW);
    if (g.GW > 15 || g.n_cells + 160 >= 1024) return false;
    for (uint32_t cell = 0; cell < 1024; ++cell)
        if (((cell * g.div_magic) >> 16) != cell / static_cast<uint32_t>(g.GW)) return false;
    return SmemLayout(g.raw_stage_bytes).total <= 232448u;
}

inline cudaError_t launch_fwd(const float* X, const float* W, const float* bias, float* Y, const Geom& g, int relu, int sms,
                              cudaStream_t st) {
    const SmemLayout L(g.raw_stage_bytes);
    static uint32_t attr_bytes = 0;
    if (L.total > attr_bytes) {
        cudaError_t e = cudaFuncSetAttribute(conv2_s2d_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(L.total));
        if (e != cudaSuccess) return e;
        attr_bytes = L.total;
    }
    const int grid = g.n_img < sms ? g.n_img : sms;
    conv2_s2d_fwd_kernel<<<static_cast<unsigned>(grid), kThreads, L.total, st>>>(X, W, bias, Y, g, relu);
    return cudaGetLastError();
}


// ====================================================================================================
// Input gradient of the same layer (ConvolutionBackward wrt input) in the same cell space:
//     dcell(Y,X)[(c,dy,dx)] = sum_{by,bx} sum_oc g[oc, Y-by, X-bx] * w[oc, c, 2by+dy, 2bx+dx]
// (g = the ReLU-masked output gradient, zero outside [0,OH) x [0,OW)): one GEMM row per INPUT cell, K = 32 output
// channels per tap, N = 64 = the cell's (c,dy,dx) values.  The A operand is the gradient itself as cell-grid rows
// [r][32 oc] (one 128-byte K-major row per position, zero rows for X = OW / Y = OH / before the image), staged
// GW+1 rows down so that the four taps are the four row-shifted views (GW+1) - (by*GW + bx) of the same tile.
// The v1 kernel ran four parity-class GEMMs with 4-byte gathers (600 us per 8192-sample minibatch, 0.08 of its
// HBM floor).  Here the gradient image streams in by cp.async.bulk, the input-gradient image is assembled in
// shared memory and leaves by one bulk store (every pixel belongs to exactly one cell: no read-modify-write,
// fully coalesced).
namespace dg {

constexpr int kN = 64;                         // (c, dy, dx)
constexpr int kBTileD = kN * 128;              // one tap: [64 rows x 32 oc floats], 8 KiB
constexpr int kBBytesD = 4 * kBTileD;          // one term
constexpr int kAPlane = kSlotRows * 128;       // [144 rows x 32 oc floats]
constexpr int kASlots = 2, kGRaw = 2, kOStages = 2;
constexpr int kTmemColsD = 256;                // 2 buffers x 2 accumulators (by = 0 | 1) x 64 columns

struct SmemLayout {
    uint32_t b_off, a_off, g_off, o_off, bar_off, total, g_stage_bytes, o_stage_bytes;
    __host__ __device__ SmemLayout(uint32_t g_bytes, uint32_t o_bytes) {
        g_stage_bytes = (g_bytes + 127u) & ~127u;
        o_stage_bytes = (o_bytes + 127u) & ~127u;
        b_off = 0;                                   // [hi | lo] x [tap]
        a_off = 2 * kBBytesD;                        // [slot][hi | lo]
        g_off = a_off + kASlots * 2 * kAPlane;
        o_off = g_off + kGRaw * g_stage_bytes;
        bar_off = o_off + kOStages * o_stage_bytes;
        total = bar_off + 256 + 1024;
    }
};

__device__ __forceinline__ void bulk_store(void* dst_global, uint32_t src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(reinterpret_cast<uint64_t>(dst_global)),
                 "r"(src_smem), "r"(bytes) : "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
conv2_s2d_dgrad_kernel(const float* __restrict__ G, const float* __restrict__ Wg, float* __restrict__ dX, Geom g,
                       unsigned int* __restrict__ chan_absmax) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int P = g.OH * g.OW;
    const uint32_t g_bytes = static_cast<uint32_t>(kOC * P * 4);
    const SmemLayout L(g_bytes, g.img_bytes);
    const uint32_t smem_u = smem_u32(smem);
    const uint32_t a_u32 = smem_u + L.a_off, g_u32 = smem_u + L.g_off, o_u32 = smem_u + L.o_off;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bar_off);
    uint64_t* g_full = bars;                         // [kGRaw]
    uint64_t* g_empty = g_full + kGRaw;
    uint64_t* a_full = g_empty + kGRaw;              // [kASlots]
    uint64_t* a_empty = a_full + kASlots;
    uint64_t* acc_full = a_empty + kASlots;          // [2]
    uint64_t* acc_empty = acc_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    constexpr uint32_t kIdesc = make_idesc_tf32(kRows, kN);
    const int halo = g.GW + 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kGRaw; ++s) { mbar_init(&g_full[s], 1); mbar_init(&g_empty[s], kProducerWarps); }
        for (int s = 0; s < kASlots; ++s) { mbar_init(&a_full[s], kProducerWarps); mbar_init(&a_empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kMmaWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(kTmemColsD));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    // filter bank -> [tap][n = c*4 + dy*2 + dx][k = oc], hi and lo; one 16-byte chunk = 4 consecutive oc
    for (int idx = threadIdx.x; idx < 4 * kN * 8; idx += kThreads) {
        const int j = idx & 7, n = (idx >> 3) & 63, tap = idx >> 9;
        const int by = tap >> 1, bx = tap & 1, c = n >> 2, dy = (n >> 1) & 1, dx = n & 1;
        const float* wp = Wg + ((4 * j * kC + c) * 4 + 2 * by + dy) * 4 + 2 * bx + dx;     // w[4j + i][c][2by+dy][2bx+dx]
        const float v[4] = {wp[0], wp[kC * 16], wp[2 * kC * 16], wp[3 * kC * 16]};
        const uint32_t off = static_cast<uint32_t>(tap * kBTileD + n * 128 + ((j ^ (n & 7)) << 4));
        split_store(smem_u + L.b_off + off, smem_u + L.b_off + kBBytesD + off, v);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = uniform_u32(*tmem_slot);
    const int n_local = (g.n_img - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);

    if (warp == kLoadWarp) {
        // ================================================================ gradient image loader
        if (elect_one()) {
            for (int i = 0; i < n_local; ++i) {
                const int s = i % kGRaw;
                mbar_wait(&g_empty[s], ((i / kGRaw) & 1) ^ 1);
                const int64_t n = static_cast<int64_t>(blockIdx.x) + static_cast<int64_t>(i) * gridDim.x;
                mbar_expect_tx(&g_full[s], g_bytes);
                bulk_load(g_u32 + static_cast<uint32_t>(s) * L.g_stage_bytes, G + n * (static_cast<int64_t>(kOC) * P), g_bytes,
                          &g_full[s]);
            }
        }
        __syncwarp();
    } else if (warp < kProducerWarps) {
        // ================================================================ gradient [oc][pos] -> cell-grid rows [r][oc] (hi, lo)
        // warp = 16-byte chunk j (oc 4j .. 4j+3), lane = slot row within a 32-row group; slot row s of tile t is the
        // grid position 128 t + s - (GW+1).
        const int j = warp;
        uint32_t it = 0;
        for (int i = 0; i < n_local; ++i) {
            const int gs = i % kGRaw;
            mbar_wait(&g_full[gs], (i / kGRaw) & 1);
            const uint32_t g_base = g_u32 + static_cast<uint32_t>(gs) * L.g_stage_bytes + static_cast<uint32_t>(4 * j * P) * 4u;
            for (int t = 0; t < g.n_tiles; ++t, ++it) {
                const int slot = it % kASlots;
                float v[5][4];
#pragma unroll
                for (int u = 0; u < 5; ++u) {
                    const int srow = 32 * u + lane;
                    const int cell = t * kRows + srow - halo;
                    const int cc = cell < 0 ? 0 : cell;
                    const int Yc = static_cast<int>((static_cast<uint32_t>(cc) * g.div_magic) >> 16);
                    const int Xc = cc - Yc * g.GW;
                    const bool ok = srow < kSlotRows && cell >= 0 && Yc < g.OH && Xc < g.OW;
                    const uint32_t src = g_base + static_cast<uint32_t>(Yc * g.OW + Xc) * 4u;
#pragma unroll
                    for (int k = 0; k < 4; ++k) v[u][k] = ok ? __uint_as_float(lds32(src + static_cast<uint32_t>(k * P) * 4u)) : 0.0f;
                }
                mbar_wait(&a_empty[slot], ((it / kASlots) & 1) ^ 1);
                const uint32_t hi_base = a_u32 + static_cast<uint32_t>(slot * 2 * kAPlane);
#pragma unroll
                for (int u = 0; u < 5; ++u) {
                    const int srow = 32 * u + lane;
                    if (srow < kSlotRows) {
                        const uint32_t off = static_cast<uint32_t>(srow * 128 + ((j ^ (srow & 7)) << 4));
                        split_store(hi_base + off, hi_base + kAPlane + off, v[u]);
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&a_full[slot]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&g_empty[gs]);
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
                    const uint8_t* a_hi = smem + L.a_off + slot * 2 * kAPlane;
#pragma unroll
                    for (int tap = 0; tap < 4; ++tap) {
                        const int row0 = halo - ((tap >> 1) * g.GW + (tap & 1));
                        const uint32_t acc = tmem_base + static_cast<uint32_t>(buf * 2 * kN + (tap >> 1) * kN);
                        const uint64_t dah = make_desc(a_hi + row0 * 128), dal = make_desc(a_hi + kAPlane + row0 * 128);
                        const uint64_t dbh = make_desc(smem + L.b_off + tap * kBTileD);
                        const uint64_t dbl = make_desc(smem + L.b_off + kBBytesD + tap * kBTileD);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint64_t adv = static_cast<uint64_t>(2 * k);
                            umma_tf32(acc, dah + adv, dbh + adv, kIdesc, ((tap & 1) | k) ? 1u : 0u);
                            umma_tf32(acc, dah + adv, dbl + adv, kIdesc, 1u);
                            umma_tf32(acc, dal + adv, dbh + adv, kIdesc, 1u);
                        }
                    }
                    umma_commit(&a_empty[slot]);
                    umma_commit(&acc_full[buf]);
                }
                __syncwarp();
            }
        }
    } else {
        // ================================================================ epilogue (warps 8..11): TMEM -> image in smem -> bulk store
        const int qw = warp - kEpiWarp0;
        const uint32_t lane_base = static_cast<uint32_t>(qw * 32) << 16;
        const int plane = g.IH * g.IW;
        // chan_absmax (optional): max |dX[:, c]| per input channel over the whole call, as the bit pattern of a
        // non-negative float - the scale the first layer's kind::i8 weight gradient quantises this gradient against
        // (conv1_i8.cuh wg: S_c = 2^e > max |g[:, c]|), produced here for free instead of by a 43 us pass over dX.
        float cmax[kC];
#pragma unroll
        for (int c = 0; c < kC; ++c) cmax[c] = 0.0f;
        uint32_t it = 0;
        for (int i = 0; i < n_local; ++i) {
            const int64_t n = static_cast<int64_t>(blockIdx.x) + static_cast<int64_t>(i) * gridDim.x;
            const uint32_t o_base = o_u32 + static_cast<uint32_t>(i & 1) * L.o_stage_bytes;
            for (int t = 0; t < g.n_tiles; ++t, ++it) {
                const int buf = it & 1;
                mbar_wait(&acc_full[buf], (it >> 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const int cell = t * kRows + qw * 32 + lane;
                const bool warp_live = t * kRows + qw * 32 < g.n_cells;       // warp-uniform
                const int Yc = static_cast<int>((static_cast<uint32_t>(cell) * g.div_magic) >> 16);
                const int Xc = cell - Yc * g.GW;
                const bool row_ok = cell < g.n_cells;
#pragma unroll
                for (int h = 0; h < 2; ++h) {                                  // columns [32h, 32h+32) = channels 8h .. 8h+7
                    uint32_t r0[32], r1[32];
                    if (warp_live) {
                        tmem_ld32(tmem_base + lane_base + static_cast<uint32_t>(buf * 2 * kN + h * 32), r0);
                        tmem_ld32(tmem_base + lane_base + static_cast<uint32_t>(buf * 2 * kN + kN + h * 32), r1);
                    }
                    if (warp_live && row_ok) {
#pragma unroll
                        for (int col = 0; col < 32; ++col) {
                            const int nn = h * 32 + col, c = nn >> 2, dy = (nn >> 1) & 1, dx = nn & 1;
                            const int iy = 2 * Yc - 1 + dy, ix = 2 * Xc - 1 + dx;
                            if (iy >= 0 && iy < g.IH && ix >= 0 && ix < g.IW) {
                                const float v = __uint_as_float(r0[col]) + __uint_as_float(r1[col]);
                                sts32u(o_base + static_cast<uint32_t>(c * plane + iy * g.IW + ix) * 4u, __float_as_uint(v));
                                cmax[c] = fmaxf(cmax[c], fabsf(v));
                            }
                        }
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[buf]);
            }
            // the image is complete: make the generic-proxy writes visible to the bulk engine, then one thread stores it
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (warp == kEpiWarp0 && lane == 0) {
                bulk_store(reinterpret_cast<uint8_t*>(dX) + n * g.img_bytes, o_base, g.img_bytes);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // the other stage's store has been read out
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
        }
        if (warp == kEpiWarp0 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        if (chan_absmax != nullptr) {
#pragma unroll
            for (int c = 0; c < kC; ++c) {
                float m = cmax[c];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
                if (lane == 0 && m > 0.0f) atomicMax(chan_absmax + c, __float_as_uint(m));
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == kMmaWarp)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemColsD));
}

inline bool smem_ok(const Geom& g) {
    return SmemLayout(static_cast<uint32_t>(kOC * g.OH * g.OW * 4), g.img_bytes).total <= 232448u;
}

inline cudaError_t launch_dgrad(const float* G, const float* W, float* dX, const Geom& g, int sms, cudaStream_t st,
                                float* chan_absmax = nullptr) {
    const SmemLayout L(static_cast<uint32_t>(kOC * g.OH * g.OW * 4), g.img_bytes);
    static uint32_t attr_bytes = 0;
    if (L.total > attr_bytes) {
        cudaError_t e = cudaFuncSetAttribute(conv2_s2d_dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(L.total));
        if (e != cudaSuccess) return e;
        attr_bytes = L.total;
    }
    const int grid = g.n_img < sms ? g.n_img : sms;
    if (chan_absmax != nullptr) {
        cudaError_t e = cudaMemsetAsync(chan_absmax, 0, kC * sizeof(float), st);
        if (e != cudaSuccess) return e;
    }
    conv2_s2d_dgrad_kernel<<<static_cast<unsigned>(grid), kThreads, L.total, st>>>(G, W, dX, g, reinterpret_cast<unsigned int*>(chan_absmax));
    return cudaGetLastError();
}

}  // namespace dg


// ====================================================================================================
// Weight + bias gradient of the same layer in cell space:
//     dW[oc, c, 2by+dy, 2bx+dx] = sum_{n, cells r} g[n][r][oc] * cell_n[r + by*GW + bx][(c,dy,dx)]
// (g = the ReLU-masked output gradient on the cell grid, zero for the dropped cells).  Both operands are read
// MN-major with K = cells: A = the cell rows [r][64 floats] the forward kernel stages (two 128-byte planes = the
// two 32-element M atoms of an M = 64 operand, LBO = the plane stride), shifted in the K direction per tap;
// B = the gradient rows [r][32 oc] the input-gradient kernel stages.  D_tap = [64 x 32] per tap in TMEM (an M = 64
// accumulator lives in lanes 0-15 / 32-47 / 64-79 / 96-111, tcgen05_shift_probe).  The tensor-core accumulator
// truncates, so every image's partial (K = 128 cells) is promoted into fp32 registers by eight accumulator warps
// (two TMEM buffers: the drain of image i overlaps the MMAs of image i+1); per-CTA partial sums are reduced over
// CTAs in a fixed order (deterministic).  3-term TF32 split.
// The two bx taps of a by row share ONE MMA: sum_r A[r+s+1] G[r] = sum_r' A[r'+s] G[r'-1], so with the gradient rows
// staged 4 rows down (rows 0-3 of the plane stay zero) the B operand is N = 64 = two overlapping 32-column atoms,
// atom 0 = the plane from row 3 (G[r'-1]: bx = 1), atom 1 = LBO = 128 bytes further (G[r']: bx = 0).  96 MMAs
// (M=64, N=64, K=8) per image instead of 192 with N=32 - the kernel was bound by the tensor pipe's per-instruction
// cost at small N (ncu: pipe 65 % busy at 240 us).
namespace wg2 {

constexpr int kThreadsW = 576;                 // warps 0-7 re-layout, 8-15 accumulators, 16 MMA + TMEM, 17 loader
constexpr int kAccWarp0 = 8, kMmaWarpW = 16, kLoadWarpW = 17;
constexpr int kGPlane = kSlotRows * 128;       // gradient rows [144][32 oc], one term
constexpr int kTmemColsW = 256;                // 2 buffers x 4 taps x 32 columns

struct SmemLayout {
    uint32_t a_off, g_off, x_off, gr_off, bar_off, total, g_stage_bytes;
    __host__ __device__ SmemLayout(uint32_t raw_stage_bytes, uint32_t g_bytes) {
        g_stage_bytes = (g_bytes + 127u) & ~127u;
        a_off = 0;                                   // [plane][hi | lo] x 144 rows (as in the forward kernel)
        g_off = a_off + 4 * kPlaneBytes;             // [hi | lo] x 144 rows
        x_off = g_off + 2 * kGPlane;
        gr_off = x_off + kRaw * raw_stage_bytes;
        bar_off = gr_off + kRaw * g_stage_bytes;
        total = bar_off + 256 + 1024;
    }
};

// MN-major descriptor for 32-bit operands: layout SWIZZLE_128B_BASE32B (cute::UMMA::LayoutType 1, Swizzle<2,5,2>:
// the 32-byte unit index of a 128-byte row is XORed with the row index mod 4; a K atom is 4 rows = 512 bytes, SBO).
// The ordinary SWIZZLE_128B pattern (16-byte units, row mod 8) is only valid MN-major for 8/16-bit types - with it a
// kind::tf32 MMA returns zeros (tools/probes/conv2_s2d_probe.cu: mn_major_probes).  `lbo_bytes` = distance between
// the 32-element (128-byte) atoms along M/N.
__device__ __forceinline__ uint64_t make_desc_mn(const void* smem_tile, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_u32(smem_tile) & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
    d |= static_cast<uint64_t>(512 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(1) << 61;
    return d;
}
// byte offset of the 16-byte chunk j (4 floats) of row q in that layout
__device__ __forceinline__ uint32_t mn32_chunk_off(int q, int j) {
    return static_cast<uint32_t>(q * 128 + ((((j >> 1) ^ (q & 3)) << 5) | ((j & 1) << 4)));
}

__global__ void __launch_bounds__(kThreadsW, 1)
conv2_s2d_wgrad_kernel(const float* __restrict__ X, const float* __restrict__ G, float* __restrict__ partial,
                       float* __restrict__ partial_bias, Geom g) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int P = g.OH * g.OW;
    const uint32_t g_bytes = static_cast<uint32_t>(kOC * P * 4);
    const SmemLayout L(g.raw_stage_bytes, g_bytes);
    const uint32_t smem_u = smem_u32(smem);
    const uint32_t a_u32 = smem_u + L.a_off, gt_u32 = smem_u + L.g_off, x_u32 = smem_u + L.x_off, gr_u32 = smem_u + L.gr_off;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bar_off);
    uint64_t* x_full = bars;                         // [kRaw]
    uint64_t* x_empty = x_full + kRaw;
    uint64_t* g_full = x_empty + kRaw;               // [kRaw]
    uint64_t* g_empty = g_full + kRaw;
    uint64_t* op_full = g_empty + kRaw;              // [1] operand tiles written -> MMA
    uint64_t* op_empty = op_full + 1;                // [1] MMAs done reading
    uint64_t* acc_full = op_empty + 1;               // [2]
    uint64_t* acc_empty = acc_full + 2;              // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    // A and B MN-major (bits 15, 16), M = 64, N = 64 (two bx taps x 32 oc)
    constexpr uint32_t kIdesc = make_idesc_tf32(64, 2 * kOC) | (1u << 15) | (1u << 16);
    constexpr int kGRow0 = 4;                        // gradient row r is staged at plane row r + 4; rows 0-3 stay zero

    if (threadIdx.x == 0) {
        for (int s = 0; s < kRaw; ++s) {
            mbar_init(&x_full[s], 1); mbar_init(&x_empty[s], kProducerWarps);
            mbar_init(&g_full[s], 1); mbar_init(&g_empty[s], kProducerWarps);
        }
        mbar_init(op_full, kProducerWarps);
        mbar_init(op_empty, 1);
        for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kMmaWarpW) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(kTmemColsW));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = uniform_u32(*tmem_slot);
    const int n_local = (g.n_img - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);

    if (warp == kLoadWarpW) {
        // ================================================================ loader: activation and gradient images
        if (elect_one()) {
            for (int i = 0; i < n_local; ++i) {
                const int s = i % kRaw;
                const uint32_t ph = ((i / kRaw) & 1) ^ 1;
                const int64_t n = static_cast<int64_t>(blockIdx.x) + static_cast<int64_t>(i) * gridDim.x;
                mbar_wait(&x_empty[s], ph);
                mbar_expect_tx(&x_full[s], g.img_bytes);
                bulk_load(x_u32 + static_cast<uint32_t>(s) * g.raw_stage_bytes, reinterpret_cast<const uint8_t*>(X) + n * g.img_bytes,
                          g.img_bytes, &x_full[s]);
                mbar_wait(&g_empty[s], ph);
                mbar_expect_tx(&g_full[s], g_bytes);
                bulk_load(gr_u32 + static_cast<uint32_t>(s) * L.g_stage_bytes, G + n * (static_cast<int64_t>(kOC) * P), g_bytes, &g_full[s]);
            }
        }
        __syncwarp();
    } else if (warp < kProducerWarps) {
        // ================================================================ re-layout (the forward's cell rows + the dgrad's gradient rows)
        const int q_end = kRows + g.GW + 1;
        float bias_acc[4] = {0.f, 0.f, 0.f, 0.f};            // channels 4*warp .. 4*warp+3, this lane's rows
        if (warp == 0) {                                     // the zero rows in front of the gradient planes (hi, lo), once
            const uint32_t z[4] = {0u, 0u, 0u, 0u};
            sts128u(gt_u32 + static_cast<uint32_t>(lane) * 16u, z);
            sts128u(gt_u32 + kGPlane + static_cast<uint32_t>(lane) * 16u, z);
        }
        uint32_t it = 0;
        for (int i = 0; i < n_local; ++i) {
            const int s = i % kRaw;
            mbar_wait(&x_full[s], (i / kRaw) & 1);
            mbar_wait(&g_full[s], (i / kRaw) & 1);
            const uint32_t raw_base = x_u32 + static_cast<uint32_t>(s) * g.raw_stage_bytes;
            const uint32_t g_base = gr_u32 + static_cast<uint32_t>(s) * L.g_stage_bytes + static_cast<uint32_t>(4 * warp * P) * 4u;
            for (int t = 0; t < g.n_tiles; ++t, ++it) {
                // ---- loads of both operands first (activations: channel warp of each plane; gradient: chunk warp)
                float va[2][5][4], vg[5][4];
#pragma unroll
                for (int u = 0; u < 5; ++u) {
                    const int q = 32 * u + lane;
                    const int cell = t * kRows + q;
                    const int Yc = static_cast<int>((static_cast<uint32_t>(cell) * g.div_magic) >> 16);
                    const int Xc = cell - Yc * g.GW;
                    const bool in = q < q_end && cell < g.n_cells;
#pragma unroll
                    for (int plane = 0; plane < 2; ++plane) {
                        const uint32_t ch_base = raw_base + static_cast<uint32_t>((warp + 8 * plane) * g.IH * g.IW) * 4u;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const int iy = 2 * Yc - 1 + (k >> 1), ix = 2 * Xc - 1 + (k & 1);
                            const bool ok = in && iy >= 0 && iy < g.IH && ix >= 0 && ix < g.IW;
                            va[plane][u][k] = ok ? __uint_as_float(lds32(ch_base + static_cast<uint32_t>(iy * g.IW + ix) * 4u)) : 0.0f;
                        }
                    }
                    const bool gok = q < kRows && cell < g.n_cells && Yc < g.OH && Xc < g.OW;
                    const uint32_t src = g_base + static_cast<uint32_t>(Yc * g.OW + Xc) * 4u;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        vg[u][k] = gok ? __uint_as_float(lds32(src + static_cast<uint32_t>(k * P) * 4u)) : 0.0f;
                        bias_acc[k] += vg[u][k];
                    }
                }
                mbar_wait(op_empty, (it & 1) ^ 1);
#pragma unroll
                for (int u = 0; u < 5; ++u) {
                    const int q = 32 * u + lane;
                    if (q < q_end && q < kSlotRows) {
                        const uint32_t off = mn32_chunk_off(q, warp);
#pragma unroll
                        for (int plane = 0; plane < 2; ++plane) {
                            const uint32_t hi = a_u32 + static_cast<uint32_t>(plane * 2 * kPlaneBytes) + off;
                            split_store(hi, hi + kPlaneBytes, va[plane][u]);
                        }
                        if (q < kRows) {
                            const uint32_t goff = mn32_chunk_off(q + kGRow0, warp);
                            split_store(gt_u32 + goff, gt_u32 + kGPlane + goff, vg[u]);
                        }
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(op_full);
            }
            __syncwarp();
            if (lane == 0) { mbar_arrive(&x_empty[s]); mbar_arrive(&g_empty[s]); }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float sum = warp_sum(bias_acc[k]);
            if (lane == 0) partial_bias[static_cast<int64_t>(blockIdx.x) * kOC + 4 * warp + k] = sum;
        }
    } else if (warp == kMmaWarpW) {
        // ================================================================ MMA issuer
        uint32_t it = 0;
        for (int i = 0; i < n_local; ++i) {
            for (int t = 0; t < g.n_tiles; ++t, ++it) {
                const int buf = it & 1;
                mbar_wait(&acc_empty[buf], ((it >> 1) & 1) ^ 1);
                mbar_wait(op_full, it & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (elect_one()) {
                    const uint8_t* a_hi = smem + L.a_off;            // plane 0 hi; plane 1 hi is 2*kPlaneBytes further (= LBO)
                    const uint8_t* g_hi = smem + L.g_off;
#pragma unroll
                    for (int by = 0; by < 2; ++by) {
                        const int shift = by * g.GW;                  // bx rides on the B operand's two atoms
                        const uint32_t acc = tmem_base + static_cast<uint32_t>(buf * 4 * kOC + by * 2 * kOC);
                        const uint64_t dah = make_desc_mn(a_hi + shift * 128, 2 * kPlaneBytes);
                        const uint64_t dal = make_desc_mn(a_hi + kPlaneBytes + shift * 128, 2 * kPlaneBytes);
                        // atom 0 = rows from kGRow0 - 1 (G[r-1] -> bx = 1), atom 1 = 128 bytes further (G[r] -> bx = 0)
                        const uint64_t dgh = make_desc_mn(g_hi + (kGRow0 - 1) * 128, 128);
                        const uint64_t dgl = make_desc_mn(g_hi + kGPlane + (kGRow0 - 1) * 128, 128);
#pragma unroll 4
                        for (int k = 0; k < 16; ++k) {                // K = 8 cells per MMA: one 1024-byte atom of rows
                            const uint64_t adv = static_cast<uint64_t>(64 * k);
                            umma_tf32(acc, dah + adv, dgh + adv, kIdesc, k ? 1u : 0u);
                            umma_tf32(acc, dah + adv, dgl + adv, kIdesc, 1u);
                            umma_tf32(acc, dal + adv, dgh + adv, kIdesc, 1u);
                        }
                    }
                    umma_commit(op_empty);
                    umma_commit(&acc_full[buf]);
                }
                __syncwarp();
            }
        }
    } else {
        // ================================================================ accumulator warps 8..15
        // warp a: TMEM lane quarter a & 3 (valid rows: lanes 0..15 = M rows 16*(a&3) + lane), taps 2*(a>>2), 2*(a>>2)+1
        const int a = warp - kAccWarp0;
        const int quarter = a & 3, half = a >> 2;
        const uint32_t lane_base = static_cast<uint32_t>(quarter * 32) << 16;
        float acc[2][kOC];
#pragma unroll
        for (int tp = 0; tp < 2; ++tp)
#pragma unroll
            for (int j = 0; j < kOC; ++j) acc[tp][j] = 0.0f;
        uint32_t it = 0;
        for (int i = 0; i < n_local; ++i) {
            for (int t = 0; t < g.n_tiles; ++t, ++it) {
                const int buf = it & 1;
                mbar_wait(&acc_full[buf], (it >> 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                for (int tp = 0; tp < 2; ++tp) {
                    uint32_t r[32];
                    tmem_ld32(tmem_base + lane_base + static_cast<uint32_t>(buf * 4 * kOC + (2 * half + tp) * kOC), r);
#pragma unroll
                    for (int j = 0; j < kOC; ++j) acc[tp][j] += __uint_as_float(r[j]);
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[buf]);
            }
        }
        if (lane < 16) {                                    // partial[cta][tap][m = (c,dy,dx)][oc]
            const int m = 16 * quarter + lane;
#pragma unroll
            for (int tp = 0; tp < 2; ++tp) {
                // column block 2*half + tp of the accumulators = tap (by = half, bx = 1 - tp): atom 0 of B is the bx = 1 tap
                float4* dst = reinterpret_cast<float4*>(partial + ((static_cast<int64_t>(blockIdx.x) * 4 + 2 * half + (1 - tp)) * 64 + m) * kOC);
#pragma unroll
                for (int j = 0; j < kOC / 4; ++j)
                    dst[j] = make_float4(acc[tp][4 * j], acc[tp][4 * j + 1], acc[tp][4 * j + 2], acc[tp][4 * j + 3]);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == kMmaWarpW)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemColsW));
}

// dW[oc][c][2by+dy][2bx+dx] = sum_cta partial[cta][tap][m][oc] (fp64, fixed order); db[oc] = sum_cta partial_bias
__global__ void __launch_bounds__(256)
wgrad2_reduce_kernel(const float* __restrict__ partial, const float* __restrict__ partial_bias, int n_cta, float* __restrict__ dW,
                     float* __restrict__ db) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;     // 4 * 64 * 32 = 8192 outputs, oc fastest
    if (tid >= 4 * 64 * kOC) return;
    const int oc = tid & 31, m = (tid >> 5) & 63, tap = tid >> 11;
    double acc = 0.0;
    for (int cta = 0; cta < n_cta; ++cta) acc += static_cast<double>(partial[((static_cast<int64_t>(cta) * 4 + tap) * 64 + m) * kOC + oc]);
    const int by = tap >> 1, bx = tap & 1, c = m >> 2, dy = (m >> 1) & 1, dx = m & 1;
    dW[((oc * kC + c) * 4 + 2 * by + dy) * 4 + 2 * bx + dx] = static_cast<float>(acc);
    if (db != nullptr && tid < kOC) {
        double b = 0.0;
        for (int cta = 0; cta < n_cta; ++cta) b += static_cast<double>(partial_bias[cta * kOC + tid]);
        db[tid] = static_cast<float>(b);
    }
}

inline size_t scratch_bytes(int sms) { return static_cast<size_t>(sms) * (4 * 64 * kOC + kOC) * sizeof(float); }
inline bool smem_ok(const Geom& g) {
    return SmemLayout(g.raw_stage_bytes, static_cast<uint32_t>(kOC * g.OH * g.OW * 4)).total <= 232448u;
}

inline cudaError_t launch_wgrad(const float* X, const float* G, float* dW, float* db, const Geom& g, int sms, void* scratch,
                                cudaStream_t st) {
    const SmemLayout L(g.raw_stage_bytes, static_cast<uint32_t>(kOC * g.OH * g.OW * 4));
    static uint32_t attr_bytes = 0;
    if (L.total > attr_bytes) {
        cudaError_t e = cudaFuncSetAttribute(conv2_s2d_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(L.total));
        if (e != cudaSuccess) return e;
        attr_bytes = L.total;
    }
    const int grid = g.n_img < sms ? g.n_img : sms;
    float* partial = static_cast<float*>(scratch);
    float* partial_bias = partial + static_cast<size_t>(grid) * 4 * 64 * kOC;
    conv2_s2d_wgrad_kernel<<<static_cast<unsigned>(grid), kThreadsW, L.total, st>>>(X, G, partial, partial_bias, g);
    wgrad2_reduce_kernel<<<(4 * 64 * kOC + 255) / 256, 256, 0, st>>>(partial, partial_bias, grid, dW, db);
    return cudaGetLastError();
}

}  // namespace wg2

}  // namespace c2s
}  // namespace rl
