// Convolution forward as an IMPLICIT GEMM on the 5th-gen tensor cores (tcgen05, kind::tf32 with the
// fp32-accurate hi/lo split of gemm_tf32x3.cu) for the two conv layers of the AtariFf network.
//
// Reference: rlpyt/models/conv2d.py:36-44 with the defaults of rlpyt/models/pg/atari_ff_model.py:31-35:
//   layer 1: uint8 frames [N,4,H,W] -> *1/255 -> Conv2d(4->16, k8, s4, p0) + ReLU
//   layer 2: fp32 [N,16,IH,IW]      ->           Conv2d(16->32, k4, s2, p1) + ReLU
// Both are GEMMs with K = C*KH*KW = 256:  Y[m, oc] = sum_k A[m, k] * W[oc, k], one row m per output
// position (n, oy, ox), A = im2col(X) never materialised.
//
//   * A producer (8 warps, two threads per output position of the 128-row tile, each owning 4 of the 8
//     chunks; the gathers of k-block i+1 are issued before k-block i is written, so two round trips
//     to L2/HBM are always in flight - the first version, one thread per row without prefetch, spent
//     90 % of its issue slots stalled on the scoreboard, profiles/r01_conv_tc_first.txt): for each
//     32-wide k-block it gathers chunks of 4 consecutive taps (4 pixels of one filter row; coalesced across
//     the warp because neighbouring lanes are neighbouring positions), splits them into TF32 hi/lo
//     and writes them straight into the K-major SWIZZLE_128B layout the MMA expects
//     (chunk j of row r at r*128 + ((j ^ (r & 7)) << 4)) in a 4- / 8-stage ring; layer 1 keeps the pixels as
//     the integers 0..255 - exact in TF32, so A needs no lo term - and applies 1/255 in the epilogue.
//   * B (the filter bank, [oc][256] = the weight tensor as stored) is split and swizzled into shared
//     memory once per persistent CTA.
//   * MMA warp: per k-block 4 x (2|3) tcgen05.mma (M=128, N=16|32, K=8); the two halves of K go to
//     separate TMEM accumulators that the epilogue adds in fp32 (halves the accumulator truncation),
//     and the accumulators are double buffered so the epilogue of tile i overlaps the MMAs of tile i+1.
//   * epilogue (4 warps): tcgen05.ld -> (+bias, *scale, ReLU) -> NCHW stores, coalesced over positions.
// The same kernel also computes the INPUT GRADIENT of layer 2 (policy Dgrad2): the transposed
// convolution is split by the parity (iy&1, ix&1) of the input pixel - each parity class sees a fixed
// 2x2 subset of the 4x4 taps - into four dense GEMMs with K = 32 channels x 4 taps = 128 whose rows are
// the input pixels of that parity, A gathers the ReLU-masked output gradient, B_p[c][oc*4+t] is the
// matching slice of the filter bank (rl_conv2_dgrad_tc prepares the four B_p).
// Measured per 8192-sample minibatch: layer-2 forward 0.33 ms (cuDNN fp32 0.68 ms), layer-1 forward
// 0.61 ms (fp32 SIMT kernel of conv1.cu 0.82 ms; cuDNN incl. the u8->f32 conversion 3.2 ms).
#include "tc_common.cuh"

namespace rl {
namespace convtc {

using namespace tc;

constexpr int kRows = 128;                 // GEMM M tile = output positions per tile
constexpr int kTileBytes = kRows * 128;    // 16 KiB: 128 rows x 128 B
constexpr int kProducerThreads = 256, kEpilogueThreads = 128;
constexpr int kThreads = 416;              // warps 0-7 producers, 8-11 epilogue, 12 MMA + TMEM alloc
constexpr int kMmaWarp = 12, kEpiWarp0 = 8;
constexpr int kTmemCols = 128;

struct Geom {
    int n_img, C, H, W, OH, OW, P;         // P = OH*OW  (forward: rows per image)
    int64_t m_total;                       // forward: n_img * P;  dgrad: rows per parity class
    int BH, BW;                            // dgrad: 2x2 input blocks per image
};

struct RowCtx {                            // one GEMM row = one output position (or input pixel for dgrad)
    const void* xn;                        // image / gradient plane base of this row's sample
    int64_t n;
    int oy, ox;                            // forward: output position;  dgrad: block (a, b)
    int py, px;                            // dgrad: parity of the input pixel
    bool valid;
};

// forward layers: tiles run over n_img*P rows, one B set
template <class L>
struct FwdRows {
    static constexpr int kBSets = 1;
    __device__ static int64_t num_tiles(const Geom& g) { return (g.m_total + kRows - 1) / kRows; }
    __device__ static int bset(const Geom&, int64_t) { return 0; }
    __device__ static RowCtx decode(const typename L::In* X, const int64_t* rows, const Geom& g, int64_t tile, int r) {
        RowCtx c;
        const int64_t m = tile * kRows + r;
        c.valid = m < g.m_total;
        c.n = c.valid ? m / g.P : 0;
        const int pos = c.valid ? static_cast<int>(m - c.n * g.P) : 0;
        c.oy = pos / g.OW;
        c.ox = pos - c.oy * g.OW;
        c.py = c.px = 0;
        const int64_t img = rows != nullptr ? rows[c.n] : c.n;
        c.xn = X + img * (static_cast<int64_t>(g.C) * g.H * g.W);
        return c;
    }
    __device__ static void store(const RowCtx& c, const Geom& g, float* __restrict__ Y, const float* __restrict__ bias,
                                 const float (&acc)[L::kN], int relu) {
        float* yb = Y + c.n * (static_cast<int64_t>(L::kN) * g.P) + c.oy * g.OW + c.ox;
#pragma unroll
        for (int oc = 0; oc < L::kN; ++oc) {
            float v = acc[oc] * L::kScale + bias[oc];
            if (relu) v = fmaxf(v, 0.0f);
            yb[static_cast<int64_t>(oc) * g.P] = v;
        }
    }
};

// ---- layer policies -------------------------------------------------------------------------------
struct Layer2 {                            // fp32 input, k4 s2 p1, 16 -> 32
    static constexpr int kN = 32, kTerms = 3, KH = 4, KW = 4, S = 2, PAD = 1;
    static constexpr int kStages = 4;      // 4 x 32 KiB (hi + lo) + 64 KiB filter bank
    static constexpr int kKB = 8;          // K = 16*4*4 = 256
    using In = float;
    using Raw = float4;
    using Rows = FwdRows<Layer2>;
    __device__ static float4 expand(const float4& v) { return v; }
    __device__ static float4 zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
    __device__ static float4 gather(const RowCtx& rc, const Geom& g, int kb, int j) {
        const float* __restrict__ xn = static_cast<const float*>(rc.xn);
        const int oy = rc.oy, ox = rc.ox;
        const int c = kb * 2 + (j >> 2), ky = j & 3;
        const int iy = oy * S + ky - PAD, ix0 = ox * S - PAD;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (iy >= 0 && iy < g.H) {
            const float* row = xn + (c * g.H + iy) * g.W;
            if (ix0 >= 0) v.x = row[ix0];
            v.y = row[ix0 + 1];
            if (ix0 + 2 < g.W) v.z = row[ix0 + 2];
            if (ix0 + 3 < g.W) v.w = row[ix0 + 3];
        }
        return v;
    }
    static constexpr float kScale = 1.0f;
};

struct Layer1 {                            // uint8 input, k8 s4 p0, 4 -> 16; pixels stay integers
    static constexpr int kN = 16, kTerms = 2, KH = 8, KW = 8, S = 4, PAD = 0;
    static constexpr int kStages = 8;      // 8 x 16 KiB (hi only) + 32 KiB filter bank
    static constexpr int kKB = 8;          // K = 4*8*8 = 256
    using In = uint8_t;
    using Rows = FwdRows<Layer1>;
    using Raw = uint32_t;                  // 4 packed pixels; converted only when written to smem so the
                                           // load stays in flight across the previous stage's stores
    __device__ static uint32_t zero() { return 0u; }
    __device__ static float4 expand(uint32_t v) {
        return make_float4(static_cast<float>(v & 0xffu), static_cast<float>((v >> 8) & 0xffu),
                           static_cast<float>((v >> 16) & 0xffu), static_cast<float>(v >> 24));
    }
    __device__ static uint32_t gather(const RowCtx& rc, const Geom& g, int kb, int j) {
        const uint8_t* __restrict__ xn = static_cast<const uint8_t*>(rc.xn);
        const int oy = rc.oy, ox = rc.ox;
        const int k0 = kb * 32 + j * 4;
        const int c = k0 >> 6, ky = (k0 >> 3) & 7, kx0 = k0 & 7;
        return *reinterpret_cast<const uint32_t*>(xn + (c * g.H + oy * S + ky) * g.W + ox * S + kx0);
    }
    static constexpr float kScale = 1.0f / 255.0f;
};

// Input gradient of layer 2 by input-pixel parity (see the header comment): rows = (parity, n, a, b).
struct Dgrad2 {
    static constexpr int kN = 16, kTerms = 3, kStages = 4, kKB = 4;   // K = 32 channels x 2x2 taps = 128
    static constexpr float kScale = 1.0f;
    using In = float;                      // ReLU-masked output gradient [N,32,OH,OW]
    using Raw = float4;
    __device__ static float4 expand(const float4& v) { return v; }
    __device__ static float4 zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
    // parity 0 (even pixel) takes taps {1,3} from output rows {a, a-1}; parity 1 taps {0,2} from {a+1, a}
    __device__ static int dpos(int parity, int t) { return parity == 0 ? (t == 0 ? 0 : -1) : (t == 0 ? 1 : 0); }
    __device__ static float4 gather(const RowCtx& rc, const Geom& g, int kb, int j) {
        const float* __restrict__ gp = static_cast<const float*>(rc.xn) + (kb * 8 + j) * g.P;   // channel oc
        float v[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int oy = rc.oy + dpos(rc.py, t >> 1), ox = rc.ox + dpos(rc.px, t & 1);
            v[t] = (oy >= 0 && oy < g.OH && ox >= 0 && ox < g.OW) ? gp[oy * g.OW + ox] : 0.0f;
        }
        return make_float4(v[0], v[1], v[2], v[3]);
    }
    struct Rows {
        static constexpr int kBSets = 4;
        __device__ static int64_t tiles_per_parity(const Geom& g) { return (g.m_total + kRows - 1) / kRows; }
        __device__ static int64_t num_tiles(const Geom& g) { return 4 * tiles_per_parity(g); }
        // tiles are parity-interleaved (tile = 4*row_block + parity): the four parity classes of one
        // region of the gradient run on neighbouring CTAs at the same time, so the gradient is read
        // from HBM once and the stride-2 stores of the four classes merge in L2 (parity-major order
        // re-read it 4x: 910 MB instead of 105 MB, profiles/r01_tc_kernels_full_summary.json)
        __device__ static int bset(const Geom&, int64_t tile) { return static_cast<int>(tile & 3); }
        __device__ static RowCtx decode(const float* X, const int64_t*, const Geom& g, int64_t tile, int r) {
            RowCtx c;
            const int par = static_cast<int>(tile & 3);
            const int64_t m = (tile >> 2) * kRows + r;
            c.py = par >> 1; c.px = par & 1;
            c.valid = m < g.m_total;
            const int per_img = g.BH * g.BW;
            c.n = c.valid ? m / per_img : 0;
            const int blk = c.valid ? static_cast<int>(m - c.n * per_img) : 0;
            c.oy = blk / g.BW;             // a
            c.ox = blk - c.oy * g.BW;      // b
            c.xn = X + c.n * (static_cast<int64_t>(32) * g.P);
            return c;
        }
        __device__ static void store(const RowCtx& c, const Geom& g, float* __restrict__ dX, const float*,
                                     const float (&acc)[kN], int) {
            const int iy = 2 * c.oy + c.py, ix = 2 * c.ox + c.px;
            if (iy >= g.H || ix >= g.W) return;
            float* xb = dX + c.n * (static_cast<int64_t>(kN) * g.H * g.W) + iy * g.W + ix;
#pragma unroll
            for (int ch = 0; ch < kN; ++ch) xb[static_cast<int64_t>(ch) * g.H * g.W] = acc[ch];
        }
    };
};

template <class L>
struct Smem {
    static constexpr int kBBytes = L::Rows::kBSets * L::kKB * L::kN * 128;   // one term of B, all sets / k-blocks
    static constexpr int kATerms = (L::kTerms == 3) ? 2 : 1;          // A hi (+ lo)
    static constexpr int kStageBytes = kATerms * kTileBytes;
    static constexpr int kTotal = 2 * kBBytes + L::kStages * kStageBytes + 1024 + 256;
};

template <class L>
__global__ void __launch_bounds__(kThreads, 1)
conv_fwd_tc_kernel(const typename L::In* __restrict__ X, const int64_t* __restrict__ rows,
                   const float* __restrict__ Wg, const float* __restrict__ bias, float* __restrict__ Y,
                   Geom g, int relu) {
    using S = Smem<L>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* b_hi = smem;
    uint8_t* b_lo = smem + S::kBBytes;
    uint8_t* a_ring = smem + 2 * S::kBBytes;
    constexpr int kStages = L::kStages;
    uint64_t* bars = reinterpret_cast<uint64_t*>(a_ring + kStages * S::kStageBytes);
    uint64_t* a_full = bars;                   // [kStages] producers -> MMA
    uint64_t* a_empty = bars + kStages;        // [kStages] MMA -> producers
    uint64_t* acc_full = bars + 2 * kStages;   // [2] MMA -> epilogue
    uint64_t* acc_empty = acc_full + 2;        // [2] epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    using Rows = typename L::Rows;
    constexpr int kKB = L::kKB, kK = L::kKB * 32;
    const int64_t num_tiles = Rows::num_tiles(g);
    constexpr uint32_t kIdesc = make_idesc_tf32(kRows, L::kN);

    // ---- one-time setup: barriers, TMEM, filter bank -> swizzled hi/lo tiles
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&a_full[s], kProducerThreads / 32);   // one elected arrive per producer warp
            mbar_init(&a_empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&acc_full[b], 1);
            mbar_init(&acc_empty[b], kEpilogueThreads / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kMmaWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    for (int i = threadIdx.x; i < Rows::kBSets * L::kN * (kK / 4); i += kThreads) {   // one 16-byte chunk each
        const int set = i / (L::kN * (kK / 4)), rem = i % (L::kN * (kK / 4));
        const int oc = rem / (kK / 4), ch = rem % (kK / 4);            // ch = chunk index within the row (k = 4*ch)
        const int kb = (set * kKB) + (ch >> 3), j = ch & 7;            // B tiles are stored set-major, then k-block
        const float4 w = *reinterpret_cast<const float4*>(Wg + (static_cast<int64_t>(set) * L::kN + oc) * kK + ch * 4);
        float4 hi, lo;
        split_tf32(w.x, hi.x, lo.x); split_tf32(w.y, hi.y, lo.y);
        split_tf32(w.z, hi.z, lo.z); split_tf32(w.w, hi.w, lo.w);
        const int off = kb * L::kN * 128 + oc * 128 + ((j ^ (oc & 7)) << 4);
        *reinterpret_cast<float4*>(b_hi + off) = hi;
        *reinterpret_cast<float4*>(b_lo + off) = lo;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 8) {
        // ================================================================ A producers
        const int r = threadIdx.x & 127;           // row of the tile
        const int jh = (threadIdx.x >> 7) * 4;     // this thread's chunks: jh .. jh+3
        const int swz = r & 7;
        int s = 0;
        uint32_t ph = 0;

        auto decode = [&](int64_t tile) { return Rows::decode(X, rows, g, tile, r); };
        using Raw = typename L::Raw;
        auto fetch = [&](const RowCtx& c, int kb, Raw (&v)[4]) {
#pragma unroll
            for (int q = 0; q < 4; ++q) v[q] = c.valid ? L::gather(c, g, kb, jh + q) : L::zero();
        };

        auto put = [&](const Raw (&c4)[4]) {       // write this thread's 4 chunks of one k-block
            mbar_wait(&a_empty[s], ph ^ 1);
            uint8_t* st = a_ring + s * S::kStageBytes + r * 128;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 v = L::expand(c4[q]);
                const int off = ((jh + q) ^ swz) << 4;
                if (L::kTerms == 3) {
                    float4 hi, lo;
                    split_tf32(v.x, hi.x, lo.x); split_tf32(v.y, hi.y, lo.y);
                    split_tf32(v.z, hi.z, lo.z); split_tf32(v.w, hi.w, lo.w);
                    *reinterpret_cast<float4*>(st + off) = hi;
                    *reinterpret_cast<float4*>(st + kTileBytes + off) = lo;
                } else {
                    *reinterpret_cast<float4*>(st + off) = v;   // integers 0..255: exact in TF32
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&a_full[s]);   // one arrive per warp (256 smem atomics would serialise)
            if (++s == kStages) { s = 0; ph ^= 1; }
        };

        int64_t tile = blockIdx.x;
        RowCtx ctx = decode(tile);
        if (sizeof(Raw) == 4) {
            // small raw chunks (uint8 layer): prefetch a WHOLE TILE ahead - 32 words per thread - so a
            // full memory round trip overlaps the 8 k-blocks of the current tile
            Raw curT[kKB][4], nxtT[kKB][4];
            if (tile < num_tiles) {
#pragma unroll
                for (int kb = 0; kb < kKB; ++kb) fetch(ctx, kb, curT[kb]);
            }
            while (tile < num_tiles) {
                const int64_t tile_next = tile + gridDim.x;
                if (tile_next < num_tiles) {
                    ctx = decode(tile_next);
#pragma unroll
                    for (int kb = 0; kb < kKB; ++kb) fetch(ctx, kb, nxtT[kb]);
                }
#pragma unroll
                for (int kb = 0; kb < kKB; ++kb) put(curT[kb]);
#pragma unroll
                for (int kb = 0; kb < kKB; ++kb)
#pragma unroll
                    for (int q = 0; q < 4; ++q) curT[kb][q] = nxtT[kb][q];
                tile = tile_next;
            }
        } else {
            // 16-byte raw chunks (fp32 layer): one k-block ahead (register budget)
            Raw cur[4], nxt[4];
            if (tile < num_tiles) fetch(ctx, 0, cur);
            while (tile < num_tiles) {
                const int64_t tile_next = tile + gridDim.x;
                RowCtx ctx_next = ctx;
                for (int kb = 0; kb < kKB; ++kb) {
                    if (kb + 1 < kKB) {
                        fetch(ctx, kb + 1, nxt);
                    } else if (tile_next < num_tiles) {
                        ctx_next = decode(tile_next);
                        fetch(ctx_next, 0, nxt);
                    }
                    put(cur);
#pragma unroll
                    for (int q = 0; q < 4; ++q) cur[q] = nxt[q];
                }
                tile = tile_next;
                ctx = ctx_next;
            }
        }
    } else if (warp == kMmaWarp) {
        // ================================================================ MMA issuer
        uint32_t it = 0, tcount = 0;
        for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tcount) {
            const int buf = tcount & 1;
            const int set_kb0 = Rows::bset(g, tile) * kKB;
            mbar_wait(&acc_empty[buf], ((tcount >> 1) & 1) ^ 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (int kb = 0; kb < kKB; ++kb, ++it) {
                const int s = it % kStages;
                const uint32_t ph = (it / kStages) & 1;
                mbar_wait(&a_full[s], ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (lane == 0) {
                    const int half = kb / (kKB / 2);
                    const uint32_t acc = tmem_base + static_cast<uint32_t>(buf * 2 * L::kN + half * L::kN);
                    const uint8_t* st = a_ring + s * S::kStageBytes;
                    const uint64_t da_hi = make_desc(st), da_lo = make_desc(st + kTileBytes);
                    const uint64_t db_hi = make_desc(b_hi + (set_kb0 + kb) * L::kN * 128), db_lo = make_desc(b_lo + (set_kb0 + kb) * L::kN * 128);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t adv = static_cast<uint64_t>(k * 2);
                        umma_tf32(acc, da_hi + adv, db_hi + adv, kIdesc, ((kb % (kKB / 2)) > 0 || k > 0) ? 1u : 0u);
                        umma_tf32(acc, da_hi + adv, db_lo + adv, kIdesc, 1u);
                        if (L::kTerms == 3) umma_tf32(acc, da_lo + adv, db_hi + adv, kIdesc, 1u);
                    }
                    umma_commit(&a_empty[s]);
                    if (kb == kKB - 1) umma_commit(&acc_full[buf]);
                }
                __syncwarp();
            }
        }
    } else {
        // ================================================================ epilogue (warps 8..11)
        const int q = warp - kEpiWarp0;
        const int r = q * 32 + lane;
        const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
        uint32_t tcount = 0;
        for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tcount) {
            const int buf = tcount & 1;
            mbar_wait(&acc_full[buf], (tcount >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t r0[32], r1[32];
            const uint32_t col = static_cast<uint32_t>(buf * 2 * L::kN);
            if (L::kN == 32) {
                tmem_ld32(tmem_base + lane_base + col, r0);
                tmem_ld32(tmem_base + lane_base + col + 32, r1);
            } else {
                tmem_ld32(tmem_base + lane_base + col, r0);       // 2 x 16 columns: both halves in one load
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);
            const RowCtx rc = Rows::decode(X, rows, g, tile, r);
            if (rc.valid) {
                float acc[L::kN];
#pragma unroll
                for (int oc = 0; oc < L::kN; ++oc)
                    acc[oc] = (L::kN == 32) ? __uint_as_float(r0[oc]) + __uint_as_float(r1[oc])
                                            : __uint_as_float(r0[oc]) + __uint_as_float(r0[16 + oc]);
                Rows::store(rc, g, Y, bias, acc, relu);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == kMmaWarp) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols));
    }
}

template <class L>
static int launch(const typename L::In* X, const int64_t* rows, const float* W, const float* bias, float* Y,
                  Geom g, int relu, cudaStream_t st) {
    using S = Smem<L>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(conv_fwd_tc_kernel<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal);
        attr_set = true;
    }
    int sms = sm_count();
    if (sms <= 0) sms = 148;
    const int64_t num_tiles = L::Rows::kBSets * ((g.m_total + kRows - 1) / kRows);
    const int64_t grid = num_tiles < sms ? num_tiles : sms;
    conv_fwd_tc_kernel<L><<<static_cast<unsigned>(grid), kThreads, S::kTotal, st>>>(X, rows, W, bias, Y, g, relu);
    return check_launch("conv_fwd_tc_kernel");
}

// B_p[par][c][oc*4 + ty*2 + tx] = W[oc][c][ky(py,ty)][kx(px,tx)], tap(parity 0) = {1,3}, tap(parity 1) = {0,2}
__global__ void conv2_dgrad_prep_kernel(const float* __restrict__ W, float* __restrict__ Bp) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;       // over 4*16*128
    if (i >= 4 * 16 * 128) return;
    const int par = i / (16 * 128), c = (i / 128) % 16, k = i % 128;
    const int oc = k >> 2, ty = (k >> 1) & 1, tx = k & 1;
    const int py = par >> 1, px = par & 1;
    const int ky = py == 0 ? (ty == 0 ? 1 : 3) : (ty == 0 ? 0 : 2);
    const int kx = px == 0 ? (tx == 0 ? 1 : 3) : (tx == 0 ? 0 : 2);
    Bp[i] = W[((oc * 16 + c) * 4 + ky) * 4 + kx];
}

}  // namespace convtc
}  // namespace rl

extern "C" {

int rl_conv1_u8_forward_tc(const uint8_t* obs, const int64_t* rows, const float* weight, const float* bias,
                           float* out, int64_t N, int C, int H, int W, int relu, void* stream) {
    RL_REQUIRE(obs && weight && bias && out, RL_EINVAL, "rl_conv1_u8_forward_tc: null pointer");
    RL_REQUIRE(N >= 0 && C == 4 && H >= 8 && W >= 8 && W % 4 == 0, RL_EINVAL,
               "rl_conv1_u8_forward_tc: needs C=4, W %% 4 == 0 (got C=%d H=%d W=%d)", C, H, W);
    RL_REQUIRE(rl::aligned(obs, 4) && rl::aligned(weight, 16), RL_EALIGN, "rl_conv1_u8_forward_tc: alignment");
    if (N == 0) return RL_OK;
    rl::convtc::Geom g;
    g.n_img = static_cast<int>(N); g.C = C; g.H = H; g.W = W;
    g.OH = (H - 8) / 4 + 1; g.OW = (W - 8) / 4 + 1; g.P = g.OH * g.OW; g.m_total = N * g.P; g.BH = g.BW = 0;
    return rl::convtc::launch<rl::convtc::Layer1>(obs, rows, weight, bias, out, g, relu, rl::as_stream(stream));
}

int rl_conv2_forward_tc(const float* x, const float* weight, const float* bias, float* out, int64_t N,
                        int C, int IH, int IW, int relu, void* stream) {
    RL_REQUIRE(x && weight && bias && out, RL_EINVAL, "rl_conv2_forward_tc: null pointer");
    RL_REQUIRE(N >= 0 && C == 16 && IH >= 2 && IW >= 2, RL_EINVAL, "rl_conv2_forward_tc: needs C=16 (got C=%d %dx%d)",
               C, IH, IW);
    RL_REQUIRE(rl::aligned(weight, 16), RL_EALIGN, "rl_conv2_forward_tc: weight must be 16B aligned");
    if (N == 0) return RL_OK;
    rl::convtc::Geom g;
    g.n_img = static_cast<int>(N); g.C = C; g.H = IH; g.W = IW;
    g.OH = (IH + 2 - 4) / 2 + 1; g.OW = (IW + 2 - 4) / 2 + 1; g.P = g.OH * g.OW; g.m_total = N * g.P; g.BH = g.BW = 0;
    return rl::convtc::launch<rl::convtc::Layer2>(x, nullptr, weight, bias, out, g, relu, rl::as_stream(stream));
}

int64_t rl_conv2_dgrad_tc_scratch_bytes(void) { return 4 * 16 * 128 * static_cast<int64_t>(sizeof(float)); }

int rl_conv2_dgrad_tc(const float* grad_out_masked, const float* weight, float* grad_x, int64_t N, int C,
                      int IH, int IW, void* scratch, void* stream) {
    RL_REQUIRE(grad_out_masked && weight && grad_x && scratch, RL_EINVAL, "rl_conv2_dgrad_tc: null pointer");
    RL_REQUIRE(N >= 0 && C == 16 && IH >= 2 && IW >= 2, RL_EINVAL, "rl_conv2_dgrad_tc: needs C=16 (got C=%d %dx%d)", C,
               IH, IW);
    RL_REQUIRE(rl::aligned(scratch, 16), RL_EALIGN, "rl_conv2_dgrad_tc: scratch must be 16B aligned");
    if (N == 0) return RL_OK;
    cudaStream_t st = rl::as_stream(stream);
    float* Bp = static_cast<float*>(scratch);
    rl::convtc::conv2_dgrad_prep_kernel<<<(4 * 16 * 128 + 255) / 256, 256, 0, st>>>(weight, Bp);
    int rc = rl::check_launch("conv2_dgrad_prep_kernel");
    if (rc != RL_OK) return rc;
    rl::convtc::Geom g;
    g.n_img = static_cast<int>(N); g.C = C; g.H = IH; g.W = IW;
    g.OH = (IH + 2 - 4) / 2 + 1; g.OW = (IW + 2 - 4) / 2 + 1; g.P = g.OH * g.OW;
    g.BH = (IH + 1) / 2; g.BW = (IW + 1) / 2;
    g.m_total = N * static_cast<int64_t>(g.BH) * g.BW;       // rows per parity class
    return rl::convtc::launch<rl::convtc::Dgrad2>(grad_out_masked, nullptr, Bp, nullptr, grad_x, g, 0, st);
}

}  // extern "C"
