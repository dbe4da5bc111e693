// Convolutions as IMPLICIT GEMMs on the 5th-gen tensor cores (tcgen05, kind::tf32 with the
// fp32-accurate hi/lo split of gemm_tf32x3.cu) for the two conv layers of the AtariFf network:
// forward of both layers, input gradient of layer 2, weight + bias gradients of both layers
// (the weight-gradient kernel has its own header comment further down).
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
//   * MMA warp: per k-block 4 x (2|3) tcgen05.mma (M=128, N=16|32, K=8), issued from warp-uniform code
//     under elect.sync so descriptors stay in uniform registers (tc_common.cuh); the two halves of K go to
//     separate TMEM accumulators that the epilogue adds in fp32 (halves the accumulator truncation),
//     and the accumulators are double buffered so the epilogue of tile i overlaps the MMAs of tile i+1.
//   * epilogue (4 warps): tcgen05.ld -> (+bias, *scale, ReLU) -> NCHW stores, coalesced over positions.
// The same kernel also computes the INPUT GRADIENT of layer 2 (policy Dgrad2): the transposed
// convolution is split by the parity (iy&1, ix&1) of the input pixel - each parity class sees a fixed
// 2x2 subset of the 4x4 taps - into four dense GEMMs with K = 32 channels x 4 taps = 128 whose rows are
// the input pixels of that parity, A gathers the ReLU-masked output gradient, B_p[c][oc*4+t] is the
// matching slice of the filter bank (rl_conv2_dgrad_tc prepares the four B_p).
// Measured per 8192-sample minibatch (round 1): layer-1 forward 0.45 ms (fp32 SIMT
// kernel of conv1.cu 0.82 ms; cuDNN incl. the u8->f32 conversion 3.2 ms), layer-2 forward 0.29 ms (cuDNN
// fp32 0.68 ms), layer-2 input gradient 0.60 ms (cuDNN 1.06 ms), weight gradients 0.78 / 0.30 ms
// (SIMT 1.06 ms / cuDNN 0.79 ms).
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
    // split addressing for the weight-gradient producers: element = image + row_off(position) + tap.off,
    // the tap part is a per-thread constant there
    struct Tap { int off, ky; };
    __device__ static Tap tap(const Geom& g, int kb, int j) {
        const int c = kb * 2 + (j >> 2), ky = j & 3;
        return Tap{(c * g.H + ky - PAD) * g.W - PAD, ky};
    }
    __device__ static int row_off(const Geom& g, int oy, int ox) { return oy * S * g.W + ox * S; }
    __device__ static float4 load(const float* __restrict__ p, const Tap& t, const Geom& g, int oy, int ox) {
        const int iy = oy * S + t.ky - PAD, ix0 = ox * S - PAD;
        const float* q = p + t.off;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (iy >= 0 && iy < g.H) {
            if (ix0 >= 0) v.x = q[0];
            v.y = q[1];
            if (ix0 + 2 < g.W) v.z = q[2];
            if (ix0 + 3 < g.W) v.w = q[3];
        }
        return v;
    }
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
    __device__ static float4 expand(uint32_t v) {   // I2F.U8; a PRMT+FADD variant measured 7% slower
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
    struct Tap { int off; };
    __device__ static Tap tap(const Geom& g, int kb, int j) {
        const int k0 = kb * 32 + j * 4;
        return Tap{((k0 >> 6) * g.H + ((k0 >> 3) & 7)) * g.W + (k0 & 7)};
    }
    __device__ static int row_off(const Geom& g, int oy, int ox) { return oy * S * g.W + ox * S; }
    __device__ static uint32_t load(const uint8_t* __restrict__ p, const Tap& t, const Geom&, int, int) {
        return *reinterpret_cast<const uint32_t*>(p + t.off);
    }
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

    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
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
    const uint32_t tmem_base = uniform_u32(*tmem_slot);

    if (warp < 8) {
        // ================================================================ A producers
        const int r = threadIdx.x & 127;           // row of the tile
        const int jh = (threadIdx.x >> 7) * 4;     // this thread's chunks: jh .. jh+3
        const int swz = r & 7;
        const uint32_t a_ring_u32 = smem_u32(a_ring);
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
            const uint32_t st = a_ring_u32 + static_cast<uint32_t>(s * S::kStageBytes + r * 128);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 v = L::expand(c4[q]);
                const uint32_t off = static_cast<uint32_t>(((jh + q) ^ swz) << 4);
                if (L::kTerms == 3) {
                    float4 hi, lo;
                    split_tf32(v.x, hi.x, lo.x); split_tf32(v.y, hi.y, lo.y);
                    split_tf32(v.z, hi.z, lo.z); split_tf32(v.w, hi.w, lo.w);
                    sts128(st + off, hi);
                    sts128(st + kTileBytes + off, lo);
                } else {
                    sts128(st + off, v);                        // integers 0..255: exact in TF32
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
                if (elect_one()) {
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


// ====================================================================================================
// Weight gradient as a tcgen05 GEMM over all output positions:
//     D[tap, oc] = sum_m A[m, tap] * G[m, oc]        (A = im2col(X), G = ReLU-masked output gradient)
// i.e. M_gemm = 256 taps (two 128-row MMA tiles), N_gemm = oc, K_gemm = m.  Both operands must be
// K_gemm-major in shared memory (32 consecutive positions per 128-byte row):
//   * A^T tile [128 taps x 32 m] per half: a producer thread gathers a 4-position x 4-tap block (four of
//     the forward pass' row chunks), transposes it in registers and writes four tap rows; lanes run
//     over consecutive positions of one tap chunk so the gathers stay coalesced;
//   * G^T tile [oc x 32 m]: 16-byte reads of the NCHW gradient (positions are contiguous per channel).
// Each persistent CTA owns a contiguous range of k-blocks (positions), accumulates in TMEM with the
// same drain-to-fp32-registers promotion as gemm_tf32x3.cu (every kPromote k-blocks, two TMEM
// buffers), and writes its partial D; wgrad_reduce_kernel sums the partials in CTA order
// (deterministic), applies the layer scale and transposes to the [oc][tap] weight layout.
constexpr int kWgStagesL2 = 2, kWgStagesL1 = 4, kPromote = 4;
constexpr int kRowTab = 2048;              // row-index entries staged in shared memory per CTA (kRowMode 1)

template <class L>
struct WgSmem {
    static constexpr int kATerms = (L::kTerms == 3) ? 2 : 1;
    static constexpr int kABytes = 2 * kATerms * kTileBytes;          // two tap halves x (hi [+ lo])
    static constexpr int kBTile = L::kN * 128;                        // G^T: oc rows x 128 B
    static constexpr int kStageBytes = kABytes + 2 * kBTile;          // + G hi, lo
    static constexpr int kStages = (L::kTerms == 3) ? kWgStagesL2 : kWgStagesL1;
    static constexpr int kTotal = kStages * kStageBytes + 1024 + 256 + kRowTab * 8;
};

// kRowMode: 0 = images in order (rows == nullptr), 1 = this CTA's slice of the row-index table staged in
// shared memory, 2 = row indices read from global memory (CTA spans more than kRowTab images).
// A compile-time mode, not a runtime branch: a dependent global load in the fetch loop - even a
// predicated-off one - makes its consumer wait on a scoreboard shared with the gathers still in flight
// from the previous k-blocks, which serialises the whole prefetch (profiles/r01_wgrad_tc_notes.md).
template <class L, int kRowMode>
__global__ void __launch_bounds__(kThreads, 1)
conv_wgrad_tc_kernel(const typename L::In* __restrict__ X, const int64_t* __restrict__ rows,
                     const float* __restrict__ Out, const float* __restrict__ Gm, float* __restrict__ partial,
                     float* __restrict__ partial_bias, Geom g) {
    using S = WgSmem<L>;
    using Rows = typename L::Rows;
    constexpr int kStages = S::kStages;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * S::kStageBytes);
    uint64_t* s_full = bars;
    uint64_t* s_empty = bars + kStages;
    uint64_t* acc_full = bars + 2 * kStages;
    uint64_t* acc_empty = acc_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    int64_t* row_tab = reinterpret_cast<int64_t*>(smem + kStages * S::kStageBytes + 256);
    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    constexpr uint32_t kIdesc = make_idesc_tf32(kRows, L::kN);
    constexpr int kDepth = (sizeof(typename L::Raw) == 4) ? 3 : 2;   // k-blocks held in registers per producer

    const int64_t total_kb = (g.m_total + 31) / 32;
    const int64_t per_cta = (total_kb + gridDim.x - 1) / gridDim.x;
    const int64_t kb_begin = static_cast<int64_t>(blockIdx.x) * per_cta;
    const int64_t kb_end = kb_begin + per_cta < total_kb ? kb_begin + per_cta : total_kb;
    const int64_t my_kb = kb_end > kb_begin ? kb_end - kb_begin : 0;
    const int64_t num_chunks = (my_kb + kPromote - 1) / kPromote;
    const int64_t n_lo = (kb_begin * 32) / g.P;            // first image this CTA touches
    if (kRowMode == 1) {
        for (int i = threadIdx.x; i < kRowTab; i += kThreads) {
            const int64_t n = n_lo + i;
            row_tab[i] = n < g.n_img ? rows[n] : 0;
        }
    }

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&s_full[s], kProducerThreads / 32);
            mbar_init(&s_empty[s], 1);
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
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = uniform_u32(*tmem_slot);

    if (warp < 8) {
        // ================================================================ producers
        // thread -> (4-position chunk mc, 4-tap group tg of each half).  Lanes (2mc, 2mc+1) hold the
        // two tap groups that share a 32-byte sector, 16 lanes cover the 32 positions of the k-block
        // (coalesced gathers); the 8 lanes of one tap group write one full 128-byte row per store
        // (conflict-free).  The same thread also owns kGP positions of one G^T row, chosen so that
        // they are a subset of its own four positions (one decode serves both operands) and the G work
        // is spread over all 256 threads (kN=16: two positions each, kN=32: four).
        using Raw = typename L::Raw;
        constexpr int kGP = L::kN / 8;                         // G positions per thread
        const int mc = (lane >> 1) & 7;
        const int tg = (lane & 1) + 2 * (lane >> 4) + 4 * warp;
        const int g_oc = (tg * kGP) >> 2, g_r0 = (tg * kGP) & 3;
        const uint32_t smem_base_u32 = smem_u32(smem);
        int s = 0;
        uint32_t ph = 0;
        float bias_acc = 0.0f;                         // sum of this thread's masked gradients (channel g_oc)

        // running (image, output row, output column) of this thread's first position in the next
        // k-block to fetch; advanced by 32 positions per fetch (fetches are issued in k-block order)
        // with carries only - the loop has no division and the tap part of every address is a
        // per-thread constant
        const typename L::Tap tap0 = L::tap(g, tg >> 3, tg & 7), tap1 = L::tap(g, 4 + (tg >> 3), tg & 7);
        const int adv_y = 32 / g.OW, adv_x = 32 - adv_y * g.OW;
        int n_next, oy_next, ox_next;
        {
            const int64_t m0 = kb_begin * 32 + 4 * mc;
            n_next = static_cast<int>(m0 / g.P);
            const int pos = static_cast<int>(m0 - static_cast<int64_t>(n_next) * g.P);
            oy_next = pos / g.OW;
            ox_next = pos - oy_next * g.OW;
        }
        const int64_t img_stride = static_cast<int64_t>(g.C) * g.H * g.W;
        const int n_lo32 = static_cast<int>(n_lo);
        const uint32_t row_tab_u32 = smem_u32(row_tab);
        auto image_ptr = [&](int n) -> const typename L::In* {
            int64_t img = n;
            if (kRowMode == 1) img = lds_s64(row_tab_u32 + static_cast<uint32_t>(min(max(n - n_lo32, 0), kRowTab - 1)) * 8u);
            if (kRowMode == 2) img = n < g.n_img ? rows[n] : 0;
            return X + img * img_stride;
        };
        const float* __restrict__ OutOrG = Out != nullptr ? Out : Gm;    // unconditional second load; the
        const bool has_mask = Out != nullptr;                             // mask is applied when stored

        struct Held { Raw a[2][4]; float gv[kGP]; float ov[kGP]; };
        auto fetch = [&](Held& hd) {
            int n = n_next, oy = oy_next, ox = ox_next;
            const typename L::In* img_ptr = image_ptr(n);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const bool valid = n < g.n_img;
                const typename L::In* p = img_ptr + L::row_off(g, oy, ox);
                hd.a[0][r] = L::zero();
                hd.a[1][r] = L::zero();
                if (valid) {
                    hd.a[0][r] = L::load(p, tap0, g, oy, ox);
                    hd.a[1][r] = L::load(p, tap1, g, oy, ox);
                }
                if (++ox == g.OW) {
                    ox = 0;
                    if (++oy == g.OH) { oy = 0; ++n; img_ptr = image_ptr(n); }
                }
            }
            // this thread's kGP positions of the G^T row start g_r0 positions into the chunk; they are
            // walked separately so that the destination registers are static (a select on a runtime
            // row index would consume the load at once and stall on its scoreboard)
            {
                int gn = n_next, goy = oy_next, gox = ox_next + g_r0;
                while (gox >= g.OW) {
                    gox -= g.OW;
                    if (++goy == g.OH) { goy = 0; ++gn; }
                }
#pragma unroll
                for (int j = 0; j < kGP; ++j) {
                    float v = 0.0f, o = 1.0f;
                    if (gn < g.n_img) {
                        const int64_t gi = static_cast<int64_t>(gn * L::kN + g_oc) * g.P + (goy * g.OW + gox);
                        v = Gm[gi];
                        o = OutOrG[gi];
                    }
                    hd.gv[j] = v;
                    hd.ov[j] = o;
                    if (++gox == g.OW) {
                        gox = 0;
                        if (++goy == g.OH) { goy = 0; ++gn; }
                    }
                }
            }
            ox_next += adv_x;
            oy_next += adv_y;
            if (ox_next >= g.OW) { ox_next -= g.OW; ++oy_next; }
            while (oy_next >= g.OH) { oy_next -= g.OH; ++n_next; }
        };
        auto put = [&](const Held& hd) {
            mbar_wait(&s_empty[s], ph ^ 1);
            const uint32_t st = smem_base_u32 + static_cast<uint32_t>(s * S::kStageBytes);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float4 f[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) f[r] = L::expand(hd.a[h][r]);
                // transpose the 4 positions x 4 taps block: tap i <- (f[0][i], f[1][i], f[2][i], f[3][i])
                const float4 t[4] = {make_float4(f[0].x, f[1].x, f[2].x, f[3].x), make_float4(f[0].y, f[1].y, f[2].y, f[3].y),
                                     make_float4(f[0].z, f[1].z, f[2].z, f[3].z), make_float4(f[0].w, f[1].w, f[2].w, f[3].w)};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int row = 4 * tg + i;
                    const uint32_t dst = st + static_cast<uint32_t>(h * S::kATerms * kTileBytes + row * 128 + ((mc ^ (row & 7)) << 4));
                    if (L::kTerms == 3) {
                        float4 hi, lo;
                        split_tf32(t[i].x, hi.x, lo.x); split_tf32(t[i].y, hi.y, lo.y);
                        split_tf32(t[i].z, hi.z, lo.z); split_tf32(t[i].w, hi.w, lo.w);
                        sts128(dst, hi);
                        sts128(dst + kTileBytes, lo);
                    } else {
                        sts128(dst, t[i]);
                    }
                }
            }
            {
                float hi[kGP], lo[kGP];
#pragma unroll
                for (int r = 0; r < kGP; ++r) {
                    const float v = (!has_mask || hd.ov[r] > 0.0f) ? hd.gv[r] : 0.0f;
                    bias_acc += v;
                    split_tf32(v, hi[r], lo[r]);
                }
                const uint32_t dst = st + static_cast<uint32_t>(S::kABytes + g_oc * 128 + ((mc ^ (g_oc & 7)) << 4) + g_r0 * 4);
                if (kGP == 4) {
                    sts128(dst, make_float4(hi[0], hi[1], hi[kGP - 2], hi[kGP - 1]));
                    sts128(dst + S::kBTile, make_float4(lo[0], lo[1], lo[kGP - 2], lo[kGP - 1]));
                } else {
                    sts64(dst, make_float2(hi[0], hi[1]));
                    sts64(dst + S::kBTile, make_float2(lo[0], lo[1]));
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_full[s]);
            if (++s == kStages) { s = 0; ph ^= 1; }
        };
        Held held[kDepth];
#pragma unroll
        for (int d = 0; d < kDepth; ++d)
            if (kb_begin + d < kb_end) fetch(held[d]);
        for (int64_t kb = kb_begin; kb < kb_end; kb += kDepth) {
#pragma unroll
            for (int d = 0; d < kDepth; ++d) {
                if (kb + d < kb_end) {
                    put(held[d]);
                    if (kb + d + kDepth < kb_end) fetch(held[d]);
                }
            }
        }
        // bias gradient: the threads of one channel are the lanes that differ in mc (and, for two
        // positions per thread, in the low tap-group bit)
        if (kGP == 2) bias_acc += __shfl_xor_sync(0xffffffffu, bias_acc, 1);
        bias_acc += __shfl_xor_sync(0xffffffffu, bias_acc, 2);
        bias_acc += __shfl_xor_sync(0xffffffffu, bias_acc, 4);
        bias_acc += __shfl_xor_sync(0xffffffffu, bias_acc, 8);
        if ((lane & 0xe) == 0 && (kGP == 4 || (lane & 1) == 0)) partial_bias[blockIdx.x * L::kN + g_oc] = bias_acc;
    } else if (warp == kMmaWarp) {
        // ================================================================ MMA issuer
        int s = 0;
        uint32_t ph = 0;
        for (int64_t c = 0; c < num_chunks; ++c) {
            const int buf = static_cast<int>(c & 1);
            mbar_wait(&acc_empty[buf], ((c >> 1) & 1) ^ 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int64_t k0 = c * kPromote, k1 = (k0 + kPromote < my_kb) ? k0 + kPromote : my_kb;
            for (int64_t kk = k0; kk < k1; ++kk) {
                mbar_wait(&s_full[s], ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (elect_one()) {
                    const uint8_t* st = smem + s * S::kStageBytes;
                    const uint64_t db_hi = make_desc(st + S::kABytes), db_lo = make_desc(st + S::kABytes + S::kBTile);
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const uint32_t acc = tmem_base + static_cast<uint32_t>(buf * 2 * L::kN + h * L::kN);
                        const uint64_t da_hi = make_desc(st + h * S::kATerms * kTileBytes);
                        const uint64_t da_lo = make_desc(st + h * S::kATerms * kTileBytes + kTileBytes);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint64_t adv = static_cast<uint64_t>(k * 2);
                            umma_tf32(acc, da_hi + adv, db_hi + adv, kIdesc, (kk > k0 || k > 0) ? 1u : 0u);
                            umma_tf32(acc, da_hi + adv, db_lo + adv, kIdesc, 1u);
                            if (L::kTerms == 3) umma_tf32(acc, da_lo + adv, db_hi + adv, kIdesc, 1u);
                        }
                    }
                    umma_commit(&s_empty[s]);
                    if (kk == k1 - 1) umma_commit(&acc_full[buf]);
                }
                __syncwarp();
                if (++s == kStages) { s = 0; ph ^= 1; }
            }
        }
    } else {
        // ================================================================ drain (warps 8..11): TMEM -> fp32 registers
        const int q = warp - kEpiWarp0;
        const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
        float acc[2 * L::kN];
#pragma unroll
        for (int j = 0; j < 2 * L::kN; ++j) acc[j] = 0.0f;
        for (int64_t c = 0; c < num_chunks; ++c) {
            const int buf = static_cast<int>(c & 1);
            mbar_wait(&acc_full[buf], (c >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t r0[32];
            tmem_ld32(tmem_base + lane_base + static_cast<uint32_t>(buf * 2 * L::kN), r0);
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (j < 2 * L::kN) acc[j] += __uint_as_float(r0[j]);
            if (L::kN == 32) {
                uint32_t r1[32];
                tmem_ld32(tmem_base + lane_base + static_cast<uint32_t>(buf * 2 * L::kN + 32), r1);
#pragma unroll
                for (int j = 0; j < 32; ++j) acc[32 + j] += __uint_as_float(r1[j]);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);
        }
        // partial[cta][tap][oc], tap = h*128 + 32*q + lane
        float* out = partial + static_cast<int64_t>(blockIdx.x) * (256 * L::kN);
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int oc = 0; oc < L::kN; ++oc) out[(h * 128 + q * 32 + lane) * L::kN + oc] = acc[h * L::kN + oc];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == kMmaWarp) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols));
    }
}

// dW[oc][tap] = scale * sum_cta partial[cta][tap][oc];  db[oc] = sum_cta partial_bias[cta][oc].
// 64 outputs x 4 partial groups per block: each thread sums every 4th partial in CTA order, the four
// group sums are combined in a fixed order (deterministic); reads coalesced over oc.
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float* __restrict__ partial, const float* __restrict__ partial_bias, int nparts, int N,
                    float scale, float* __restrict__ dW, float* __restrict__ db) {
    __shared__ float sh[4][64];
    const int o = threadIdx.x & 63, grp = threadIdx.x >> 6;
    const int i = blockIdx.x * 64 + o;                        // over tap*N + oc, then N bias entries
    const int total = 256 * N;
    float s = 0.0f;
    if (i < total) {
        for (int b = grp; b < nparts; b += 4) s += partial[static_cast<int64_t>(b) * total + i];
    } else if (i < total + N && db != nullptr) {
        for (int b = grp; b < nparts; b += 4) s += partial_bias[b * N + (i - total)];
    }
    sh[grp][o] = s;
    __syncthreads();
    if (grp == 0) {
        const float r = (sh[0][o] + sh[1][o]) + (sh[2][o] + sh[3][o]);
        if (i < total) {
            const int tap = i / N, oc = i - tap * N;
            dW[oc * 256 + tap] = r * scale;
        } else if (i < total + N && db != nullptr) {
            db[i - total] = r;
        }
    }
}

template <class L>
static int launch_wgrad(const typename L::In* X, const int64_t* rows, const float* Out, const float* Gm, float* dW,
                        float* db, float* scratch, Geom g, cudaStream_t st) {
    using S = WgSmem<L>;
    int sms = sm_count();
    if (sms <= 0) sms = 148;
    const int64_t total_kb = (g.m_total + 31) / 32;
    const int64_t grid = total_kb < sms ? total_kb : sms;
    float* partial_bias = scratch + static_cast<int64_t>(sms) * 256 * 32;
    const int64_t per_cta = (total_kb + grid - 1) / grid;
    const int mode = rows == nullptr ? 0 : ((per_cta * 32 + 35) / g.P + 2 <= kRowTab ? 1 : 2);
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(conv_wgrad_tc_kernel<L, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal);
        cudaFuncSetAttribute(conv_wgrad_tc_kernel<L, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal);
        cudaFuncSetAttribute(conv_wgrad_tc_kernel<L, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal);
        attr_set = true;
    }
    auto kern = mode == 0 ? conv_wgrad_tc_kernel<L, 0> : mode == 1 ? conv_wgrad_tc_kernel<L, 1> : conv_wgrad_tc_kernel<L, 2>;
    kern<<<static_cast<unsigned>(grid), kThreads, S::kTotal, st>>>(X, rows, Out, Gm, scratch, partial_bias, g);
    int rc = check_launch("conv_wgrad_tc_kernel");
    if (rc != RL_OK) return rc;
    wgrad_reduce_kernel<<<(256 * L::kN + L::kN + 63) / 64, 256, 0, st>>>(scratch, partial_bias, static_cast<int>(grid), L::kN,
                                                                   L::kScale, dW, db);
    return check_launch("wgrad_reduce_kernel");
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

int64_t rl_conv_wgrad_tc_scratch_bytes(void) {
    int sms = rl::sm_count();
    if (sms <= 0) sms = 148;
    return static_cast<int64_t>(sms) * (256 * 32 + 32) * static_cast<int64_t>(sizeof(float));
}

int rl_conv1_u8_wgrad_tc(const uint8_t* obs, const int64_t* rows, const float* out, const float* grad_out,
                         float* grad_weight, float* grad_bias, int64_t N, int C, int H, int W, void* scratch,
                         void* stream) {
    RL_REQUIRE(obs && grad_out && grad_weight && scratch, RL_EINVAL, "rl_conv1_u8_wgrad_tc: null pointer");
    RL_REQUIRE(N >= 1 && C == 4 && H >= 8 && W >= 8 && W % 4 == 0, RL_EINVAL,
               "rl_conv1_u8_wgrad_tc: needs C=4, W %% 4 == 0 (got C=%d H=%d W=%d)", C, H, W);
    rl::convtc::Geom g;
    g.n_img = static_cast<int>(N); g.C = C; g.H = H; g.W = W;
    g.OH = (H - 8) / 4 + 1; g.OW = (W - 8) / 4 + 1; g.P = g.OH * g.OW; g.m_total = N * g.P; g.BH = g.BW = 0;
    RL_REQUIRE(g.m_total < (int64_t(1) << 31) - 64, RL_EINVAL, "rl_conv1_u8_wgrad_tc: N*OH*OW must be < 2^31");
    return rl::convtc::launch_wgrad<rl::convtc::Layer1>(obs, rows, out, grad_out, grad_weight, grad_bias,
                                                        static_cast<float*>(scratch), g, rl::as_stream(stream));
}

int rl_conv2_wgrad_tc(const float* x, const float* out, const float* grad_out, float* grad_weight,
                      float* grad_bias, int64_t N, int C, int IH, int IW, void* scratch, void* stream) {
    RL_REQUIRE(x && grad_out && grad_weight && scratch, RL_EINVAL, "rl_conv2_wgrad_tc: null pointer");
    RL_REQUIRE(N >= 1 && C == 16 && IH >= 2 && IW >= 2, RL_EINVAL, "rl_conv2_wgrad_tc: needs C=16 (got C=%d %dx%d)", C,
               IH, IW);
    rl::convtc::Geom g;
    g.n_img = static_cast<int>(N); g.C = C; g.H = IH; g.W = IW;
    g.OH = (IH + 2 - 4) / 2 + 1; g.OW = (IW + 2 - 4) / 2 + 1; g.P = g.OH * g.OW; g.m_total = N * g.P; g.BH = g.BW = 0;
    RL_REQUIRE(g.m_total < (int64_t(1) << 31) - 64, RL_EINVAL, "rl_conv2_wgrad_tc: N*OH*OW must be < 2^31");
    return rl::convtc::launch_wgrad<rl::convtc::Layer2>(x, nullptr, out, grad_out, grad_weight, grad_bias,
                                                        static_cast<float*>(scratch), g, rl::as_stream(stream));
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
