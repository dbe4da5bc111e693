// Round-2 experiment 2: "v2" weight gradient of the first conv layer WITHOUT im2col
// expansion.  Needs no new hardware semantics (only ordinary descriptors), so it does not depend on the
// outcome of tcgen05_shift_probe.cu.
//
//   dW4[(by,bx)][ch, oc] = sum_pos X4[pos, ch] * G[pos - (by*GW + bx), oc]       (tests/test_conv_layout_math.py)
// with X4 the space-to-depth input (ch = (c, ky', kx'), 64 channels) and G the ReLU-masked output gradient,
// zero outside the OH x OW valid positions.  One K-block = one grid row (n, Y) padded to 32 positions:
//   A^T tile [128 x 32]: rows 0..63 = ch, K = X (21 valid of 32); rows 64..127 stay zero (M = 128 MMA)
//       four adjacent words = 4 positions x 4 kx' of one (c, ky') input row -> register transpose -> 4 chunks
//       (96 threads x 4 word loads per K-block of 21 positions; the v1 kernel issues 2048 gathers per 32 positions)
//   B tile  [64 x 32]: rows = (tap, oc), element X = G[n, oc, Y - by, X - bx] (hi and lo tiles); 160 threads, one
//       16-byte load of G and of the activation each, written once aligned (bx = 0) and once shifted (bx = 1)
//   MMA: D[128 x 64] += A^T . B^T, K slices that hold only padding are skipped; D in TMEM, promoted to
//       fp32 registers every kPromote K-blocks (double-buffered), per-CTA partials reduced in CTA order.
// Standalone: nvcc only; checks dW and db against a CPU fp64 loop on the first images and times N = 8192:
//   tools/probes/_bin/conv1_wgrad_v2 [N]
// NOT validated: written at the end of round 1 after the GPU budget was spent.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "tc_common.cuh"

using namespace rl::tc;

namespace w2 {

constexpr int kThreads = 416, kProducerThreads = 256, kDrainThreads = 128, kMmaWarp = 12, kDrainWarp0 = 8;
constexpr int kATile = 128 * 128;           // 16 KiB (rows 64..127 zero)
constexpr int kBTile = 64 * 128;            // 8 KiB
constexpr int kStageBytes = kATile + 2 * kBTile;   // A | G hi | G lo = 32 KiB
constexpr int kStages = 4, kPromote = 8;
constexpr int kN = 64;                      // (tap, oc)
constexpr int kTmemCols = 128;              // 2 buffers x 64 columns
constexpr int kSmemBytes = kStages * kStageBytes + 256 + 1024 + 1024 * 8;   // + row table
constexpr int kDepthA = 3, kDepthG = 2;

struct Geom {
    int n_img, H, W, GH, GW, OH, OW;
    int n16;                                // 16-byte units per input row = ceil(GW / 4)
    int nG;                                 // 4-column chunks per output row = ceil(OW / 4)  (<= 5: 32 x nG G-units on 160 threads)
    int n_slices;                           // K slices (8 positions) that hold valid positions = ceil(GW / 8)
};

// kRowMode 0: images in order; 1: minibatch row indices (rows[n] = image of sample n), this CTA's slice staged in
// shared memory so that the fetch loop has no dependent global load (see conv_tc.cu, kRowMode).
constexpr int kRowTab = 1024;

template <int kRowMode>
__global__ void __launch_bounds__(kThreads, 1)
conv1_wgrad_v2_kernel(const uint8_t* __restrict__ X, const int64_t* __restrict__ rows, const float* __restrict__ Out,
                      const float* __restrict__ Gr, float* __restrict__ partial, float* __restrict__ partial_bias, Geom g) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
    uint64_t* s_full = bars;
    uint64_t* s_empty = bars + kStages;
    uint64_t* acc_full = bars + 2 * kStages;
    uint64_t* acc_empty = acc_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    int64_t* row_tab = reinterpret_cast<int64_t*>(smem + kStages * kStageBytes + 256);
    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    constexpr uint32_t kIdesc = make_idesc_tf32(128, kN);

    const int64_t total_kb = static_cast<int64_t>(g.n_img) * g.GH;          // one K-block per grid row
    const int64_t per_cta = (total_kb + gridDim.x - 1) / gridDim.x;
    const int64_t kb_begin = static_cast<int64_t>(blockIdx.x) * per_cta;
    const int64_t kb_end = kb_begin + per_cta < total_kb ? kb_begin + per_cta : total_kb;
    const int64_t my_kb = kb_end > kb_begin ? kb_end - kb_begin : 0;
    const int64_t num_chunks = (my_kb + kPromote - 1) / kPromote;

    const int64_t n_lo = kb_begin / g.GH;                      // first sample this CTA touches
    if (kRowMode == 1) {
        for (int i = threadIdx.x; i < kRowTab; i += kThreads) row_tab[i] = (n_lo + i < g.n_img) ? rows[n_lo + i] : 0;
    }
    // zero the whole ring once: padding rows / chunks are never written afterwards
    for (int i = threadIdx.x; i < kStages * kStageBytes / 16; i += kThreads)
        reinterpret_cast<float4*>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&s_full[s], kProducerThreads / 32);
            mbar_init(&s_empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&acc_full[b], 1);
            mbar_init(&acc_empty[b], kDrainThreads / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kMmaWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = uniform_u32(*tmem_slot);

    if (warp < 8) {
        // ================================================================ producers
        const uint32_t ring = smem_u32(smem);
        const uint32_t row_tab_u32 = smem_u32(row_tab);
        const int tid = threadIdx.x;
        // A^T role: unit u = tid < 16 * n16: (c, ky', j16) -> one 16-byte load per K-block
        const bool a_thread = tid < 16 * g.n16;
        const int a_c = tid / (4 * g.n16), a_ky = (tid / g.n16) & 3, a_j = tid % g.n16;
        const int a_words = min(4, (g.W - a_j * 16) / 4);            // valid 4-byte words of this unit (tail unit: fewer)
        // G role: threads 96..255 (warps 3..7): unit (row = (by, oc), chunk j of 4 output columns); each unit
        // feeds BOTH tap columns: bx = 0 as one 16-byte store, bx = 1 as four 4-byte stores one element to the right
        const int gt = tid - 96;
        const int g_row = gt / g.nG, g_j = gt - g_row * g.nG;
        const bool g_thread = gt >= 0 && g_row < 32;
        const int g_oc = g_row & 15, g_by = (g_row >> 4) & 1;
        const int64_t img_bytes = static_cast<int64_t>(4) * g.H * g.W;
        const int P = g.OH * g.OW;
        const bool vec4_rows = (g.OW % 4 == 0) && (P % 4 == 0);
        float bias_acc = 0.0f;
        int s = 0;
        uint32_t ph = 0;

        struct HeldA { uint4 v; };
        struct HeldG { float4 gv, ov; };                              // raw gradient / activation: masked when stored
        auto fetch_a = [&](int64_t kb, HeldA& h) {
            h.v = make_uint4(0u, 0u, 0u, 0u);
            if (!a_thread) return;
            const int64_t n = kb / g.GH;
            const int Yg = static_cast<int>(kb - n * g.GH);
            int64_t img = n;
            if (kRowMode == 1) img = lds_s64(row_tab_u32 + static_cast<uint32_t>(n - n_lo) * 8u);
            const uint8_t* p = X + img * img_bytes + (static_cast<int64_t>(a_c) * g.H + 4 * Yg + a_ky) * g.W + a_j * 16;
            // input rows are only 4-byte aligned in general (W = 84): four word loads, not one 16-byte load
            const uint32_t* q = reinterpret_cast<const uint32_t*>(p);
            if (a_words > 0) h.v.x = q[0];
            if (a_words > 1) h.v.y = q[1];
            if (a_words > 2) h.v.z = q[2];
            if (a_words > 3) h.v.w = q[3];
        };
        auto fetch_g = [&](int64_t kb, HeldG& h) {
            h.gv = make_float4(0.f, 0.f, 0.f, 0.f);
            h.ov = make_float4(0.f, 0.f, 0.f, 0.f);                     // o = 0 -> masked out
            if (!g_thread) return;
            const int64_t n = kb / g.GH;
            const int oy = static_cast<int>(kb - n * g.GH) - g_by;
            if (oy < 0 || oy >= g.OH) return;
            const int64_t base = (n * 16 + g_oc) * P + static_cast<int64_t>(oy) * g.OW + 4 * g_j;
            if (vec4_rows) {                                              // OW % 4 == 0: 16-byte aligned, fully valid chunks
                h.gv = *reinterpret_cast<const float4*>(Gr + base);
                h.ov = *reinterpret_cast<const float4*>(Out + base);
            } else {
                const int left = g.OW - 4 * g_j;                          // valid columns in this chunk
                if (left > 0) { h.gv.x = Gr[base]; h.ov.x = Out[base]; }
                if (left > 1) { h.gv.y = Gr[base + 1]; h.ov.y = Out[base + 1]; }
                if (left > 2) { h.gv.z = Gr[base + 2]; h.ov.z = Out[base + 2]; }
                if (left > 3) { h.gv.w = Gr[base + 3]; h.ov.w = Out[base + 3]; }
            }
        };
        auto put = [&](const HeldA& ha, const HeldG& hg) {
            mbar_wait(&s_empty[s], ph ^ 1);
            const uint32_t st = ring + static_cast<uint32_t>(s * kStageBytes);
            if (a_thread) {
                // 16 bytes = 4 positions x 4 kx'; row ch = c*16 + ky'*4 + kx' gets chunk a_j = (pos0..pos3)[kx']
                const uint32_t w[4] = {ha.v.x, ha.v.y, ha.v.z, ha.v.w};
#pragma unroll
                for (int kx = 0; kx < 4; ++kx) {
                    const float4 f = make_float4(static_cast<float>((w[0] >> (8 * kx)) & 0xffu),
                                                 static_cast<float>((w[1] >> (8 * kx)) & 0xffu),
                                                 static_cast<float>((w[2] >> (8 * kx)) & 0xffu),
                                                 static_cast<float>((w[3] >> (8 * kx)) & 0xffu));
                    const int row = a_c * 16 + a_ky * 4 + kx;
                    sts128(st + static_cast<uint32_t>(row * 128 + ((a_j ^ (row & 7)) << 4)), f);
                }
            }
            if (g_thread) {
                const float e[4] = {hg.ov.x > 0.0f ? hg.gv.x : 0.0f, hg.ov.y > 0.0f ? hg.gv.y : 0.0f,
                                    hg.ov.z > 0.0f ? hg.gv.z : 0.0f, hg.ov.w > 0.0f ? hg.gv.w : 0.0f};
                if (g_by == 0) bias_acc += (e[0] + e[1]) + (e[2] + e[3]);
                float hi[4], lo[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) split_tf32(e[t], hi[t], lo[t]);
                // bx = 0: tile row (by*2 + 0)*16 + oc, elements 4j .. 4j+3
                const int row0 = (g_by * 2) * 16 + g_oc;
                const uint32_t off0 = static_cast<uint32_t>(kATile + row0 * 128 + ((g_j ^ (row0 & 7)) << 4));
                sts128(st + off0, make_float4(hi[0], hi[1], hi[2], hi[3]));
                sts128(st + off0 + kBTile, make_float4(lo[0], lo[1], lo[2], lo[3]));
                // bx = 1: tile row (by*2 + 1)*16 + oc, element X = 4j + t + 1 holds G[.., X - 1]
                const int row1 = row0 + 16;
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int xe = 4 * g_j + t + 1;
                    const uint32_t off1 = static_cast<uint32_t>(kATile + row1 * 128 + (((xe >> 2) ^ (row1 & 7)) << 4) + (xe & 3) * 4);
                    sts32(st + off1, hi[t]);
                    sts32(st + off1 + kBTile, lo[t]);
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_full[s]);
            if (++s == kStages) { s = 0; ph ^= 1; }
        };
        // software pipeline: A loads kDepthA K-blocks ahead, G rows kDepthG ahead (static register rings)
        HeldA ha[kDepthA];
        HeldG hg[kDepthG];
#pragma unroll
        for (int d = 0; d < kDepthA; ++d)
            if (kb_begin + d < kb_end) fetch_a(kb_begin + d, ha[d]);
#pragma unroll
        for (int d = 0; d < kDepthG; ++d)
            if (kb_begin + d < kb_end) fetch_g(kb_begin + d, hg[d]);
        constexpr int kUnroll = kDepthA * kDepthG;                    // 6: both ring indices static
        for (int64_t kb0 = kb_begin; kb0 < kb_end; kb0 += kUnroll) {
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const int64_t kb = kb0 + u;
                if (kb < kb_end) {
                    put(ha[u % kDepthA], hg[u % kDepthG]);
                    if (kb + kDepthA < kb_end) fetch_a(kb + kDepthA, ha[u % kDepthA]);
                    if (kb + kDepthG < kb_end) fetch_g(kb + kDepthG, hg[u % kDepthG]);
                }
            }
        }
        // bias gradient partials: one per (oc, chunk) unit of the by = 0 rows (zero where there is none)
        if (g_thread && g_by == 0) partial_bias[(blockIdx.x * 16 + g_oc) * 8 + g_j] = bias_acc;
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
                    const uint8_t* st = smem + s * kStageBytes;
                    const uint64_t da = make_desc(st), dbh = make_desc(st + kATile), dbl = make_desc(st + kATile + kBTile);
                    const uint32_t acc = tmem_base + static_cast<uint32_t>(buf * kN);
                    for (int k = 0; k < g.n_slices; ++k) {
                        const uint64_t adv = static_cast<uint64_t>(k * 2);
                        umma_tf32(acc, da + adv, dbh + adv, kIdesc, (kk > k0 || k > 0) ? 1u : 0u);
                        umma_tf32(acc, da + adv, dbl + adv, kIdesc, 1u);
                    }
                    umma_commit(&s_empty[s]);
                    if (kk == k1 - 1) umma_commit(&acc_full[buf]);
                }
                __syncwarp();
                if (++s == kStages) { s = 0; ph ^= 1; }
            }
        }
    } else {
        // ================================================================ drain (warps 8..11)
        const int q = warp - kDrainWarp0;
        const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
        float acc[kN];
#pragma unroll
        for (int j = 0; j < kN; ++j) acc[j] = 0.0f;
        for (int64_t c = 0; c < num_chunks; ++c) {
            const int buf = static_cast<int>(c & 1);
            mbar_wait(&acc_full[buf], (c >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t r[32];
                tmem_ld32(tmem_base + lane_base + static_cast<uint32_t>(buf * kN + h * 32), r);
#pragma unroll
                for (int j = 0; j < 32; ++j) acc[h * 32 + j] += __uint_as_float(r[j]);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);
        }
        if (q < 2) {                          // TMEM lanes 0..63 = the 64 channels; lanes 64..127 are the zero rows
            float* out = partial + (static_cast<int64_t>(blockIdx.x) * 64 + q * 32 + lane) * kN;
#pragma unroll
            for (int j = 0; j < kN; ++j) out[j] = acc[j];
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == kMmaWarp)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols));
}

// dW[oc][c][4*by + ky'][4*bx + kx'] = (1/255) * sum_cta partial[cta][ch = c*16 + ky'*4 + kx'][tap*16 + oc]
__global__ void wgrad_v2_reduce_kernel(const float* __restrict__ partial, const float* __restrict__ partial_bias,
                                       int nparts, float* __restrict__ dW, float* __restrict__ db) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;       // over ch*64 + col
    if (i < 64 * 64) {
        float s = 0.0f;
        for (int b = 0; b < nparts; ++b) s += partial[static_cast<int64_t>(b) * 4096 + i];
        const int ch = i >> 6, col = i & 63;
        const int tap = col >> 4, oc = col & 15, by = tap >> 1, bx = tap & 1;
        const int c = ch >> 4, kyp = (ch >> 2) & 3, kxp = ch & 3;
        dW[((oc * 4 + c) * 8 + 4 * by + kyp) * 8 + 4 * bx + kxp] = s * (1.0f / 255.0f);
    }
    if (i < 16) {
        float s = 0.0f;
        for (int b = 0; b < nparts; ++b)
            for (int j = 0; j < 8; ++j) s += partial_bias[(b * 16 + i) * 8 + j];
        db[i] = s;
    }
}

}  // namespace w2

int main(int argc, char** argv) {
    const int N = argc > 1 ? atoi(argv[1]) : 8192;
    const int H = 84, W = 84;
    w2::Geom g;
    g.n_img = N; g.H = H; g.W = W; g.GH = H / 4; g.GW = W / 4; g.OH = (H - 8) / 4 + 1; g.OW = (W - 8) / 4 + 1;
    g.n16 = (g.GW + 3) / 4; g.n_slices = (g.GW + 7) / 8; g.nG = (g.OW + 3) / 4;
    if (g.GW > 32 || g.nG > 5 || 16 * g.n16 > 96 || (W % 4) != 0) { printf("unsupported geometry\n"); return 1; }
    const int P = g.OH * g.OW;
    std::vector<uint8_t> hx(static_cast<size_t>(N) * 4 * H * W + 64);
    std::vector<float> hg(static_cast<size_t>(N) * 16 * P), ho(hg.size());
    uint32_t seed = 777u;
    auto rnd = [&]() { seed = seed * 1664525u + 1013904223u; return seed >> 8; };
    for (auto& v : hx) v = static_cast<uint8_t>(rnd() & 0xff);
    for (auto& v : hg) v = static_cast<float>(rnd() & 0xffff) / 32768.0f - 1.0f;
    for (auto& v : ho) v = static_cast<float>(rnd() & 0xffff) / 32768.0f - 1.0f;
    uint8_t* dx; float *dg, *dout, *dpart, *dpb, *dw, *db;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaMalloc(&dx, hx.size()); cudaMalloc(&dg, hg.size() * 4); cudaMalloc(&dout, ho.size() * 4);
    cudaMalloc(&dpart, static_cast<size_t>(sms) * 4096 * 4); cudaMalloc(&dpb, static_cast<size_t>(sms) * 16 * 8 * 4); cudaMemset(dpb, 0, static_cast<size_t>(sms) * 16 * 8 * 4);
    cudaMalloc(&dw, 1024 * 4); cudaMalloc(&db, 16 * 4);
    cudaMemcpy(dx, hx.data(), hx.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(dg, hg.data(), hg.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dout, ho.data(), ho.size() * 4, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(w2::conv1_wgrad_v2_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, w2::kSmemBytes);
    cudaFuncSetAttribute(w2::conv1_wgrad_v2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, w2::kSmemBytes);
    // minibatch row indices: a permutation of the images (sample n reads image hrows[n])
    std::vector<int64_t> hrows(N);
    for (int i = 0; i < N; ++i) hrows[i] = i;
    for (int i = N - 1; i > 0; --i) { const int j = rnd() % (i + 1); std::swap(hrows[i], hrows[j]); }
    int64_t* drows;
    cudaMalloc(&drows, static_cast<size_t>(N) * 8);
    cudaMemcpy(drows, hrows.data(), static_cast<size_t>(N) * 8, cudaMemcpyHostToDevice);
    int use_rows = 0;
    auto launch = [&](int n_img) {
        w2::Geom gg = g;
        gg.n_img = n_img;
        const int64_t total_kb = static_cast<int64_t>(n_img) * g.GH;
        const unsigned grid = static_cast<unsigned>(total_kb < sms ? total_kb : sms);
        const int64_t per_cta = (total_kb + grid - 1) / grid;
        if (per_cta / g.GH + 2 > w2::kRowTab) { printf("row table too small\n"); exit(1); }
        if (use_rows)
            w2::conv1_wgrad_v2_kernel<1><<<grid, w2::kThreads, w2::kSmemBytes>>>(dx, drows, dout, dg, dpart, dpb, gg);
        else
            w2::conv1_wgrad_v2_kernel<0><<<grid, w2::kThreads, w2::kSmemBytes>>>(dx, nullptr, dout, dg, dpart, dpb, gg);
        w2::wgrad_v2_reduce_kernel<<<16, 256>>>(dpart, dpb, static_cast<int>(grid), dw, db);
    };
    cudaEvent_t e0, e1;
    for (use_rows = 0; use_rows < 2; ++use_rows) {
    // ---- correctness on the first n_chk images against fp64 loops
    const int n_chk = N < 48 ? N : 48;
    launch(n_chk);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<float> gw(1024), gb(16);
    cudaMemcpy(gw.data(), dw, 1024 * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(gb.data(), db, 16 * 4, cudaMemcpyDeviceToHost);
    std::vector<double> rw(1024, 0.0), rws(1024, 0.0), rb(16, 0.0), rbs(16, 0.0);
    for (int n = 0; n < n_chk; ++n)
        for (int oc = 0; oc < 16; ++oc)
            for (int oy = 0; oy < g.OH; ++oy)
                for (int ox = 0; ox < g.OW; ++ox) {
                    const size_t gi = ((static_cast<size_t>(n) * 16 + oc) * g.OH + oy) * g.OW + ox;
                    if (!(ho[gi] > 0.0f)) continue;
                    const double gv = hg[gi];
                    rb[oc] += gv; rbs[oc] += std::fabs(gv);
                    for (int c = 0; c < 4; ++c)
                        for (int ky = 0; ky < 8; ++ky)
                            for (int kx = 0; kx < 8; ++kx) {
                                const double a = hx[((static_cast<size_t>(use_rows ? hrows[n] : n) * 4 + c) * H + 4 * oy + ky) * W + 4 * ox + kx] / 255.0;
                                rw[((oc * 4 + c) * 8 + ky) * 8 + kx] += a * gv;
                                rws[((oc * 4 + c) * 8 + ky) * 8 + kx] += std::fabs(a * gv);
                            }
                }
    double worst = 0.0;
    for (int i = 0; i < 1024; ++i) worst = std::fmax(worst, std::fabs(gw[i] - rw[i]) / (rws[i] + 1e-30));
    double worst_b = 0.0;
    for (int i = 0; i < 16; ++i) worst_b = std::fmax(worst_b, std::fabs(gb[i] - rb[i]) / (rbs[i] + 1e-30));
    printf("rows=%d n_chk=%d  dW max err / term scale = %.3e, db = %.3e -> %s\n", use_rows, n_chk, worst, worst_b,
           (worst <= 1e-5 && worst_b <= 1e-5) ? "OK" : "MISMATCH");
    // ---- timing at full size
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) launch(N);
    cudaEventRecord(e0);
    for (int i = 0; i < 10; ++i) launch(N);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    e = cudaGetLastError();
    printf("conv1 wgrad v2: %.1f us per launch at N=%d (v1 tcgen05 kernel: ~780 us at N=8192) %s\n", ms * 100.0f, N,
           e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
    return 0;
}
