// Round-2 experiment 1: "v2" first conv layer WITHOUT im2col expansion.
//
//   Conv2d(4->16, k8, s4) on uint8 frames = a 2x2 stride-1 convolution over the space-to-depth input
//   X4[g, (c,ky',kx')] = x[n, c, 4Y+ky', 4X+kx'],  g = n*GH*GW + Y*GW + X  (GH = H/4, GW = W/4).
//   Y[g, oc] = sum over taps (by,bx) of  X4[g + by*GW + bx, :] . W4[by,bx][oc, :]
//   (tests/test_conv_layout_math.py::test_space_to_depth_shifted_gemm pins this identity).
//
// One shared-memory stage holds the X4 rows [g0, g0 + 160) of a 128-row tile as two K-major
// SWIZZLE_128B k-block tiles (32 of the 64 channels each); the four taps are FOUR ROW-SHIFTED
// DESCRIPTORS into the same stage (shift = by*GW + bx rows), so each input byte is loaded, converted
// and stored once (2400 chunks per tile instead of 8192).  Whether a descriptor may start at a row
// that is not a multiple of 8, and with which base-offset field, is what tcgen05_shift_probe.cu
// answers; `bo_mode` selects the encoding here (0: base_offset = 0, 1: (addr >> 7) & 7).
//
// Standalone: builds with nvcc only, checks a sample of outputs against a CPU fp64 convolution and
// times the kernel with CUDA events:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I rlpyt_b200/csrc \
//        -o tools/probes/_bin/conv1_v2 tools/probes/conv1_v2_probe.cu && tools/probes/_bin/conv1_v2 [N] [bo_mode]
// NOT validated: written at the end of round 1 after the GPU budget was spent.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "tc_common.cuh"

using namespace rl::tc;

namespace v2 {

constexpr int kRows = 128;                  // output rows (grid positions) per tile
constexpr int kStageRows = 160;             // X4 rows staged per tile: 128 + GW + 1 <= 160  (GW <= 31)
constexpr int kKbTile = kStageRows * 128;   // one k-block tile: 160 rows x 128 B = 20 KiB
constexpr int kStageBytes = 2 * kKbTile;    // channels 0..31 | 32..63
constexpr int kStages = 3;
constexpr int kN = 16;                      // output channels
constexpr int kBTile = kN * 128;            // one [16 oc x 32 ch] K-major tile
constexpr int kBBytes = 4 * 2 * kBTile;     // 4 taps x 2 k-blocks, one term (hi or lo)
constexpr int kThreads = 416;               // warps 0-7 producers, 8-11 epilogue, 12 MMA
constexpr int kProducerThreads = 256, kEpilogueThreads = 128, kMmaWarp = 12, kEpiWarp0 = 8;
constexpr int kTmemCols = 64;               // 2 buffers x 2 accumulators (by = 0 / 1) x 16 columns
constexpr int kItems = kStageRows * 16 / kProducerThreads;   // 10 chunks per producer thread per tile
constexpr int kSmemBytes = 2 * kBBytes + kStages * kStageBytes + 2 * kStageRows * 8 + 256 + 1024;

struct Geom {
    int n_img, H, W, GH, GW, OH, OW;
    int64_t g_total;                        // n_img * GH * GW
};

__global__ void __launch_bounds__(kThreads, 1)
conv1_v2_kernel(const uint8_t* __restrict__ X, const int64_t* __restrict__ rows, const float* __restrict__ Wg,
                const float* __restrict__ bias, float* __restrict__ Y, Geom g, int relu, int bo_mode) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* b_hi = smem;
    uint8_t* b_lo = smem + kBBytes;
    uint8_t* a_ring = smem + 2 * kBBytes;
    int64_t* row_off = reinterpret_cast<int64_t*>(a_ring + kStages * kStageBytes);   // [2][kStageRows] byte offsets
    uint64_t* bars = reinterpret_cast<uint64_t*>(row_off + 2 * kStageRows);
    uint64_t* a_full = bars;                 // [kStages] producers -> MMA
    uint64_t* a_empty = bars + kStages;      // [kStages] MMA -> producers
    uint64_t* acc_full = bars + 2 * kStages; // [2]
    uint64_t* acc_empty = acc_full + 2;      // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
    const int64_t num_tiles = (g.g_total + kRows - 1) / kRows;
    const int G = g.GH * g.GW;
    constexpr uint32_t kIdesc = make_idesc_tf32(kRows, kN);

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&a_full[s], kProducerThreads / 32);
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
    // filter bank: W4[(by,bx)][oc][ch = c*16 + ky'*4 + kx'] = w[oc, c, 4*by + ky', 4*bx + kx'], split hi/lo,
    // as 4 taps x 2 k-blocks of [16 oc x 32 ch] K-major SWIZZLE_128B tiles
    for (int idx = threadIdx.x; idx < 4 * 2 * kN * 8; idx += kThreads) {
        const int j = idx & 7, oc = (idx >> 3) & 15, kb = (idx >> 7) & 1, tap = idx >> 8;
        const int by = tap >> 1, bx = tap & 1;
        const int ch0 = kb * 32 + j * 4;                       // 4 consecutive channels = 4 consecutive kx'
        const int c = ch0 >> 4, kyp = (ch0 >> 2) & 3;
        const float* wp = Wg + ((oc * 4 + c) * 8 + (4 * by + kyp)) * 8 + 4 * bx;
        float4 hi, lo;
        split_tf32(wp[0], hi.x, lo.x); split_tf32(wp[1], hi.y, lo.y);
        split_tf32(wp[2], hi.z, lo.z); split_tf32(wp[3], hi.w, lo.w);
        const int off = (tap * 2 + kb) * kBTile + oc * 128 + ((j ^ (oc & 7)) << 4);
        *reinterpret_cast<float4*>(b_hi + off) = hi;
        *reinterpret_cast<float4*>(b_lo + off) = lo;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = uniform_u32(*tmem_slot);

    if (warp < 8) {
        // ================================================================ producers
        // item i of a tile: X4 row p = i % 160 (lanes run over consecutive grid positions = consecutive
        // 4-byte words of one input row), chunk q = i / 160 = c*4 + ky' (k-block q >> 3, chunk q & 7)
        const uint32_t ring_u32 = smem_u32(a_ring);
        const uint32_t row_off_u32 = smem_u32(row_off);
        const int64_t img_bytes = static_cast<int64_t>(4) * g.H * g.W;
        int s = 0;
        uint32_t ph = 0;
        // byte offset of x[n, 0, 4Y, 4X] for stage row p of `tile`, or -1 past the end (one decode per row per tile)
        auto fill_row_table = [&](int64_t tile, int buf) {
            if (threadIdx.x < kStageRows) {
                const int64_t gg = tile * kRows + threadIdx.x;
                int64_t off = -1;
                if (gg < g.g_total) {
                    const int64_t n = gg / G;
                    const int pos = static_cast<int>(gg - n * G);
                    const int Yg = pos / g.GW, Xg = pos - Yg * g.GW;
                    const int64_t img = rows != nullptr ? rows[n] : n;
                    off = img * img_bytes + static_cast<int64_t>(4 * Yg) * g.W + 4 * Xg;
                }
                row_off[buf * kStageRows + threadIdx.x] = off;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(kProducerThreads) : "memory");   // producers only
        };
        auto fetch = [&](int buf, uint32_t (&v)[kItems]) {
#pragma unroll
            for (int i = 0; i < kItems; ++i) {
                const int item = threadIdx.x + i * kProducerThreads;
                const int p = item % kStageRows, q = item / kStageRows;
                const int64_t off = lds_s64(row_off_u32 + static_cast<uint32_t>(buf * kStageRows + p) * 8u);
                const int c = q >> 2, kyp = q & 3;
                v[i] = 0u;
                if (off >= 0) v[i] = *reinterpret_cast<const uint32_t*>(X + off + (static_cast<int64_t>(c) * g.H + kyp) * g.W);
            }
        };
        auto put = [&](const uint32_t (&v)[kItems]) {
            mbar_wait(&a_empty[s], ph ^ 1);
            const uint32_t st = ring_u32 + static_cast<uint32_t>(s * kStageBytes);
#pragma unroll
            for (int i = 0; i < kItems; ++i) {
                const int item = threadIdx.x + i * kProducerThreads;
                const int p = item % kStageRows, q = item / kStageRows;
                const uint32_t w = v[i];
                const float4 f = make_float4(static_cast<float>(w & 0xffu), static_cast<float>((w >> 8) & 0xffu),
                                             static_cast<float>((w >> 16) & 0xffu), static_cast<float>(w >> 24));
                sts128(st + static_cast<uint32_t>((q >> 3) * kKbTile + p * 128 + (((q & 7) ^ (p & 7)) << 4)), f);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&a_full[s]);
            if (++s == kStages) { s = 0; ph ^= 1; }
        };
        uint32_t cur[kItems], nxt[kItems];
        int64_t tile = blockIdx.x;
        int buf = 0;
        if (tile < num_tiles) {
            fill_row_table(tile, buf);
            fetch(buf, cur);
        }
        while (tile < num_tiles) {
            const int64_t tile_next = tile + gridDim.x;
            if (tile_next < num_tiles) {
                fill_row_table(tile_next, buf ^ 1);
                fetch(buf ^ 1, nxt);
            }
            put(cur);
#pragma unroll
            for (int i = 0; i < kItems; ++i) cur[i] = nxt[i];
            tile = tile_next;
            buf ^= 1;
        }
    } else if (warp == kMmaWarp) {
        // ================================================================ MMA issuer
        int s = 0;
        uint32_t ph = 0, tcount = 0;
        for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tcount) {
            const int buf = tcount & 1;
            mbar_wait(&acc_empty[buf], ((tcount >> 1) & 1) ^ 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            mbar_wait(&a_full[s], ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (elect_one()) {
                const uint8_t* st = a_ring + s * kStageBytes;
#pragma unroll
                for (int tap = 0; tap < 4; ++tap) {
                    const int shift = (tap >> 1) * g.GW + (tap & 1);
                    const uint32_t acc = tmem_base + static_cast<uint32_t>(buf * 2 * kN + (tap >> 1) * kN);
#pragma unroll
                    for (int kb = 0; kb < 2; ++kb) {
                        const uint8_t* a_src = st + kb * kKbTile + shift * 128;
                        uint64_t da = make_desc(a_src);
                        if (bo_mode == 1) da |= static_cast<uint64_t>((smem_u32(a_src) >> 7) & 7u) << 49;
                        const uint64_t dbh = make_desc(b_hi + (tap * 2 + kb) * kBTile);
                        const uint64_t dbl = make_desc(b_lo + (tap * 2 + kb) * kBTile);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint64_t adv = static_cast<uint64_t>(k * 2);
                            const uint32_t first = ((tap & 1) == 0 && kb == 0 && k == 0) ? 0u : 1u;   // per accumulator
                            umma_tf32(acc, da + adv, dbh + adv, kIdesc, first);
                            umma_tf32(acc, da + adv, dbl + adv, kIdesc, 1u);
                        }
                    }
                }
                umma_commit(&a_empty[s]);
                umma_commit(&acc_full[buf]);
            }
            __syncwarp();
            if (++s == kStages) { s = 0; ph ^= 1; }
        }
    } else {
        // ================================================================ epilogue (warps 8..11)
        const int q = warp - kEpiWarp0;
        const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
        uint32_t tcount = 0;
        for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tcount) {
            const int buf = tcount & 1;
            mbar_wait(&acc_full[buf], (tcount >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t r[32];
            tmem_ld32(tmem_base + lane_base + static_cast<uint32_t>(buf * 2 * kN), r);   // acc(by=0) | acc(by=1)
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);
            const int64_t gg = tile * kRows + q * 32 + lane;
            if (gg < g.g_total) {
                const int64_t n = gg / G;
                const int pos = static_cast<int>(gg - n * G);
                const int Yg = pos / g.GW, Xg = pos - Yg * g.GW;
                if (Yg < g.OH && Xg < g.OW) {
                    float* yo = Y + (n * kN * g.OH + Yg) * g.OW + Xg;
#pragma unroll
                    for (int oc = 0; oc < kN; ++oc) {
                        float v = (__uint_as_float(r[oc]) + __uint_as_float(r[kN + oc])) * (1.0f / 255.0f) + bias[oc];
                        if (relu) v = fmaxf(v, 0.0f);
                        yo[static_cast<int64_t>(oc) * g.OH * g.OW] = v;
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

}  // namespace v2

int main(int argc, char** argv) {
    const int N = argc > 1 ? atoi(argv[1]) : 8192;
    const int bo_mode = argc > 2 ? atoi(argv[2]) : 0;
    const int H = 84, W = 84;
    v2::Geom g;
    g.n_img = N; g.H = H; g.W = W; g.GH = H / 4; g.GW = W / 4; g.OH = (H - 8) / 4 + 1; g.OW = (W - 8) / 4 + 1;
    g.g_total = static_cast<int64_t>(N) * g.GH * g.GW;
    if (g.GW + 1 + v2::kRows > v2::kStageRows) { printf("GW too large for kStageRows\n"); return 1; }
    std::vector<uint8_t> hx(static_cast<size_t>(N) * 4 * H * W);
    std::vector<float> hw(16 * 4 * 8 * 8), hb(16);
    uint32_t seed = 12345u;
    auto rnd = [&]() { seed = seed * 1664525u + 1013904223u; return seed >> 8; };
    for (auto& v : hx) v = static_cast<uint8_t>(rnd() & 0xff);
    for (auto& v : hw) v = (static_cast<float>(rnd() & 0xffff) / 65536.0f - 0.5f) / 8.0f;
    for (auto& v : hb) v = (static_cast<float>(rnd() & 0xffff) / 65536.0f - 0.5f) / 8.0f;
    uint8_t* dx; float *dw, *db, *dy;
    const size_t ybytes = static_cast<size_t>(N) * 16 * g.OH * g.OW * sizeof(float);
    cudaMalloc(&dx, hx.size()); cudaMalloc(&dw, hw.size() * 4); cudaMalloc(&db, hb.size() * 4); cudaMalloc(&dy, ybytes);
    cudaMemcpy(dx, hx.data(), hx.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(dw, hw.data(), hw.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(db, hb.data(), hb.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dy, 0xff, ybytes);
    cudaFuncSetAttribute(v2::conv1_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, v2::kSmemBytes);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t tiles = (g.g_total + v2::kRows - 1) / v2::kRows;
    const unsigned grid = static_cast<unsigned>(tiles < sms ? tiles : sms);
    auto launch = [&]() { v2::conv1_v2_kernel<<<grid, v2::kThreads, v2::kSmemBytes>>>(dx, nullptr, dw, db, dy, g, 1, bo_mode); };
    launch();
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    // ---- check a sample of outputs against a direct fp64 convolution
    std::vector<float> hy(static_cast<size_t>(N) * 16 * g.OH * g.OW);
    cudaMemcpy(hy.data(), dy, ybytes, cudaMemcpyDeviceToHost);
    double max_err = 0.0, max_scale = 0.0;
    int checked = 0;
    for (int t = 0; t < 20000; ++t) {
        const int n = rnd() % N, oc = rnd() % 16, oy = rnd() % g.OH, ox = rnd() % g.OW;
        double acc = 0.0, scale = 0.0;
        for (int c = 0; c < 4; ++c)
            for (int ky = 0; ky < 8; ++ky)
                for (int kx = 0; kx < 8; ++kx) {
                    const double a = hx[((static_cast<size_t>(n) * 4 + c) * H + 4 * oy + ky) * W + 4 * ox + kx] / 255.0;
                    const double w = hw[((oc * 4 + c) * 8 + ky) * 8 + kx];
                    acc += a * w;
                    scale += std::fabs(a * w);
                }
        acc += hb[oc];
        const double want = acc > 0 ? acc : 0;
        const double got = hy[((static_cast<size_t>(n) * 16 + oc) * g.OH + oy) * g.OW + ox];
        max_err = std::fmax(max_err, std::fabs(got - want));
        max_scale = std::fmax(max_scale, scale);
        ++checked;
    }
    printf("N=%d bo_mode=%d checked=%d max_abs_err=%.3e (term scale %.3f) -> %s\n", N, bo_mode, checked, max_err, max_scale,
           max_err <= 3e-6 * max_scale ? "OK" : "MISMATCH");
    // ---- timing
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) launch();
    cudaEventRecord(e0);
    for (int i = 0; i < 10; ++i) launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("conv1_v2 forward: %.1f us per launch (v1 tcgen05 kernel: ~450 us at N=8192)\n", ms * 100.0f);
    return 0;
}
