// Round-2 probe: (1) tcgen05.mma kind::i8 semantics this repo relies on - u8 x s8 -> s32, K-major
// SWIZZLE_128B operands, row-shifted A descriptors, MN-major A (the weight-gradient kernel's operand) with
// a K-direction shift; (2) the kind::i8 first-layer forward kernel of rlpyt_b200/csrc/conv1_i8.cuh against
// an fp64 convolution on the host, and its time per launch.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I rlpyt_b200/csrc \
//        -o tools/probes/_bin/conv1_i8 tools/probes/conv1_i8_probe.cu && tools/probes/_bin/conv1_i8
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "conv1_i8.cuh"

using namespace rl::tc;
using namespace rl::c1i8;

static uint32_t g_seed = 12345u;
static uint32_t rnd() { g_seed = g_seed * 1664525u + 1013904223u; return g_seed >> 8; }

// ---------------------------------------------------------------------------------------------- (1)
// One CTA copies a host-built shared-memory image, issues n_mma MMAs (M=128, N=64, K=32) whose
// descriptors are (a_desc + i*a_step, b_desc + i*b_step) with start addresses relative to the
// 1024-aligned base, and dumps D[128][64].
__global__ void __launch_bounds__(128, 1)
i8_mma_probe_kernel(const uint8_t* __restrict__ img, int img_bytes, uint64_t a_desc, uint64_t b_desc, int a_step,
                    int b_step, int n_mma, uint32_t idesc, int* __restrict__ out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + img_bytes);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < img_bytes; i += blockDim.x) smem[i] = img[i];
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(64));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    if (warp == 0 && lane == 0) {
        const uint64_t base16 = static_cast<uint64_t>((smem_u32(smem) & 0x3FFFFu) >> 4);
        for (int i = 0; i < n_mma; ++i)
            umma_i8(tmem, a_desc + base16 + static_cast<uint64_t>(i * a_step), b_desc + base16 + static_cast<uint64_t>(i * b_step),
                    idesc, i > 0 ? 1u : 0u);
        umma_commit(bar);
    }
    mbar_wait(bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t r0[32], r1[32];
    tmem_ld32(tmem + (static_cast<uint32_t>(warp * 32) << 16), r0);
    tmem_ld32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + 32, r1);
    for (int n = 0; n < 32; ++n) {
        out[(warp * 32 + lane) * 64 + n] = static_cast<int>(r0[n]);
        out[(warp * 32 + lane) * 64 + 32 + n] = static_cast<int>(r1[n]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(64));
}

static uint64_t desc_hi(int lbo16, int sbo16) {      // SWIZZLE_128B, version 1
    return (static_cast<uint64_t>(lbo16) << 16) | (static_cast<uint64_t>(sbo16) << 32) | (1ull << 46) | (2ull << 61);
}

static int run_mma_case(const char* name, const std::vector<uint8_t>& img, uint64_t a_desc, uint64_t b_desc, int a_step,
                        int b_step, int n_mma, uint32_t idesc, const std::vector<long long>& want) {
    uint8_t* d_img; int* d_out;
    cudaMalloc(&d_img, img.size()); cudaMalloc(&d_out, 128 * 64 * 4);
    cudaMemcpy(d_img, img.data(), img.size(), cudaMemcpyHostToDevice);
    cudaMemset(d_out, 0xff, 128 * 64 * 4);
    const int smem_bytes = static_cast<int>(img.size()) + 64 + 1024;
    cudaFuncSetAttribute(i8_mma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    i8_mma_probe_kernel<<<1, 128, smem_bytes>>>(d_img, static_cast<int>(img.size()), a_desc, b_desc, a_step, b_step, n_mma, idesc, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("[mma] %-44s CUDA error: %s\n", name, cudaGetErrorString(e)); return -1; }
    std::vector<int> h(128 * 64);
    cudaMemcpy(h.data(), d_out, h.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0, bi = -1;
    for (int i = 0; i < 128 * 64; ++i)
        if (h[i] != want[i] && bad++ == 0) bi = i;
    if (bad) printf("[mma] %-44s mismatches %d first (m=%d,n=%d) got %d want %lld\n", name, bad, bi / 64, bi % 64, h[bi], want[bi]);
    else printf("[mma] %-44s OK\n", name);
    cudaFree(d_img); cudaFree(d_out);
    return bad;
}

static void mma_probes() {
    // B: K-major [64 rows][128 B] at offset b_off; A region before it
    const int a_rows = 192;                                  // K-major rows / MN-major k-rows
    const int a_bytes = a_rows * 128, b_off = a_bytes, b_bytes = 64 * 128;
    std::vector<int> Bl(64 * 128);
    for (auto& v : Bl) v = static_cast<int>(rnd() % 255) - 127;            // s8
    auto put_b = [&](std::vector<uint8_t>& img) {
        for (int n = 0; n < 64; ++n)
            for (int k = 0; k < 128; ++k)
                img[b_off + n * 128 + (((k >> 4) ^ (n & 7)) << 4) + (k & 15)] = static_cast<uint8_t>(static_cast<int8_t>(Bl[n * 128 + k]));
    };
    const uint32_t idesc_k = make_idesc_i8(128, 64);
    const uint32_t idesc_mn = idesc_k | (1u << 15);          // a_major = MN
    {   // K-major A, rows = GEMM rows
        std::vector<int> Al(a_rows * 128);
        for (auto& v : Al) v = static_cast<int>(rnd() & 255);              // u8 (values above 127 test signedness)
        std::vector<uint8_t> img(a_bytes + b_bytes);
        for (int r = 0; r < a_rows; ++r)
            for (int k = 0; k < 128; ++k) img[r * 128 + (((k >> 4) ^ (r & 7)) << 4) + (k & 15)] = static_cast<uint8_t>(Al[r * 128 + k]);
        put_b(img);
        for (int shift : {0, 1, 21, 22, 20}) {
            std::vector<long long> want(128 * 64);
            for (int m = 0; m < 128; ++m)
                for (int n = 0; n < 64; ++n) {
                    long long s = 0;
                    for (int k = 0; k < 128; ++k) s += static_cast<long long>(Al[(m + shift) * 128 + k]) * Bl[n * 128 + k];
                    want[m * 64 + n] = s;
                }
            char name[96];
            snprintf(name, sizeof name, "K-major A u8 x s8, K=128, row shift %d", shift);
            run_mma_case(name, img, desc_hi(1, 64) + static_cast<uint64_t>(shift * 8), desc_hi(1, 64) + static_cast<uint64_t>(b_off >> 4), 2, 2, 4,
                         idesc_k, want);
        }
    }
    {   // MN-major A: logical A[m][k], k-row of 128 B (m contiguous), 8 k-rows per 1024 B atom; B K-major, K = cells
        std::vector<int> Al(128 * a_rows);                   // [m][k]
        for (auto& v : Al) v = static_cast<int>(rnd() & 255);
        std::vector<uint8_t> img(a_bytes + b_bytes);
        for (int k = 0; k < a_rows; ++k)
            for (int m = 0; m < 128; ++m) img[k * 128 + (((m >> 4) ^ (k & 7)) << 4) + (m & 15)] = static_cast<uint8_t>(Al[m * a_rows + k]);
        put_b(img);
        for (int lbo : {1, 64}) {
            for (int shift : {0, 8, 21, 20}) {
                std::vector<long long> want(128 * 64);
                for (int m = 0; m < 128; ++m)
                    for (int n = 0; n < 64; ++n) {
                        long long s = 0;
                        for (int k = 0; k < 128; ++k) s += static_cast<long long>(Al[m * a_rows + k + shift]) * Bl[n * 128 + k];
                        want[m * 64 + n] = s;
                    }
                char name[96];
                snprintf(name, sizeof name, "MN-major A (lbo16=%d), K=128 cells, K shift %d", lbo, shift);
                // per MMA (K = 32): A advances 32 k-rows = 4096 B (256 x 16 B), B advances 32 B
                run_mma_case(name, img, desc_hi(lbo, 64) + static_cast<uint64_t>(shift * 8), desc_hi(1, 64) + static_cast<uint64_t>(b_off >> 4), 256,
                             2, 4, idesc_mn, want);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------- (2)
static int conv_case(int N, int H, int W, bool use_rows, bool timing, int dbg = 0) {
    if (!geom_ok(4, H, W)) { printf("geometry not supported\n"); return 1; }
    Geom g = make_geom(N, H, W);
    const int R = use_rows ? N + 37 : N;                      // frames in the store; rows pick N of them
    std::vector<uint8_t> hx(static_cast<size_t>(R) * 4 * H * W);
    std::vector<float> hw(16 * 4 * 8 * 8), hb(16);
    std::vector<int64_t> hrows(N);
    for (auto& v : hx) v = static_cast<uint8_t>(rnd() & 0xff);
    for (int i = 0; i < 16 * 256; ++i) {
        float w = (static_cast<float>(rnd() & 0xffff) / 65536.0f - 0.5f) / 8.0f;
        if ((i & 31) == 7) w *= 1e-3f;                        // small weights: exercise the low digits
        if ((i >> 8) == 3) w *= 37.0f;                        // a channel with a different scale
        hw[i] = w;
    }
    for (auto& v : hb) v = (static_cast<float>(rnd() & 0xffff) / 65536.0f - 0.5f) / 8.0f;
    for (auto& v : hrows) v = static_cast<int64_t>(rnd() % R);
    uint8_t* dx; float *dw, *db, *dy; int64_t* drows;
    const int P = g.OH * g.OW;
    const size_t ybytes = static_cast<size_t>(N) * 16 * P * sizeof(float);
    cudaMalloc(&dx, hx.size()); cudaMalloc(&dw, hw.size() * 4); cudaMalloc(&db, hb.size() * 4); cudaMalloc(&dy, ybytes);
    cudaMalloc(&drows, hrows.size() * 8);
    cudaMemcpy(dx, hx.data(), hx.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(dw, hw.data(), hw.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(db, hb.data(), hb.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(drows, hrows.data(), hrows.size() * 8, cudaMemcpyHostToDevice);
    cudaMemset(dy, 0xff, ybytes);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    auto launch = [&]() { return launch_fwd(dx, use_rows ? drows : nullptr, dw, db, dy, g, 1, sms, 0); };
    cudaError_t e = launch();
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("[conv] N=%d %dx%d CUDA error: %s\n", N, H, W, cudaGetErrorString(e)); return 1; }
    std::vector<float> hy(static_cast<size_t>(N) * 16 * P);
    cudaMemcpy(hy.data(), dy, ybytes, cudaMemcpyDeviceToHost);
    double max_err = 0.0, max_scale = 0.0, max_ratio = 0.0;
    const int checks = N * 16 * P < 40000 ? N * 16 * P : 40000;
    for (int t = 0; t < checks; ++t) {
        int n, oc, oy, ox;
        if (checks == N * 16 * P) { n = t / (16 * P); oc = (t / P) % 16; oy = (t % P) / g.OW; ox = t % g.OW; }
        else { n = rnd() % N; oc = rnd() % 16; oy = rnd() % g.OH; ox = rnd() % g.OW; }
        const size_t img = use_rows ? static_cast<size_t>(hrows[n]) : static_cast<size_t>(n);
        double acc = 0.0, scale = 0.0;
        for (int c = 0; c < 4; ++c)
            for (int ky = 0; ky < 8; ++ky)
                for (int kx = 0; kx < 8; ++kx) {
                    const double a = hx[((img * 4 + c) * H + 4 * oy + ky) * W + 4 * ox + kx] / 255.0;
                    const double w = hw[((oc * 4 + c) * 8 + ky) * 8 + kx];
                    acc += a * w;
                    scale += std::fabs(a * w);
                }
        acc += hb[oc];
        const double want = acc > 0 ? acc : 0;
        const double got = hy[((static_cast<size_t>(n) * 16 + oc) * g.OH + oy) * g.OW + ox];
        const double err = std::fabs(got - want);
        max_err = std::fmax(max_err, err);
        max_scale = std::fmax(max_scale, scale);
        max_ratio = std::fmax(max_ratio, err / scale);
    }
    printf("[conv] N=%d %dx%d rows=%d checked=%d max_abs_err=%.3e max err/sum|x||w|=%.3e -> %s\n", N, H, W, use_rows ? 1 : 0, checks,
           max_err, max_ratio, max_ratio <= 3e-6 ? "OK" : "MISMATCH");
    if (timing) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int i = 0; i < 3; ++i) launch();
        cudaEventRecord(e0);
        for (int i = 0; i < 20; ++i) launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double us = ms * 1000.0 / 20.0;
        const double bytes = static_cast<double>(N) * (4.0 * H * W + 16.0 * P * 4.0);
        printf("[conv] dbg=%d N=%d %dx%d: %.1f us per launch, %.0f GB/s of algorithmic bytes (v1 tcgen05 tf32 kernel: ~450 us at N=8192 84x84)\n",
               dbg, N, H, W, us, bytes / us * 1e-3);
    }
    cudaFree(dx); cudaFree(dw); cudaFree(db); cudaFree(dy); cudaFree(drows);
    return 0;
}

// ---------------------------------------------------------------------------------------------- (3)
static int wgrad_case(int N, int H, int W, bool use_rows, bool use_mask, bool check, bool timing) {
    if (!geom_ok(4, H, W)) { printf("geometry not supported\n"); return 1; }
    Geom g = make_geom(N, H, W);
    if (!wg::smem_ok(g)) { printf("[wgrad] %dx%d: shared memory does not fit\n", H, W); return 1; }
    const int R = use_rows ? N + 37 : N, P = g.OH * g.OW;
    std::vector<uint8_t> hx(static_cast<size_t>(R) * 4 * H * W);
    std::vector<float> hg(static_cast<size_t>(N) * 16 * P), ho(static_cast<size_t>(N) * 16 * P);
    std::vector<int64_t> hrows(N);
    for (auto& v : hx) v = static_cast<uint8_t>(rnd() & 0xff);
    for (size_t i = 0; i < hg.size(); ++i) {
        float v = (static_cast<float>(rnd() & 0xffff) / 65536.0f - 0.5f) * 1e-3f;
        if ((rnd() & 63) == 0) v *= 300.0f;                   // heavy tail: a few large gradients set the scale
        if (((i / P) % 16) == 5) v *= 1e-4f;                  // a channel with tiny gradients
        hg[i] = v;
        ho[i] = (rnd() & 3) ? 1.0f : 0.0f;                    // ReLU mask source: 1/4 of the outputs are off
    }
    for (auto& v : hrows) v = static_cast<int64_t>(rnd() % R);
    uint8_t* dx; float *dg, *dout, *dw, *db; int64_t* drows; void* scratch;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaMalloc(&dx, hx.size()); cudaMalloc(&dg, hg.size() * 4); cudaMalloc(&dout, ho.size() * 4);
    cudaMalloc(&dw, 4096 * 4); cudaMalloc(&db, 16 * 4); cudaMalloc(&drows, hrows.size() * 8);
    cudaMalloc(&scratch, wg::scratch_bytes(sms));
    cudaMemcpy(dx, hx.data(), hx.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(dg, hg.data(), hg.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dout, ho.data(), ho.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(drows, hrows.data(), hrows.size() * 8, cudaMemcpyHostToDevice);
    auto launch = [&]() { return wg::launch_wgrad(dx, use_rows ? drows : nullptr, use_mask ? dout : nullptr, dg, dw, db, g, sms, scratch, 0); };
    cudaError_t e = launch();
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("[wgrad] N=%d %dx%d CUDA error: %s\n", N, H, W, cudaGetErrorString(e)); return 1; }
    if (check) {
        std::vector<float> gw(4096), gb(16);
        cudaMemcpy(gw.data(), dw, 4096 * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(gb.data(), db, 16 * 4, cudaMemcpyDeviceToHost);
        std::vector<double> ref(4096, 0.0), scale(4096, 0.0), refb(16, 0.0);
        for (int n = 0; n < N; ++n) {
            const size_t img = use_rows ? static_cast<size_t>(hrows[n]) : static_cast<size_t>(n);
            for (int oc = 0; oc < 16; ++oc)
                for (int oy = 0; oy < g.OH; ++oy)
                    for (int ox = 0; ox < g.OW; ++ox) {
                        const size_t gi = (static_cast<size_t>(n) * 16 + oc) * P + oy * g.OW + ox;
                        const double gv = (!use_mask || ho[gi] > 0.0f) ? hg[gi] : 0.0;
                        if (gv == 0.0) continue;
                        refb[oc] += gv;
                        for (int c = 0; c < 4; ++c)
                            for (int ky = 0; ky < 8; ++ky) {
                                const uint8_t* xr = &hx[((img * 4 + c) * H + 4 * oy + ky) * W + 4 * ox];
                                double* rr = &ref[((oc * 4 + c) * 8 + ky) * 8];
                                double* ss = &scale[((oc * 4 + c) * 8 + ky) * 8];
                                for (int kx = 0; kx < 8; ++kx) {
                                    const double term = gv * xr[kx] / 255.0;
                                    rr[kx] += term;
                                    ss[kx] += std::fabs(term);
                                }
                            }
                    }
        }
        double max_ratio = 0.0, max_rel = 0.0, max_b = 0.0;
        for (int i = 0; i < 4096; ++i) {
            const double err = std::fabs(gw[i] - ref[i]);
            max_ratio = std::fmax(max_ratio, err / (scale[i] + 1e-300));
            max_rel = std::fmax(max_rel, err / (std::fabs(ref[i]) + 1e-300));
        }
        for (int i = 0; i < 16; ++i) max_b = std::fmax(max_b, std::fabs(gb[i] - refb[i]) / (std::fabs(refb[i]) + 1e-30));
        printf("[wgrad] N=%d %dx%d rows=%d mask=%d: max err/sum|g||x| = %.3e, max rel err = %.3e, bias max rel err = %.3e -> %s\n", N, H, W,
               use_rows ? 1 : 0, use_mask ? 1 : 0, max_ratio, max_rel, max_b, (max_ratio <= 1e-6 && max_b <= 1e-4) ? "OK" : "MISMATCH");
    }
    if (timing) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int i = 0; i < 3; ++i) launch();
        cudaEventRecord(e0);
        for (int i = 0; i < 20; ++i) launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double us = ms * 1000.0 / 20.0;
        const double bytes = static_cast<double>(N) * (4.0 * H * W + (use_mask ? 2.0 : 1.0) * 16.0 * P * 4.0);
        printf("[wgrad] N=%d %dx%d mask=%d: %.1f us per call (absmax + wgrad + reduce), %.0f GB/s of algorithmic bytes (v1: ~780 us)\n", N, H, W,
               use_mask ? 1 : 0, us, bytes / us * 1e-3);
    }
    cudaFree(dx); cudaFree(dg); cudaFree(dout); cudaFree(dw); cudaFree(db); cudaFree(drows); cudaFree(scratch);
    return 0;
}

int main(int argc, char** argv) {
    const int which = argc > 1 ? atoi(argv[1]) : 3;
    if (which & 1) mma_probes();
    if (which & 2) {
        conv_case(3, 84, 84, false, false);
        conv_case(300, 84, 84, true, false);
        conv_case(70, 104, 80, true, false);
        conv_case(256, 84, 84, false, true);
        conv_case(8192, 84, 84, true, true);
        conv_case(8192, 104, 80, false, true);
    }
    if (which & 8) conv_case(8192, 84, 84, true, true, 0);    // the ncu target
    if (which & 16) {
        wgrad_case(5, 84, 84, false, false, true, false);
        wgrad_case(300, 84, 84, true, true, true, false);
        wgrad_case(70, 104, 80, true, true, true, false);
        wgrad_case(8192, 84, 84, true, true, false, true);
        wgrad_case(8192, 84, 84, true, false, false, true);
        wgrad_case(8192, 104, 80, false, true, false, true);
    }
    if (which & 32) wgrad_case(8192, 84, 84, true, true, false, true);   // the ncu target
    return 0;
}
