// Round-2 probe: the "v2" second-layer kernels of rlpyt_b200/csrc/conv2_s2d.cuh (space-to-depth cells, row-shifted
// tcgen05 descriptors, bulk-copied images) against fp64 references on the host, and their time per launch.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I rlpyt_b200/csrc \
//        -o tools/probes/_bin/conv2_s2d tools/probes/conv2_s2d_probe.cu && tools/probes/_bin/conv2_s2d [mask]
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "conv2_s2d.cuh"

using namespace rl::c2s;
using namespace rl::tc;

static uint32_t g_seed = 777u;
static uint32_t rnd() { g_seed = g_seed * 1664525u + 1013904223u; return g_seed >> 8; }
static float frand() { return static_cast<float>(rnd() & 0xffff) / 65536.0f - 0.5f; }

static int fwd_case(int N, int IH, int IW, bool timing) {
    if (!geom_ok(16, IH, IW)) { printf("[fwd] %dx%d geometry not supported\n", IH, IW); return 1; }
    Geom g = make_geom(N, IH, IW);
    const int P = g.OH * g.OW;
    std::vector<float> hx(static_cast<size_t>(N) * 16 * IH * IW), hw(32 * 16 * 16), hb(32);
    for (auto& v : hx) { v = frand() * 4.0f; if (v < 0) v = 0; }                  // post-ReLU activations
    for (auto& v : hw) v = frand() / 8.0f;
    for (auto& v : hb) v = frand() / 8.0f;
    float *dx, *dw, *db, *dy;
    const size_t ybytes = static_cast<size_t>(N) * 32 * P * sizeof(float);
    cudaMalloc(&dx, hx.size() * 4); cudaMalloc(&dw, hw.size() * 4); cudaMalloc(&db, 32 * 4); cudaMalloc(&dy, ybytes);
    cudaMemcpy(dx, hx.data(), hx.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dw, hw.data(), hw.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(db, hb.data(), 32 * 4, cudaMemcpyHostToDevice);
    cudaMemset(dy, 0xff, ybytes);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    auto launch = [&]() { return launch_fwd(dx, dw, db, dy, g, 1, sms, 0); };
    cudaError_t e = launch();
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("[fwd] N=%d %dx%d CUDA error: %s\n", N, IH, IW, cudaGetErrorString(e)); return 1; }
    std::vector<float> hy(static_cast<size_t>(N) * 32 * P);
    cudaMemcpy(hy.data(), dy, ybytes, cudaMemcpyDeviceToHost);
    double max_ratio = 0.0;
    const int total = N * 32 * P, checks = total < 60000 ? total : 60000;
    for (int t = 0; t < checks; ++t) {
        int n, oc, oy, ox;
        if (checks == total) { n = t / (32 * P); oc = (t / P) % 32; oy = (t % P) / g.OW; ox = t % g.OW; }
        else { n = rnd() % N; oc = rnd() % 32; oy = rnd() % g.OH; ox = rnd() % g.OW; }
        double acc = 0.0, scale = 0.0;
        for (int c = 0; c < 16; ++c)
            for (int ky = 0; ky < 4; ++ky)
                for (int kx = 0; kx < 4; ++kx) {
                    const int iy = 2 * oy - 1 + ky, ix = 2 * ox - 1 + kx;
                    if (iy < 0 || iy >= IH || ix < 0 || ix >= IW) continue;
                    const double a = hx[((static_cast<size_t>(n) * 16 + c) * IH + iy) * IW + ix];
                    const double w = hw[((oc * 16 + c) * 4 + ky) * 4 + kx];
                    acc += a * w;
                    scale += std::fabs(a * w);
                }
        acc += hb[oc];
        scale += std::fabs(hb[oc]);
        const double want = acc > 0 ? acc : 0;
        const double got = hy[((static_cast<size_t>(n) * 32 + oc) * g.OH + oy) * g.OW + ox];
        max_ratio = std::fmax(max_ratio, std::fabs(got - want) / (scale + 1e-30));
    }
    printf("[fwd] N=%d %dx%d checked=%d max err/sum|x||w| = %.3e -> %s\n", N, IH, IW, checks, max_ratio, max_ratio <= 3e-6 ? "OK" : "MISMATCH");
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
        const double us = ms * 1000.0 / 20.0, bytes = static_cast<double>(N) * (16.0 * IH * IW * 4 + 32.0 * P * 4);
        printf("[fwd] N=%d %dx%d: %.1f us per launch, %.0f GB/s of algorithmic bytes (v1 tcgen05 kernel: ~290 us at N=8192 20x20)\n", N, IH, IW, us,
               bytes / us * 1e-3);
    }
    cudaFree(dx); cudaFree(dw); cudaFree(db); cudaFree(dy);
    return 0;
}

static int dgrad_case(int N, int IH, int IW, bool timing) {
    if (!geom_ok(16, IH, IW)) { printf("[dgrad] %dx%d geometry not supported\n", IH, IW); return 1; }
    Geom g = make_geom(N, IH, IW);
    if (!dg::smem_ok(g)) { printf("[dgrad] %dx%d shared memory does not fit\n", IH, IW); return 1; }
    const int P = g.OH * g.OW;
    std::vector<float> hg(static_cast<size_t>(N) * 32 * P), hw(32 * 16 * 16);
    for (auto& v : hg) { v = frand(); if ((rnd() & 3) == 0) v = 0.0f; }          // masked gradient
    for (auto& v : hw) v = frand() / 8.0f;
    float *dgp, *dw, *dx;
    const size_t xbytes = static_cast<size_t>(N) * 16 * IH * IW * sizeof(float);
    cudaMalloc(&dgp, hg.size() * 4); cudaMalloc(&dw, hw.size() * 4); cudaMalloc(&dx, xbytes);
    cudaMemcpy(dgp, hg.data(), hg.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dw, hw.data(), hw.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dx, 0xff, xbytes);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    auto launch = [&]() { return dg::launch_dgrad(dgp, dw, dx, g, sms, 0); };
    cudaError_t e = launch();
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("[dgrad] N=%d %dx%d CUDA error: %s\n", N, IH, IW, cudaGetErrorString(e)); return 1; }
    std::vector<float> hx(static_cast<size_t>(N) * 16 * IH * IW);
    cudaMemcpy(hx.data(), dx, xbytes, cudaMemcpyDeviceToHost);
    double max_ratio = 0.0;
    const int total = N * 16 * IH * IW, checks = total < 60000 ? total : 60000;
    int nans = 0;
    for (int t = 0; t < checks; ++t) {
        int n, c, y, x;
        if (checks == total) { n = t / (16 * IH * IW); c = (t / (IH * IW)) % 16; y = (t % (IH * IW)) / IW; x = t % IW; }
        else { n = rnd() % N; c = rnd() % 16; y = rnd() % IH; x = rnd() % IW; }
        double acc = 0.0, scale = 0.0;
        for (int oc = 0; oc < 32; ++oc)
            for (int ky = 0; ky < 4; ++ky)
                for (int kx = 0; kx < 4; ++kx) {
                    const int ty = y + 1 - ky, tx = x + 1 - kx;
                    if (ty < 0 || tx < 0 || (ty & 1) || (tx & 1)) continue;
                    const int oy = ty / 2, ox = tx / 2;
                    if (oy >= g.OH || ox >= g.OW) continue;
                    const double a = hg[((static_cast<size_t>(n) * 32 + oc) * g.OH + oy) * g.OW + ox];
                    const double w = hw[((oc * 16 + c) * 4 + ky) * 4 + kx];
                    acc += a * w;
                    scale += std::fabs(a * w);
                }
        const double got = hx[((static_cast<size_t>(n) * 16 + c) * IH + y) * IW + x];
        if (got != got) { ++nans; continue; }
        max_ratio = std::fmax(max_ratio, std::fabs(got - acc) / (scale + 1e-30));
    }
    printf("[dgrad] N=%d %dx%d checked=%d nans=%d max err/sum|g||w| = %.3e -> %s\n", N, IH, IW, checks, nans, max_ratio,
           (max_ratio <= 3e-6 && nans == 0) ? "OK" : "MISMATCH");
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
        const double us = ms * 1000.0 / 20.0, bytes = static_cast<double>(N) * (16.0 * IH * IW * 4 + 32.0 * P * 4);
        printf("[dgrad] N=%d %dx%d: %.1f us per launch, %.0f GB/s of algorithmic bytes (v1 tcgen05 kernel: ~600 us at N=8192 20x20)\n", N, IH, IW,
               us, bytes / us * 1e-3);
    }
    cudaFree(dgp); cudaFree(dw); cudaFree(dx);
    return 0;
}

static int wgrad_case(int N, int IH, int IW, bool check, bool timing) {
    if (!geom_ok(16, IH, IW)) { printf("[wgrad] %dx%d geometry not supported\n", IH, IW); return 1; }
    Geom g = make_geom(N, IH, IW);
    if (!wg2::smem_ok(g)) { printf("[wgrad] %dx%d shared memory does not fit\n", IH, IW); return 1; }
    const int P = g.OH * g.OW;
    std::vector<float> hx(static_cast<size_t>(N) * 16 * IH * IW), hg(static_cast<size_t>(N) * 32 * P);
    for (auto& v : hx) { v = frand() * 4.0f; if (v < 0) v = 0; }
    for (auto& v : hg) { v = frand(); if ((rnd() & 3) == 0) v = 0.0f; }
    float *dx, *dgp, *dw, *db; void* scratch;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaMalloc(&dx, hx.size() * 4); cudaMalloc(&dgp, hg.size() * 4); cudaMalloc(&dw, 32 * 256 * 4); cudaMalloc(&db, 32 * 4);
    cudaMalloc(&scratch, wg2::scratch_bytes(sms));
    cudaMemcpy(dx, hx.data(), hx.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dgp, hg.data(), hg.size() * 4, cudaMemcpyHostToDevice);
    auto launch = [&]() { return wg2::launch_wgrad(dx, dgp, dw, db, g, sms, scratch, 0); };
    cudaError_t e = launch();
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("[wgrad] N=%d %dx%d CUDA error: %s\n", N, IH, IW, cudaGetErrorString(e)); return 1; }
    if (check) {
        std::vector<float> gw(32 * 256), gb(32);
        cudaMemcpy(gw.data(), dw, gw.size() * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(gb.data(), db, 32 * 4, cudaMemcpyDeviceToHost);
        std::vector<double> ref(32 * 256, 0.0), sc(32 * 256, 0.0), refb(32, 0.0), scb(32, 0.0);
        for (int n = 0; n < N; ++n)
            for (int oc = 0; oc < 32; ++oc)
                for (int oy = 0; oy < g.OH; ++oy)
                    for (int ox = 0; ox < g.OW; ++ox) {
                        const double gv = hg[((static_cast<size_t>(n) * 32 + oc) * g.OH + oy) * g.OW + ox];
                        if (gv == 0.0) continue;
                        refb[oc] += gv; scb[oc] += std::fabs(gv);
                        for (int c = 0; c < 16; ++c)
                            for (int ky = 0; ky < 4; ++ky)
                                for (int kx = 0; kx < 4; ++kx) {
                                    const int iy = 2 * oy - 1 + ky, ix = 2 * ox - 1 + kx;
                                    if (iy < 0 || iy >= IH || ix < 0 || ix >= IW) continue;
                                    const double t = gv * hx[((static_cast<size_t>(n) * 16 + c) * IH + iy) * IW + ix];
                                    ref[((oc * 16 + c) * 4 + ky) * 4 + kx] += t;
                                    sc[((oc * 16 + c) * 4 + ky) * 4 + kx] += std::fabs(t);
                                }
                    }
        double mr = 0.0, mb = 0.0;
        for (int i = 0; i < 32 * 256; ++i) mr = std::fmax(mr, std::fabs(gw[i] - ref[i]) / (sc[i] + 1e-300));
        for (int i = 0; i < 32; ++i) mb = std::fmax(mb, std::fabs(gb[i] - refb[i]) / (scb[i] + 1e-300));
        printf("[wgrad] N=%d %dx%d: max err/sum|g||x| = %.3e, bias %.3e -> %s\n", N, IH, IW, mr, mb, (mr <= 1e-5 && mb <= 1e-5) ? "OK" : "MISMATCH");
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
        const double us = ms * 1000.0 / 20.0, bytes = static_cast<double>(N) * (16.0 * IH * IW * 4 + 32.0 * P * 4);
        printf("[wgrad] N=%d %dx%d: %.1f us per call (kernel + reduce), %.0f GB/s of algorithmic bytes (v1 tcgen05 kernel: ~300 us at N=8192 20x20)\n", N, IH,
               IW, us, bytes / us * 1e-3);
    }
    cudaFree(dx); cudaFree(dgp); cudaFree(dw); cudaFree(db); cudaFree(scratch);
    return 0;
}

// ---------------------------------------------------------------------------------------------- MN-major tf32 semantics
// One CTA copies a host-built shared-memory image, issues n_mma tf32 MMAs with descriptors (a_desc + i*a_step,
// b_desc + i*b_step) relative to the 1024-aligned base, dumps the first 32 columns of all 128 TMEM lanes.
__global__ void __launch_bounds__(128, 1)
tf32_mma_probe_kernel(const float* __restrict__ img, int img_floats, uint64_t a_desc, uint64_t b_desc, int a_step, int b_step, int n_mma,
                      uint32_t idesc, float* __restrict__ out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    float* sm = reinterpret_cast<float*>(smem);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + img_floats * 4);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < img_floats; i += blockDim.x) sm[i] = img[i];
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(32));
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
            umma_tf32(tmem, a_desc + base16 + static_cast<uint64_t>(i * a_step), b_desc + base16 + static_cast<uint64_t>(i * b_step), idesc, i > 0 ? 1u : 0u);
        umma_commit(bar);
    }
    mbar_wait(bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t r[32];
    tmem_ld32(tmem + (static_cast<uint32_t>(warp * 32) << 16), r);
    for (int n = 0; n < 32; ++n) out[(warp * 32 + lane) * 32 + n] = __uint_as_float(r[n]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(32));
}

static void mn_major_probes() {
    // A: two planes of [64 k-rows][32 floats] (plane p holds m in [32p, 32p+32)), MN-major; B: [64 k-rows][32 floats] MN-major.
    // Small integers: exact in TF32.  Expected D[m][n] = sum_k A[k + shift][m] * B[k][n] over K = 8 * n_mma.
    const int rows = 64, plane_f = rows * 32, a_off = 0, b_off = 2 * plane_f;
    std::vector<float> img(3 * plane_f, 0.0f), Al(rows * 64), Bl(rows * 32);
    for (auto& v : Al) v = static_cast<float>(static_cast<int>(rnd() % 15) - 7);
    for (auto& v : Bl) v = static_cast<float>(static_cast<int>(rnd() % 15) - 7);
    // layout_type 2: SWIZZLE_128B (16-byte units ^ row & 7), 1: SWIZZLE_128B_BASE32B (32-byte units ^ row & 3)
    auto fill = [&](int layout_type) {
        for (int k = 0; k < rows; ++k) {
            auto col = [&](int e) { return layout_type == 2 ? (((e >> 2) ^ (k & 7)) << 2) + (e & 3) : (((e >> 3) ^ (k & 3)) << 3) + (e & 7); };
            for (int m = 0; m < 64; ++m) img[a_off + (m >> 5) * plane_f + k * 32 + col(m & 31)] = Al[k * 64 + m];
            for (int n = 0; n < 32; ++n) img[b_off + k * 32 + col(n)] = Bl[k * 32 + n];
        }
    };
    float *d_img, *d_out;
    cudaMalloc(&d_img, img.size() * 4); cudaMalloc(&d_out, 128 * 32 * 4);
    const int smem_bytes = static_cast<int>(img.size() * 4) + 64 + 1024;
    cudaFuncSetAttribute(tf32_mma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    int layout_type = 2;
    auto desc = [&](uint64_t start_bytes, uint64_t lbo_bytes, uint64_t sbo_bytes) {
        return (start_bytes >> 4) | ((lbo_bytes >> 4) << 16) | ((sbo_bytes >> 4) << 32) | (1ull << 46) | (static_cast<uint64_t>(layout_type) << 61);
    };
    const uint32_t idesc = make_idesc_tf32(64, 32) | (1u << 15) | (1u << 16);
    struct V { const char* name; int layout; uint64_t a_lbo, a_sbo, b_lbo, b_sbo; };
    const V variants[] = {{"SW128 (16B units)    A: LBO=plane SBO=1024 | B: SBO=1024", 2, static_cast<uint64_t>(plane_f) * 4, 1024, 1024, 1024},
                          {"SW128_BASE32B        A: LBO=plane SBO=512  | B: SBO=512 ", 1, static_cast<uint64_t>(plane_f) * 4, 512, 512, 512},
                          {"SW128_BASE32B        A: LBO=512 SBO=plane  | B: SBO=512 ", 1, 512, static_cast<uint64_t>(plane_f) * 4, 512, 512},
                          {"SW128_BASE32B        A: LBO=plane SBO=1024 | B: SBO=1024", 1, static_cast<uint64_t>(plane_f) * 4, 1024, 1024, 1024}};
    for (const V& v : variants) {
        layout_type = v.layout;
        fill(v.layout);
        cudaMemcpy(d_img, img.data(), img.size() * 4, cudaMemcpyHostToDevice);
        for (int n_mma : {1, 4})
            for (int shift : {0, 11}) {
                tf32_mma_probe_kernel<<<1, 128, smem_bytes>>>(d_img, static_cast<int>(img.size()), desc(a_off * 4 + shift * 128, v.a_lbo, v.a_sbo),
                                                              desc(b_off * 4, v.b_lbo, v.b_sbo), 64, 64, n_mma, idesc, d_out);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("[mn] %s CUDA error: %s\n", v.name, cudaGetErrorString(e)); return; }
                std::vector<float> h(128 * 32);
                cudaMemcpy(h.data(), d_out, h.size() * 4, cudaMemcpyDeviceToHost);
                int bad = 0, first = -1;
                for (int m = 0; m < 64; ++m) {
                    const int lane = (m >> 4) * 32 + (m & 15);                    // M = 64 accumulator: lanes 0-15, 32-47, 64-79, 96-111
                    for (int n = 0; n < 32; ++n) {
                        float want = 0.0f;
                        for (int k = 0; k < 8 * n_mma; ++k) want += Al[(k + shift) * 64 + m] * Bl[k * 32 + n];
                        if (h[lane * 32 + n] != want && bad++ == 0) first = m * 32 + n;
                    }
                }
                if (bad) printf("[mn] %s n_mma=%d shift=%2d: %4d mismatches, first (m=%d,n=%d) got %g\n", v.name, n_mma, shift, bad, first / 32, first % 32,
                                h[((first / 32 >> 4) * 32 + (first / 32 & 15)) * 32 + first % 32]);
                else printf("[mn] %s n_mma=%d shift=%2d: OK\n", v.name, n_mma, shift);
            }
    }
    cudaFree(d_img); cudaFree(d_out);
}

int main(int argc, char** argv) {
    const int which = argc > 1 ? atoi(argv[1]) : 1;
    if (which & 1) {
        fwd_case(3, 20, 20, false);
        fwd_case(301, 20, 20, false);
        fwd_case(70, 25, 19, false);
        fwd_case(5, 7, 5, false);
        fwd_case(9, 8, 6, false);
        fwd_case(256, 20, 20, true);
        fwd_case(8192, 20, 20, true);
        fwd_case(8192, 25, 19, true);
    }
    if (which & 2) {
        dgrad_case(3, 20, 20, false);
        dgrad_case(301, 20, 20, false);
        dgrad_case(70, 25, 19, false);
        dgrad_case(9, 8, 6, false);
        dgrad_case(256, 20, 20, true);
        dgrad_case(8192, 20, 20, true);
        dgrad_case(8192, 25, 19, true);
    }
    if (which & 8) mn_major_probes();
    if (which & 4) {
        wgrad_case(3, 20, 20, true, false);
        wgrad_case(301, 20, 20, true, false);
        wgrad_case(70, 25, 19, true, false);
        wgrad_case(9, 8, 6, true, false);
        wgrad_case(8192, 20, 20, false, true);
        wgrad_case(8192, 25, 19, false, true);
    }
    return 0;
}
