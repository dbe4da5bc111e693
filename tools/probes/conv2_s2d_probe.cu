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
