// Round-2 probe: tcgen05.mma with the A operand in tensor memory ("TS" form), and the persistent 3xTF32 GEMM of
// rlpyt_b200/csrc/gemm_ts.cuh against fp64 references, with its time per launch on the three fc shapes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I rlpyt_b200/csrc \
//        -o tools/probes/_bin/gemm_ts tools/probes/gemm_ts_probe.cu && tools/probes/_bin/gemm_ts
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "gemm_ts.cuh"

using namespace rl::gts;
using namespace rl::tc;

static uint32_t g_seed = 4242u;
static uint32_t rnd() { g_seed = g_seed * 1664525u + 1013904223u; return g_seed >> 8; }
static float frand() { return static_cast<float>(rnd() & 0xffff) / 65536.0f - 0.5f; }

// ---- 1. how does the MMA read an A operand from TMEM?  A[lane][col] = lane*8 + col, B = first 8 rows of identity
__global__ void __launch_bounds__(128, 1) ts_decode_kernel(float* out) {
    __shared__ __align__(1024) uint8_t btile[16 * 128];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 16 * 32; i += 128) reinterpret_cast<float*>(btile)[i] = 0.0f;
    __syncthreads();
    if (threadIdx.x < 8) {
        const int n = threadIdx.x, k = n;                       // B[n][k] = 1 at k == n
        const int chunk = (k / 4) ^ (n & 7);
        reinterpret_cast<float*>(btile)[n * 32 + chunk * 4 + (k & 3)] = 1.0f;
    }
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    const uint32_t lane_base = static_cast<uint32_t>(warp * 32) << 16;
    uint32_t v[8];
    for (int j = 0; j < 8; ++j) v[j] = __float_as_uint(static_cast<float>((warp * 32 + lane) * 8 + j));
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(tmem + lane_base + 32),
                 "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp == 0) {
        if (elect_one()) {
            umma_tf32_ts(tmem, tmem + 32, make_desc(btile), make_idesc_tf32(128, 16), 0u);
            umma_commit(&bar);
        }
        __syncwarp();
    }
    mbar_wait(&bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t d[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7]), "=r"(d[8]),
                   "=r"(d[9]), "=r"(d[10]), "=r"(d[11]), "=r"(d[12]), "=r"(d[13]), "=r"(d[14]), "=r"(d[15])
                 : "r"(tmem + lane_base));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 16; ++j) out[(warp * 32 + lane) * 16 + j] = __uint_as_float(d[j]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem));
}

static int decode_case() {
    float* d;
    cudaMalloc(&d, 128 * 16 * 4);
    cudaMemset(d, 0xff, 128 * 16 * 4);
    ts_decode_kernel<<<1, 128>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("[decode] CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<float> h(128 * 16);
    cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 16; ++n) {
            const float want = n < 8 ? static_cast<float>(m * 8 + n) : 0.0f;
            if (h[m * 16 + n] != want) ++bad;
        }
    printf("[decode] A from TMEM (lane = row, column = k): %d mismatches of 2048 -> %s\n", bad, bad ? "MISMATCH" : "OK");
    if (bad) {
        for (int m : {0, 1, 2, 31, 32, 33, 64, 127}) {
            printf("  row %3d:", m);
            for (int n = 0; n < 16; ++n) printf(" %g", h[m * 16 + n]);
            printf("\n");
        }
    }
    cudaFree(d);
    return bad ? 1 : 0;
}

// ---- 2. the GEMM
static int gemm_case(const char* name, int M, int N, int K, int a_mmajor, int c_trans, int with_bias, int relu, bool timing) {
    std::vector<float> ha(static_cast<size_t>(M) * K), hb(static_cast<size_t>(N) * K), hbias(N);
    for (auto& v : ha) v = frand() * 2.0f;
    for (auto& v : hb) v = frand() / 4.0f;
    for (auto& v : hbias) v = frand();
    // device A in the requested layout
    std::vector<float> ha_dev(ha.size());
    if (a_mmajor) { for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) ha_dev[static_cast<size_t>(k) * M + m] = ha[static_cast<size_t>(m) * K + k]; }
    else ha_dev = ha;
    float *da, *db, *dlo, *dbias, *dc, *dws = nullptr;
    cudaMalloc(&da, ha.size() * 4); cudaMalloc(&db, hb.size() * 4); cudaMalloc(&dlo, hb.size() * 4); cudaMalloc(&dbias, N * 4);
    cudaMalloc(&dc, static_cast<size_t>(M) * N * 4);
    cudaMemcpy(da, ha_dev.data(), ha.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(db, hb.data(), hb.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dbias, hbias.data(), N * 4, cudaMemcpyHostToDevice);
    cudaMemset(dc, 0xff, static_cast<size_t>(M) * N * 4);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t wsb = workspace_bytes(M, N, K, sms);
    if (wsb) cudaMalloc(&dws, wsb);
    const Plan pl = make_plan(M, N, K, sms, true);
    auto run = [&]() {
        split_lo_kernel<<<sms * 4, 256>>>(db, dlo, static_cast<int64_t>(N) * K);
        return launch(da, a_mmajor, db, dlo, with_bias ? dbias : nullptr, dc, c_trans, M, N, K, relu, dws, sms, 0);
    };
    cudaError_t e = run();
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("[%s] CUDA error: %s\n", name, cudaGetErrorString(e)); return 1; }
    std::vector<float> hc(static_cast<size_t>(M) * N);
    cudaMemcpy(hc.data(), dc, hc.size() * 4, cudaMemcpyDeviceToHost);
    const int64_t total = static_cast<int64_t>(M) * N;
    const int checks = total < 20000 ? static_cast<int>(total) : 6000;
    double max_ratio = 0.0;
    for (int t = 0; t < checks; ++t) {
        int m, n;
        if (checks == total) { m = t / N; n = t % N; }
        else if (t < 64) { m = (t & 1) ? M - 1 - (t >> 1) % M : (t >> 1) % M; n = (t & 2) ? N - 1 : 0; }   // corners / edges
        else { m = rnd() % M; n = rnd() % N; }
        double acc = 0.0, scale = 0.0;
        for (int k = 0; k < K; ++k) {
            const double a = ha[static_cast<size_t>(m) * K + k], b = hb[static_cast<size_t>(n) * K + k];
            acc += a * b;
            scale += std::fabs(a * b);
        }
        if (with_bias) { acc += hbias[n]; scale += std::fabs(hbias[n]); }
        const double want = relu && acc < 0 ? 0.0 : acc;
        const double got = c_trans ? hc[static_cast<size_t>(n) * M + m] : hc[static_cast<size_t>(m) * N + n];
        const double ratio = std::fabs(got - want) / (scale + 1e-30);
        if (!(ratio <= max_ratio)) max_ratio = ratio;     // also catches NaN
    }
    printf("[%s] M=%d N=%d K=%d a_mmajor=%d c_trans=%d splits=%d grid=%d: max err/sum|a||b| = %.3e -> %s\n", name, M, N, K, a_mmajor,
           c_trans, pl.splits, pl.grid, max_ratio, max_ratio <= 3e-6 ? "OK" : "MISMATCH");
    if (timing) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        for (int i = 0; i < 3; ++i) run();
        cudaEventRecord(e0);
        for (int i = 0; i < 20; ++i) run();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double us = ms * 1000.0 / 20.0;
        cudaEventRecord(e0);
        for (int i = 0; i < 20; ++i) split_lo_kernel<<<sms * 4, 256>>>(db, dlo, static_cast<int64_t>(N) * K);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        const double us_split = ms * 1000.0 / 20.0;
        printf("[%s] %.1f us per call (of which split_lo %.1f us) = %.0f TFLOP/s useful (first kernel: 152-172 us on these shapes)\n", name, us,
               us_split, 2.0 * M * N * K / us * 1e-6);
    }
    cudaFree(da); cudaFree(db); cudaFree(dlo); cudaFree(dbias); cudaFree(dc);
    if (dws) cudaFree(dws);
    return max_ratio <= 3e-6 ? 0 : 1;
}

int main(int argc, char** argv) {
    const int mask = argc > 1 ? atoi(argv[1]) : 0xff;
    int rc = 0;
    if (mask & 1) rc |= decode_case();
    if (mask & 2) {
        rc |= gemm_case("small", 128, 128, 64, 0, 0, 0, 0, false);
        rc |= gemm_case("small-mmajor", 128, 128, 64, 1, 0, 0, 0, false);
        rc |= gemm_case("small-ctrans", 128, 128, 64, 0, 1, 1, 1, false);
        rc |= gemm_case("tails", 200, 72, 100, 0, 0, 1, 1, false);
        rc |= gemm_case("tails-mmajor-ctrans", 200, 72, 100, 1, 1, 1, 0, false);
        rc |= gemm_case("multi-tile", 1000, 384, 1056, 0, 0, 1, 0, false);
    }
    if (mask & 8) gemm_case("fc-fwd", 8192, 512, 3200, 0, 0, 1, 1, true);
    if (mask & 4) {
        rc |= gemm_case("fc-fwd", 8192, 512, 3200, 0, 0, 1, 1, true);
        rc |= gemm_case("fc-dgrad", 8192, 3200, 512, 0, 0, 0, 0, true);
        rc |= gemm_case("fc-wgrad", 3200, 512, 8192, 1, 1, 0, 0, true);
        rc |= gemm_case("step-fwd", 256, 512, 3200, 0, 0, 1, 1, true);
    }
    printf(rc ? "PROBE FAILED\n" : "PROBE OK\n");
    return rc;
}
