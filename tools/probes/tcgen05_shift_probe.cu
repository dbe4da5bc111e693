// Round-2 experiment 0 ("v2" conv kernels): can a tcgen05 K-major SWIZZLE_128B
// operand descriptor start at an arbitrary ROW of a shared-memory tile (start address 128-byte but not
// 1024-byte aligned), and which "base offset" value (descriptor bits [49,52)) does it need?
//
// One CTA fills a 192-row x 128-byte K-major SWIZZLE_128B region (chunk j of row p at
// p*128 + ((j ^ (p & 7)) << 4), exactly how the conv producers write tiles) with A[p][e] = p (pass 0)
// or A[p][e] = e (pass 1), a 16-row B tile with B[n][e] = (e == n), and issues M=128, N=16, K=8 x 4
// tf32 MMAs with the A descriptor advanced by `shift` rows.  Expected D[i][n] = A[shift + i][n]:
// pass 0 -> shift + i for every n, pass 1 -> n.  Prints the number of mismatches for every
// (shift, base_offset) pair so the right encoding can be read off one run:
//     nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I rlpyt_b200/csrc \
//          -o tools/probes/_bin/shift_probe tools/probes/tcgen05_shift_probe.cu && tools/probes/_bin/shift_probe
// Not part of the library; never measured or validated in round 1 (written after the GPU budget was spent).
#include <cstdio>
#include <vector>

#include "tc_common.cuh"

using namespace rl::tc;

constexpr int kRowsA = 192, kN = 16;

__global__ void __launch_bounds__(128, 1)
shift_probe_kernel(int shift, int base_offset, int pass, float* __restrict__ out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* a_tile = smem;                              // 192 x 128 B = 24 KiB
    uint8_t* b_tile = smem + kRowsA * 128;               // 16 x 128 B (1024-aligned: 24576 = 24 * 1024)
    uint64_t* bar = reinterpret_cast<uint64_t*>(b_tile + 2048);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    for (int idx = threadIdx.x; idx < kRowsA * 32; idx += blockDim.x) {
        const int p = idx >> 5, e = idx & 31;
        const float v = pass == 0 ? static_cast<float>(p) : static_cast<float>(e);
        *reinterpret_cast<float*>(a_tile + p * 128 + (((e >> 2) ^ (p & 7)) << 4) + (e & 3) * 4) = v;
    }
    for (int idx = threadIdx.x; idx < kN * 32; idx += blockDim.x) {
        const int n = idx >> 5, e = idx & 31;
        *reinterpret_cast<float*>(b_tile + n * 128 + (((e >> 2) ^ (n & 7)) << 4) + (e & 3) * 4) = (e == n) ? 1.0f : 0.0f;
    }
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
        constexpr uint32_t idesc = make_idesc_tf32(128, kN);
        const uint64_t bo = static_cast<uint64_t>(base_offset & 7) << 49;
        const uint64_t da = make_desc(a_tile + shift * 128) | bo;
        const uint64_t db = make_desc(b_tile);
        for (int k = 0; k < 4; ++k)
            umma_tf32(tmem, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), idesc, k > 0 ? 1u : 0u);
        umma_commit(bar);
    }
    mbar_wait(bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t r[32];
    tmem_ld32(tmem + (static_cast<uint32_t>(warp * 32) << 16), r);     // lanes 32*warp.., columns 0..31 (16 used)
    for (int n = 0; n < kN; ++n) out[(warp * 32 + lane) * kN + n] = __uint_as_float(r[n]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(32));
}

// ---- second question: where do the 64 rows of an M = 64 (cta_group::1) accumulator live in TMEM?
// An M = 128 MMA with an all-zero A tile clears lanes 0..127 of 16 columns, then an M = 64 MMA with
// A[p][e] = p + 1 (rows 0..63) and the same selector B writes D[i][n] = i + 1.  Every lane of column 0 is
// read back: lane -> row + 1 (0 = not written by the M = 64 instruction).
__global__ void __launch_bounds__(128, 1)
m64_probe_kernel(float* __restrict__ out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* a_zero = smem;                              // 128 x 128 B zeros
    uint8_t* a_tile = smem + 16384;                      // 64 x 128 B
    uint8_t* b_tile = smem + 16384 + 8192;               // 16 x 128 B
    uint64_t* bar = reinterpret_cast<uint64_t*>(b_tile + 2048);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int idx = threadIdx.x; idx < 128 * 32; idx += blockDim.x) reinterpret_cast<float*>(a_zero)[idx] = 0.0f;
    for (int idx = threadIdx.x; idx < 64 * 32; idx += blockDim.x) {
        const int p = idx >> 5, e = idx & 31;
        *reinterpret_cast<float*>(a_tile + p * 128 + (((e >> 2) ^ (p & 7)) << 4) + (e & 3) * 4) = static_cast<float>(p + 1);
    }
    for (int idx = threadIdx.x; idx < kN * 32; idx += blockDim.x) {
        const int n = idx >> 5, e = idx & 31;
        *reinterpret_cast<float*>(b_tile + n * 128 + (((e >> 2) ^ (n & 7)) << 4) + (e & 3) * 4) = (e == n) ? 1.0f : 0.0f;
    }
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
        const uint64_t db = make_desc(b_tile);
        umma_tf32(tmem, make_desc(a_zero), db, make_idesc_tf32(128, kN), 0u);          // clear 128 lanes
        const uint64_t da = make_desc(a_tile);
        for (int k = 0; k < 4; ++k)
            umma_tf32(tmem, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), make_idesc_tf32(64, kN),
                      k > 0 ? 1u : 0u);
        umma_commit(bar);
    }
    mbar_wait(bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t r[32];
    tmem_ld32(tmem + (static_cast<uint32_t>(warp * 32) << 16), r);
    out[warp * 32 + lane] = __uint_as_float(r[0]);
    out[128 + warp * 32 + lane] = __uint_as_float(r[5]);       // column 5 must hold the same row ids
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(32));
}

int main() {
    float* d_out = nullptr;
    cudaMalloc(&d_out, 128 * kN * sizeof(float));   // also large enough for the 256 floats of the M=64 probe
    const int smem_bytes = kRowsA * 128 + 2048 + 64 + 1024;
    cudaFuncSetAttribute(shift_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    std::vector<float> h(128 * kN);
    const int shifts[] = {0, 1, 2, 7, 8, 9, 21, 22, 33};
    printf("shift base_offset pass mismatches first_bad(i,n,got,want)\n");
    for (int shift : shifts) {
        for (int bo : {0, shift & 7, (8 - (shift & 7)) & 7}) {
            for (int pass = 0; pass < 2; ++pass) {
                shift_probe_kernel<<<1, 128, smem_bytes>>>(shift, bo, pass, d_out);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) {
                    printf("%d %d %d CUDA error: %s\n", shift, bo, pass, cudaGetErrorString(e));
                    return 1;
                }
                cudaMemcpy(h.data(), d_out, h.size() * sizeof(float), cudaMemcpyDeviceToHost);
                int bad = 0, bi = -1, bn = -1;
                for (int i = 0; i < 128; ++i)
                    for (int n = 0; n < kN; ++n) {
                        const float want = pass == 0 ? static_cast<float>(shift + i) : static_cast<float>(n);
                        if (h[i * kN + n] != want && bad++ == 0) { bi = i; bn = n; }
                    }
                if (bad)
                    printf("%5d %11d %4d %10d (%d,%d,%g,%g)\n", shift, bo, pass, bad, bi, bn, h[bi * kN + bn],
                           pass == 0 ? static_cast<float>(shift + bi) : static_cast<float>(bn));
                else
                    printf("%5d %11d %4d %10d\n", shift, bo, pass, 0);
            }
        }
    }
    // ---- M = 64 accumulator layout
    {
        const int smem2 = 16384 + 8192 + 2048 + 64 + 1024;
        cudaFuncSetAttribute(m64_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2);
        m64_probe_kernel<<<1, 128, smem2>>>(d_out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("m64 probe CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
        std::vector<float> m(256);
        cudaMemcpy(m.data(), d_out, 256 * sizeof(float), cudaMemcpyDeviceToHost);
        printf("M=64 accumulator: TMEM lane -> row+1 (0 = untouched), column 0 [column 5 agrees: ");
        bool same = true;
        for (int i = 0; i < 128; ++i) same = same && (m[i] == m[128 + i]);
        printf("%s]\n", same ? "yes" : "NO");
        for (int i = 0; i < 128; ++i) printf("%g%s", m[i], (i % 32 == 31) ? "\n" : " ");
    }
    cudaFree(d_out);
    return 0;
}
