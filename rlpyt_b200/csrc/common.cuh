// Shared helpers for the sm_100a kernels of librlpyt_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/rlpyt_b200.h"

namespace rl {

// Thread-local last-error text behind rl_b200_last_error().
void set_error(const char* fmt, ...);
int check_launch(const char* what);  // cudaGetLastError -> status code (+ error text)

#define RL_REQUIRE(cond, code, ...)          \
    do {                                     \
        if (!(cond)) {                       \
            ::rl::set_error(__VA_ARGS__);    \
            return (code);                   \
        }                                    \
    } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static inline bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

// B200: 148 SMs (2 dies x 74).  Queried once; grids for persistent-style kernels are
// sized in multiples of this.
int sm_count();

// ---- device helpers -------------------------------------------------------------
// Streaming (read-once / write-once) accesses: keep them out of L1 so the [T,B] sweep does
// not thrash lines that neighbouring warps still need.
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float ldg_stream(const float* p) {
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t ldg_stream(const uint32_t* p) {
    uint32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ uint8_t ldg_stream(const uint8_t* p) {
    uint32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.u8 %0, [%1];" : "=r"(r) : "l"(p));
    return static_cast<uint8_t>(r);
}
__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream(float4* p, const float4& v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void stg_stream(float* p, float v) {
    asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" :: "l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ void stg_stream(uint4* p, const uint4& v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ bool aligned_dev(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// Warp reductions (fixed shuffle tree => deterministic).
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace rl
