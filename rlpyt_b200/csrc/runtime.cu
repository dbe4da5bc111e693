// Error plumbing + device queries for the C ABI (no kernels here).
#if defined(__x86_64__) || defined(_M_X64)
#include <emmintrin.h>
#define RL_HAVE_SSE2 1
#endif
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace rl {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return static_cast<int>(e);
    }
    return RL_OK;
}

int sm_count() {
    static int cached = 0;
    if (cached > 0) return cached;
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    cached = n;
    return n;
}

}  // namespace rl

#ifdef RL_HAVE_SSE2
#include <immintrin.h>
namespace rl {
// Whole 64-byte lines with non-temporal stores; the widest vector unit the host has (one full line per store on
// AVX-512 keeps a write-combining buffer busy for a single instruction).  Returns the number of bytes copied.
__attribute__((target("avx512f"))) static int64_t stream_lines_avx512(uint8_t* d, const uint8_t* s, int64_t n) {
    int64_t i = 0;
    for (; i + 64 <= n; i += 64) _mm512_stream_si512(reinterpret_cast<__m512i*>(d + i), _mm512_loadu_si512(s + i));
    return i;
}
__attribute__((target("avx2"))) static int64_t stream_lines_avx2(uint8_t* d, const uint8_t* s, int64_t n) {
    int64_t i = 0;
    for (; i + 64 <= n; i += 64) {
        const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + i));
        const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + i + 32));
        _mm256_stream_si256(reinterpret_cast<__m256i*>(d + i), a);
        _mm256_stream_si256(reinterpret_cast<__m256i*>(d + i + 32), b);
    }
    return i;
}
static int64_t stream_lines_sse2(uint8_t* d, const uint8_t* s, int64_t n) {
    int64_t i = 0;
    for (; i + 64 <= n; i += 64) {
        const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i));
        const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i + 16));
        const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i + 32));
        const __m128i e = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i + 48));
        _mm_stream_si128(reinterpret_cast<__m128i*>(d + i), a);
        _mm_stream_si128(reinterpret_cast<__m128i*>(d + i + 16), b);
        _mm_stream_si128(reinterpret_cast<__m128i*>(d + i + 32), c);
        _mm_stream_si128(reinterpret_cast<__m128i*>(d + i + 48), e);
    }
    return i;
}
static int64_t stream_copy_lines(uint8_t* d, const uint8_t* s, int64_t n) {
    // Default: 128-bit stores.  Measured on the B200 host with the alternating sampler (14 env workers on the two
    // hardware threads of 7 cores): with 512-bit non-temporal stores an env step took 2.4x as long (wait_envs
    // 26 -> 303 us per batch step, profiles/r02_sampler_configs.txt) - wide-vector frequency licence + sibling
    // interference cost more than the copy gained in isolation (8.0 -> 6.7 us per 28 KB frame).  RLPYT_B200_STREAM_COPY =
    // avx2 | avx512 selects the wider paths for hosts where they pay.
    static const int level = [] {
        const char* e = getenv("RLPYT_B200_STREAM_COPY");
        if (e != nullptr && strcmp(e, "avx512") == 0 && __builtin_cpu_supports("avx512f")) return 2;
        if (e != nullptr && strcmp(e, "avx2") == 0 && __builtin_cpu_supports("avx2")) return 1;
        return 0;
    }();
    const int64_t i = level == 2 ? stream_lines_avx512(d, s, n) : (level == 1 ? stream_lines_avx2(d, s, n) : stream_lines_sse2(d, s, n));
    _mm_sfence();
    return i;
}
}  // namespace rl
#endif

extern "C" {

int rl_b200_abi_version(void) { return 2; }   // 2: + rl_conv1_u8_*_i8

const char* rl_b200_last_error(void) { return rl::g_err; }

int rl_b200_sm_count(void) {
    int n = rl::sm_count();
    if (n < 0) {
        cudaError_t e = cudaGetLastError();
        rl::set_error("rl_b200_sm_count: %s", cudaGetErrorString(e));
    }
    return n;
}

/* Host-side helper for the env workers (no CUDA call inside): copy an observation into the
 * page-locked step buffer with non-temporal stores, so the lines go to memory instead of staying
 * Modified in the writing core's private L2.  The GPU copy engine then reads them at PCIe rate
 * (measured on the B200 host: 7.2 MB H2D takes 1.2 ms when 32 worker cores hold the lines dirty,
 * 0.17 ms otherwise).  Falls back to memcpy for unaligned tails. */
int rl_host_stream_copy(void* dst, const void* src, int64_t nbytes) {
    if (dst == nullptr || src == nullptr || nbytes < 0) return RL_EINVAL;
    uint8_t* d = static_cast<uint8_t*>(dst);
    const uint8_t* s = static_cast<const uint8_t*>(src);
    int64_t i = 0;
#ifdef RL_HAVE_SSE2      // x86 hosts only; on other hosts (Grace/ARM) the whole copy is the memcpy below
    if ((reinterpret_cast<uintptr_t>(d) & 63) == 0) i = rl::stream_copy_lines(d, s, nbytes);
#endif
    if (i < nbytes) memcpy(d + i, s + i, static_cast<size_t>(nbytes - i));
    return RL_OK;
}

/* Asynchronous host -> device copy of one worker's observation rows out of the page-locked step buffer
 * (rlpyt/samplers/parallel/gpu/action_server.py:46-48 uploads the whole step buffer after ALL workers are done; here
 * the master uploads each worker's rows as soon as that worker has signalled, so only the last worker's 0.5 MB is on
 * the critical path).  A thin cudaMemcpyAsync: ~2 us of host time per call through ctypes against ~10 us for
 * Tensor.copy_. */
int rl_upload_async(void* dst_device, const void* src_host, int64_t nbytes, void* stream) {
    RL_REQUIRE(dst_device && src_host && nbytes >= 0, RL_EINVAL, "rl_upload_async: bad argument");
    if (nbytes == 0) return RL_OK;
    cudaError_t e = cudaMemcpyAsync(dst_device, src_host, static_cast<size_t>(nbytes), cudaMemcpyHostToDevice, rl::as_stream(stream));
    if (e != cudaSuccess) {
        rl::set_error("rl_upload_async: %s", cudaGetErrorString(e));
        return static_cast<int>(e);
    }
    return RL_OK;
}

}  // extern "C"
