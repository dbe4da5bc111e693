// Error plumbing + device queries for the C ABI (no kernels here).
#include <stdarg.h>
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

extern "C" {

int rl_b200_abi_version(void) { return 1; }

const char* rl_b200_last_error(void) { return rl::g_err; }

int rl_b200_sm_count(void) {
    int n = rl::sm_count();
    if (n < 0) {
        cudaError_t e = cudaGetLastError();
        rl::set_error("rl_b200_sm_count: %s", cudaGetErrorString(e));
    }
    return n;
}

}  // extern "C"
