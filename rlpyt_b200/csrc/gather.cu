// Indexed row gathers: the PPO minibatch former (rlpyt/algos/pg/ppo.py:94-100 fancy-indexes every
// field of LossInputs with [T_idxs, B_idxs]; observations are 231 MB per minibatch at the config)
// and the generic "x[idx]" used by the replay extraction of small fields.
//
// Pure HBM-bound byte movement: read row_bytes + write row_bytes per gathered row (+8 B index).
// Rows are contiguous, so the only question is keeping enough 16 B requests in flight:
//   * gather_rows_vec16_kernel: rows of >= 256 B whose size and base are 16 B multiples
//     (observations: 4*84*84 = 28224 B = 1764 uint4).  A CTA copies one 4 KiB slab of one row:
//     256 threads x 1 uint4 x kUnroll(=4) independent loads issued before the stores.
//   * gather_rows_small_kernel: 4-byte-word granularity for the scalar fields (action, return,
//     advantage, valid, old prob[A]), any number of fields (<= 8) in ONE launch.
#include "common.cuh"

namespace rl {

constexpr int kGatherThreads = 256;
constexpr int kGatherUnroll = 4;

__global__ void __launch_bounds__(kGatherThreads)
gather_rows_vec16_kernel(const uint4* __restrict__ src, const int64_t* __restrict__ idx,
                         uint4* __restrict__ dst, int64_t row_vec, int slabs_per_row) {
    const int64_t row = blockIdx.x / slabs_per_row;
    const int slab = blockIdx.x % slabs_per_row;
    const int64_t s = idx[row];
    const uint4* in = src + s * row_vec;
    uint4* out = dst + row * row_vec;
    const int64_t base = static_cast<int64_t>(slab) * kGatherThreads * kGatherUnroll + threadIdx.x;
    uint4 v[kGatherUnroll];
#pragma unroll
    for (int u = 0; u < kGatherUnroll; ++u) {
        const int64_t j = base + static_cast<int64_t>(u) * kGatherThreads;
        if (j < row_vec) v[u] = ldg_stream(in + j);
    }
#pragma unroll
    for (int u = 0; u < kGatherUnroll; ++u) {
        const int64_t j = base + static_cast<int64_t>(u) * kGatherThreads;
        if (j < row_vec) stg_stream(out + j, v[u]);
    }
}

// Byte-granular fallback (row sizes that are not 16 B multiples, e.g. 1x1 test frames).
__global__ void gather_rows_bytes_kernel(const uint8_t* __restrict__ src, const int64_t* __restrict__ idx,
                                         uint8_t* __restrict__ dst, int64_t row_bytes, int64_t n) {
    const int64_t total = n * row_bytes;
    for (int64_t j = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; j < total;
         j += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t row = j / row_bytes, off = j % row_bytes;
        dst[j] = src[idx[row] * row_bytes + off];
    }
}

struct GatherFields {
    const void* src[RL_GATHER_MAX_FIELDS];
    void* dst[RL_GATHER_MAX_FIELDS];
    int32_t row_words[RL_GATHER_MAX_FIELDS];  // row size in 4-byte words
    int32_t n_fields;
};

__global__ void gather_rows_small_kernel(GatherFields f, const int64_t* __restrict__ idx, int64_t n) {
    const int field = blockIdx.y;
    const int rw = f.row_words[field];
    const uint32_t* __restrict__ src = static_cast<const uint32_t*>(f.src[field]);
    uint32_t* __restrict__ dst = static_cast<uint32_t*>(f.dst[field]);
    const int64_t total = n * rw;
    for (int64_t j = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; j < total;
         j += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t row = j / rw;
        const int w = static_cast<int>(j - row * rw);
        dst[j] = src[idx[row] * rw + w];
    }
}

}  // namespace rl

extern "C" {

int rl_gather_rows(const void* src, const int64_t* idx, void* dst, int64_t n, int64_t row_bytes,
                   void* stream) {
    RL_REQUIRE(src && idx && dst, RL_EINVAL, "rl_gather_rows: null pointer");
    RL_REQUIRE(n >= 0 && row_bytes >= 1, RL_EINVAL, "rl_gather_rows: n=%lld row_bytes=%lld",
               (long long)n, (long long)row_bytes);
    if (n == 0) return RL_OK;
    cudaStream_t st = rl::as_stream(stream);
    if (row_bytes % 16 == 0 && row_bytes >= 256 && rl::aligned(src, 16) && rl::aligned(dst, 16)) {
        const int64_t row_vec = row_bytes / 16;
        const int per = rl::kGatherThreads * rl::kGatherUnroll;
        const int slabs = static_cast<int>((row_vec + per - 1) / per);
        const int64_t blocks = n * slabs;
        RL_REQUIRE(blocks < (1LL << 31), RL_EINVAL, "rl_gather_rows: too many blocks");
        rl::gather_rows_vec16_kernel<<<static_cast<unsigned>(blocks), rl::kGatherThreads, 0, st>>>(
            static_cast<const uint4*>(src), idx, static_cast<uint4*>(dst), row_vec, slabs);
        return rl::check_launch("gather_rows_vec16_kernel");
    }
    if (row_bytes % 4 == 0 && rl::aligned(src, 4) && rl::aligned(dst, 4)) {
        rl::GatherFields f{};
        f.src[0] = src; f.dst[0] = dst; f.row_words[0] = static_cast<int32_t>(row_bytes / 4); f.n_fields = 1;
        const int64_t total = n * (row_bytes / 4);
        int64_t blocks = (total + 255) / 256;
        if (blocks > 148 * 16) blocks = 148 * 16;
        rl::gather_rows_small_kernel<<<dim3(static_cast<unsigned>(blocks), 1), 256, 0, st>>>(f, idx, n);
        return rl::check_launch("gather_rows_small_kernel");
    }
    int64_t blocks = (n * row_bytes + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    rl::gather_rows_bytes_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(
        static_cast<const uint8_t*>(src), idx, static_cast<uint8_t*>(dst), row_bytes, n);
    return rl::check_launch("gather_rows_bytes_kernel");
}

int rl_gather_rows_multi(int n_fields, const void* const* src, void* const* dst, const int64_t* row_bytes,
                         const int64_t* idx, int64_t n, void* stream) {
    RL_REQUIRE(src && dst && row_bytes && idx, RL_EINVAL, "rl_gather_rows_multi: null pointer");
    RL_REQUIRE(n_fields >= 1 && n_fields <= RL_GATHER_MAX_FIELDS, RL_EINVAL,
               "rl_gather_rows_multi: n_fields=%d (max %d)", n_fields, RL_GATHER_MAX_FIELDS);
    if (n <= 0) return RL_OK;
    rl::GatherFields f{};
    int64_t max_words = 0;
    for (int k = 0; k < n_fields; ++k) {
        RL_REQUIRE(src[k] && dst[k], RL_EINVAL, "rl_gather_rows_multi: field %d null", k);
        RL_REQUIRE(row_bytes[k] >= 4 && row_bytes[k] % 4 == 0 && rl::aligned(src[k], 4) && rl::aligned(dst[k], 4),
                   RL_EALIGN, "rl_gather_rows_multi: field %d rows must be 4B multiples (got %lld)", k,
                   (long long)row_bytes[k]);
        f.src[k] = src[k]; f.dst[k] = dst[k];
        f.row_words[k] = static_cast<int32_t>(row_bytes[k] / 4);
        if (f.row_words[k] > max_words) max_words = f.row_words[k];
    }
    f.n_fields = n_fields;
    int64_t blocks = (n * max_words + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    rl::gather_rows_small_kernel<<<dim3(static_cast<unsigned>(blocks), n_fields), 256, 0,
                                   rl::as_stream(stream)>>>(f, idx, n);
    return rl::check_launch("gather_rows_small_kernel");
}

}  // extern "C"
