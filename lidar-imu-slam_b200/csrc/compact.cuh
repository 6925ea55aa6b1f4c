// compact.cuh -- stable (index-order) stream compaction of a flag array.
//
// The reference's "first point per voxel" / "inlier" / "within range" selections all keep survivors in
// input order (serial loops, icp.cpp:13-19, :112-121; voxel_hash_map.cpp:114-125), so every selection on
// the device is a flag pass followed by this order-preserving compaction:
//   pass A  one block per 1024 flags -> per-block survivor count
//   pass B  each block sums the counts of the blocks before it (they sit in L2), ranks its own flags
//           with ballots and writes the surviving INDICES; the last block publishes the total.
// n is read from device memory when n_dev != nullptr so a chain of stages needs no host round trip.
#pragma once
#include "common.cuh"

namespace limu {

constexpr int COMPACT_BLOCK = 1024;

__device__ __forceinline__ int block_exclusive_scan_flag(int flag, int *total, int *warp_sums /* 32 ints shared */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned bal = __ballot_sync(0xFFFFFFFFu, flag);
    const int in_warp = __popc(bal & ((1u << lane) - 1u));
    if (lane == 0) warp_sums[warp] = __popc(bal);
    __syncthreads();
    if (warp == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        int v = lane < nw ? warp_sums[lane] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += t; }
        warp_sums[lane] = incl - v;              // exclusive prefix per warp
        if (lane == 31) *total = incl;
    }
    __syncthreads();
    return warp_sums[warp] + in_warp;
}

static __global__ void __launch_bounds__(COMPACT_BLOCK) k_compact_count(const unsigned char *flags, int64_t n_max, const int *n_dev, int *block_counts) {
    __shared__ int ws[32];
    __shared__ int total;
    const int64_t n = n_dev ? (int64_t)*n_dev : n_max;
    const int64_t i = (int64_t)blockIdx.x * COMPACT_BLOCK + threadIdx.x;
    const int f = (i < n) ? (flags[i] != 0) : 0;
    block_exclusive_scan_flag(f, &total, ws);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = total;
}

static __global__ void __launch_bounds__(COMPACT_BLOCK) k_compact_scatter(const unsigned char *flags, int64_t n_max, const int *n_dev, const int *block_counts,
                                                                    int *out_idx, int *out_count) {
    __shared__ int ws[32];
    __shared__ int total;
    __shared__ int base_s;
    // base = sum of the counts of the preceding blocks
    int part = 0;
    for (int b = threadIdx.x; b < (int)blockIdx.x; b += COMPACT_BLOCK) part += block_counts[b];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_down_sync(0xFFFFFFFFu, part, o);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x < 32) {
        int v = ws[threadIdx.x];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xFFFFFFFFu, v, o);
        if (threadIdx.x == 0) base_s = v;
    }
    __syncthreads();
    const int base = base_s;
    __syncthreads();
    const int64_t n = n_dev ? (int64_t)*n_dev : n_max;
    const int64_t i = (int64_t)blockIdx.x * COMPACT_BLOCK + threadIdx.x;
    const int f = (i < n) ? (flags[i] != 0) : 0;
    const int rank = block_exclusive_scan_flag(f, &total, ws);
    if (f) out_idx[base + rank] = (int)i;
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) *out_count = base + total;
}

// flags[n_max] -> out_idx (survivor indices in order), *out_count. block_counts needs div_up(n_max,1024) ints.
inline int compact_flags(limu_ctx *c, const unsigned char *flags, int64_t n_max, const int *n_dev, int *block_counts, int *out_idx, int *out_count) {
    const int blocks = n_max > 0 ? div_up(n_max, COMPACT_BLOCK) : 1;
    k_compact_count<<<blocks, COMPACT_BLOCK, 0, c->stream>>>(flags, n_max, n_dev, block_counts);
    LIMU_LAUNCHED();
    k_compact_scatter<<<blocks, COMPACT_BLOCK, 0, c->stream>>>(flags, n_max, n_dev, block_counts, out_idx, out_count);
    LIMU_LAUNCHED();
    return LIMU_OK;
}

// out[j] = in[idx[j]] for j < *count (3 doubles per point).
static __global__ void k_gather_points(const double *in, const int *idx, const int *count, double *out, int64_t *out_idx64) {
    const int n = *count;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        const int i = idx[j];
        out[3 * (size_t)j] = in[3 * (size_t)i];
        out[3 * (size_t)j + 1] = in[3 * (size_t)i + 1];
        out[3 * (size_t)j + 2] = in[3 * (size_t)i + 2];
        if (out_idx64) out_idx64[j] = i;
    }
}

}  // namespace limu
