// bulk_gather_bench.cu -- stand-alone micro-benchmark for the next round; NOT part of the library.
// Question: the kernel-mode registration kernel gathers one 128-byte-aligned voxel block (three rows x/y/z of `capp` doubles, 384 or
// 512 bytes) per query from a map much larger than L2. Today eight lanes fetch it with 3 x ROUNDS dependent LDG.64 each (coop_scan in
// registration.cu). Does ONE cp.async.bulk (TMA 1-D bulk copy, mbarrier completion) per block into a per-warp shared-memory ring keep
// more bytes in flight and get closer to the HBM peak?
//   A  eight lanes per block, LDG.64 per lane (what the kernel does now)
//   B  per-warp ring of DEPTH slots; lane 0 issues one bulk copy per block; the warp consumes the blocks from shared memory
// Both compute the same thing: for each query the (min squared distance, rank) over the block's `count` points, summed into a checksum.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/bulk_gather tools/experimental/bulk_gather_bench.cu && /tmp/bulk_gather
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s (line %d)\n", #x, cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

constexpr int CAPP = 20;                 // cap 20 -> rows of 20 doubles
constexpr int STRIDE = 64;               // doubles per block (3 * 20 = 60 -> 64 = 512 bytes), as block_stride() in voxel_map.cuh
constexpr int BLOCK_BYTES = STRIDE * 8;
constexpr int THREADS = 256;
constexpr int DEPTH = 4;                 // ring slots per warp

__device__ __forceinline__ double sq3(double a, double b, double c) { return (a * a + b * b) + c * c; }

// ---- A: eight lanes per block ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(THREADS, 4) k_ldg(const double *__restrict__ blocks, const int *__restrict__ idx, const double *__restrict__ q, int64_t n,
                                                    double *__restrict__ out) {
    const int lane = threadIdx.x & 31, l8 = lane & 7;
    const int64_t groups = (int64_t)gridDim.x * (THREADS / 8), g0 = ((int64_t)blockIdx.x * THREADS + threadIdx.x) / 8;
    double acc = 0.0;
    for (int64_t i = g0; i < n; i += groups) {
        const double *b = blocks + (size_t)__ldg(idx + i) * STRIDE;
        const double qx = __ldg(q + 3 * i), qy = __ldg(q + 3 * i + 1), qz = __ldg(q + 3 * i + 2);
        double x[3], y[3], z[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) { const int r = l8 + 8 * k < CAPP ? l8 + 8 * k : l8; x[k] = __ldg(b + r); y[k] = __ldg(b + CAPP + r); z[k] = __ldg(b + 2 * CAPP + r); }
        double best = 1e300;
#pragma unroll
        for (int k = 0; k < 3; ++k) { const double d = sq3(qx - x[k], qy - y[k], qz - z[k]); if (l8 + 8 * k < CAPP && d < best) best = d; }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) { const double t = __shfl_xor_sync(0xFFFFFFFFu, best, o); best = t < best ? t : best; }
        if (l8 == 0) acc += best;
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
    if (lane == 0) atomicAdd(out, acc);
}

// ---- B: bulk copies into a per-warp ring ----------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n .reg .pred p;\n W_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra D_%=;\n bra W_%=;\n D_%=:\n}" ::"r"(bar), "r"(parity) : "memory");
}

__global__ void __launch_bounds__(THREADS, 4) k_bulk(const double *__restrict__ blocks, const int *__restrict__ idx, const double *__restrict__ q, int64_t n,
                                                     double *__restrict__ out) {
    __shared__ __align__(128) double ring[THREADS / 32][DEPTH][STRIDE];
    __shared__ __align__(8) unsigned long long bars[THREADS / 32][DEPTH];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t warps = (int64_t)gridDim.x * (THREADS / 32), w0 = (int64_t)blockIdx.x * (THREADS / 32) + warp;
    if (lane == 0)
        for (int s = 0; s < DEPTH; ++s) mbar_init(smem_u32(&bars[warp][s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    // prologue: fill the ring
    int64_t next = w0;
    for (int s = 0; s < DEPTH; ++s, next += warps)
        if (lane == 0 && next < n) {
            mbar_expect(smem_u32(&bars[warp][s]), BLOCK_BYTES);
            bulk_g2s(smem_u32(&ring[warp][s][0]), blocks + (size_t)__ldg(idx + next) * STRIDE, BLOCK_BYTES, smem_u32(&bars[warp][s]));
        }
    double acc = 0.0;
    int slot = 0;
    uint32_t parity = 0;
    for (int64_t i = w0; i < n; i += warps) {
        const double qx = __ldg(q + 3 * i), qy = __ldg(q + 3 * i + 1), qz = __ldg(q + 3 * i + 2);
        mbar_wait(smem_u32(&bars[warp][slot]), parity);
        const double *b = &ring[warp][slot][0];
        double best = 1e300;
        if (lane < CAPP) best = sq3(qx - b[lane], qy - b[CAPP + lane], qz - b[2 * CAPP + lane]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { const double t = __shfl_xor_sync(0xFFFFFFFFu, best, o); best = t < best ? t : best; }
        if (lane == 0) acc += best;
        __syncwarp();   // everybody has read the slot: refill it
        if (lane == 0 && next < n) {
            mbar_expect(smem_u32(&bars[warp][slot]), BLOCK_BYTES);
            bulk_g2s(smem_u32(&ring[warp][slot][0]), blocks + (size_t)__ldg(idx + next) * STRIDE, BLOCK_BYTES, smem_u32(&bars[warp][slot]));
        }
        next += warps;
        if (++slot == DEPTH) { slot = 0; parity ^= 1u; }
    }
    if (lane == 0) atomicAdd(out, acc);
}

int main(int argc, char **argv) {
    const int64_t n_blocks = argc > 1 ? atoll(argv[1]) : 4 * 1000 * 1000;   // 4 M blocks x 512 B = 2 GB >> L2
    const int64_t n = argc > 2 ? atoll(argv[2]) : 4 * 1024 * 1024;          // queries
    double *blocks, *q, *out;
    int *idx;
    CK(cudaMalloc(&blocks, (size_t)n_blocks * BLOCK_BYTES));
    CK(cudaMalloc(&q, (size_t)n * 24));
    CK(cudaMalloc(&idx, (size_t)n * 4));
    CK(cudaMalloc(&out, 2 * sizeof(double)));
    std::vector<int> hidx((size_t)n);
    std::vector<double> hq((size_t)3 * n);
    uint64_t s = 88172645463325252ull;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
    for (int64_t i = 0; i < n; ++i) { hidx[(size_t)i] = (int)(rnd() % (uint64_t)n_blocks); for (int a = 0; a < 3; ++a) hq[(size_t)(3 * i + a)] = (double)(rnd() % 1000) * 0.01; }
    CK(cudaMemcpy(idx, hidx.data(), (size_t)n * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(q, hq.data(), (size_t)n * 24, cudaMemcpyHostToDevice));
    {   // block contents: pseudo-random doubles (values do not matter for bandwidth, they must only differ)
        std::vector<double> chunk((size_t)1 << 20);
        for (size_t off = 0; off < (size_t)n_blocks * STRIDE; off += chunk.size()) {
            const size_t m = std::min(chunk.size(), (size_t)n_blocks * STRIDE - off);
            for (size_t k = 0; k < m; ++k) chunk[k] = (double)(rnd() % 100000) * 1e-3;
            CK(cudaMemcpy(blocks + off, chunk.data(), m * 8, cudaMemcpyHostToDevice));
        }
    }
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const int grid = sms * 4;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const double bytes = (double)n * (BLOCK_BYTES - 32 + 24 + 4);   // 480 useful block bytes + query + index
    for (int variant = 0; variant < 2; ++variant) {
        double sums[2] = {0, 0};
        float best_ms = 1e30f;
        for (int rep = 0; rep < 6; ++rep) {
            CK(cudaMemset(out, 0, 16));
            CK(cudaEventRecord(e0));
            if (variant == 0) k_ldg<<<grid, THREADS>>>(blocks, idx, q, n, out);
            else k_bulk<<<grid, THREADS>>>(blocks, idx, q, n, out);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            CK(cudaGetLastError());
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep > 0 && ms < best_ms) best_ms = ms;
            CK(cudaMemcpy(sums, out, 8, cudaMemcpyDeviceToHost));
        }
        printf("{\"variant\": \"%s\", \"queries\": %lld, \"blocks\": %lld, \"ms\": %.4f, \"GBps\": %.1f, \"checksum\": %.6f}\n", variant == 0 ? "ldg_8_lanes" : "bulk_ring",
               (long long)n, (long long)n_blocks, best_ms, bytes / best_ms * 1e-6, sums[0]);
    }
    return 0;
}
