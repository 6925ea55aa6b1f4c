// grid_sync.cuh -- grid-wide barrier for cooperative (co-resident) launches: one monotonically increasing counter,
// release/acquire at GPU scope. `phase` is a per-thread running count of barriers passed (uniform across the grid).
#pragma once
#include <cuda_runtime.h>

namespace limu {

__device__ __forceinline__ void grid_barrier(unsigned int *bar, unsigned int target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(bar, 1u);
        unsigned int v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
        } while (v < target);
    }
    __syncthreads();
}

struct GridSync {
    unsigned int *bar;
    unsigned int phase;
    unsigned int members;   // CTAs taking part (gridDim.x, or the leading subset that runs a latency-bound phase)
    __device__ __forceinline__ void sync() { ++phase; grid_barrier(bar, phase * members); }
};

}  // namespace limu
