// launch_gap_bench.cu -- how long the GPU idles between two kernels that sit back to back in ONE stream, for plain launches and for
// cooperative launches (cudaLaunchCooperativeKernel), and for a kernel on a second stream released by an event.
// The pipelined odometry path is a chain of short persistent kernels, so this gap is paid twice per scan.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o launch_gap_bench launch_gap_bench.cu && ./launch_gap_bench
#include <cuda_runtime.h>
#include <stdio.h>
#include <algorithm>
#include <vector>

__global__ void k_busy(unsigned long long *marks, int idx, unsigned int spin_ns) {
    unsigned long long t0, t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    if (blockIdx.x == 0 && threadIdx.x == 0) marks[2 * idx] = t0;
    do { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); } while (t - t0 < spin_ns);
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x == 0) marks[2 * idx + 1] = t;
}

static void report(const char *name, const std::vector<unsigned long long> &h, int n) {
    std::vector<double> gaps;
    for (int i = 8; i + 1 < n; ++i) gaps.push_back(((double)h[2 * (i + 1)] - (double)h[2 * i + 1]) / 1e3);
    std::sort(gaps.begin(), gaps.end());
    printf("%-64s gap end->start: min %.2f  median %.2f  p90 %.2f us\n", name, gaps.front(), gaps[gaps.size() / 2], gaps[gaps.size() * 9 / 10]);
}

int main() {
    const int N = 64;
    unsigned long long *d;
    cudaMalloc(&d, 2 * N * sizeof(unsigned long long));
    std::vector<unsigned long long> h(2 * N);
    cudaStream_t s1, s2;
    cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    for (int threads : {256, 1024}) {
        for (int grid : {sms, 89}) {
            char name[128];
            // plain launches
            for (int i = 0; i < N; ++i) k_busy<<<grid, threads, 0, s1>>>(d, i, 20000u);
            cudaStreamSynchronize(s1);
            cudaMemcpy(h.data(), d, h.size() * 8, cudaMemcpyDeviceToHost);
            snprintf(name, sizeof name, "plain launches, grid %d x %d threads", grid, threads);
            report(name, h, N);
            // cooperative launches
            for (int i = 0; i < N; ++i) {
                unsigned int ns = 20000u;
                void *args[] = {&d, &i, &ns};
                cudaLaunchCooperativeKernel((const void *)k_busy, dim3(grid), dim3(threads), args, 0, s1);
            }
            cudaStreamSynchronize(s1);
            cudaMemcpy(h.data(), d, h.size() * 8, cudaMemcpyDeviceToHost);
            snprintf(name, sizeof name, "cooperative launches, grid %d x %d threads", grid, threads);
            report(name, h, N);
        }
    }
    // alternating streams with an event between consecutive kernels (what a cross-stream dependency costs)
    {
        cudaEvent_t ev[N];
        for (int i = 0; i < N; ++i) cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming);
        for (int coop = 0; coop < 2; ++coop) {
            for (int i = 0; i < N; ++i) {
                cudaStream_t s = (i & 1) ? s2 : s1;
                if (i > 0) cudaStreamWaitEvent(s, ev[i - 1], 0);
                unsigned int ns = 20000u;
                void *args[] = {&d, &i, &ns};
                if (coop) cudaLaunchCooperativeKernel((const void *)k_busy, dim3(sms), dim3(256), args, 0, s);
                else k_busy<<<sms, 256, 0, s>>>(d, i, ns);
                cudaEventRecord(ev[i], s);
            }
            cudaDeviceSynchronize();
            cudaMemcpy(h.data(), d, h.size() * 8, cudaMemcpyDeviceToHost);
            report(coop ? "cooperative, alternating streams, event between kernels" : "plain, alternating streams, event between kernels", h, N);
        }
    }
    printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
