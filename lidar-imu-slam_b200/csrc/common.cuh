// common.cuh -- internal plumbing of liblimu_cuda: error state, context, device buffers, launch counting.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <string>

#include "../../include/limu_cuda.h"
#include "se3.cuh"

// Instrumented build only (make phase): every translation unit with kernels worth tracing keeps a ring of (id << 56 | globaltimer) records,
// appended by thread 0 of CTA 0, and exports a getter; tools/frame_phase_timing.py merges the rings into one timeline.
#ifdef LIMU_ICP_PHASE_TIMING
#define LIMU_TRACE_RING(getter)                                                                         \
    static __device__ unsigned long long g_trace[2048];                                                 \
    static __device__ unsigned int g_trace_n;                                                           \
    extern "C" int getter(unsigned long long *out /* 2048 */, unsigned int *n) {                        \
        if (cudaMemcpyFromSymbol(out, g_trace, sizeof(unsigned long long) * 2048) != cudaSuccess) return -1; \
        if (cudaMemcpyFromSymbol(n, g_trace_n, sizeof(unsigned int)) != cudaSuccess) return -1;         \
        return 0;                                                                                       \
    }
#define LIMU_TRACE(id)                                                                                  \
    do {                                                                                                \
        if (blockIdx.x == 0 && threadIdx.x == 0) {                                                      \
            unsigned long long _t;                                                                      \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(_t));                                      \
            g_trace[atomicAdd(&g_trace_n, 1u) & 2047u] = ((unsigned long long)(id) << 56) | (_t & 0x00FFFFFFFFFFFFFFull); \
        }                                                                                               \
    } while (0)
#else
#define LIMU_TRACE_RING(getter)
#define LIMU_TRACE(id) do {} while (0)
#endif

namespace limu {

void set_error(const char *fmt, ...);
extern std::atomic<uint64_t> g_launches;

#define LIMU_CUDA_TRY(expr)                                                                      \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess) {                                                                 \
            limu::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return LIMU_ERR_CUDA;                                                                \
        }                                                                                        \
    } while (0)

#define LIMU_TRY(expr)                  \
    do {                                \
        int _s = (expr);                \
        if (_s != LIMU_OK) return _s;   \
    } while (0)

#define LIMU_REQUIRE(cond, msg)                              \
    do {                                                     \
        if (!(cond)) { limu::set_error("%s", msg); return LIMU_ERR_INVALID; } \
    } while (0)

// Count a launch and check it was accepted.
#define LIMU_LAUNCHED()                                     \
    do {                                                    \
        limu::g_launches.fetch_add(1, std::memory_order_relaxed); \
        LIMU_CUDA_TRY(cudaGetLastError());                  \
    } while (0)

// Growable device buffer (never shrinks; contents are NOT preserved on growth unless asked).
struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
    int reserve(size_t want, cudaStream_t s = 0, bool keep = false) {
        if (want <= bytes) return LIMU_OK;
        size_t nb = bytes ? bytes : 256;
        while (nb < want) nb *= 2;
        void *np = nullptr;
        LIMU_CUDA_TRY(cudaMalloc(&np, nb));
        if (keep && p && bytes) LIMU_CUDA_TRY(cudaMemcpyAsync(np, p, bytes, cudaMemcpyDeviceToDevice, s));
        if (p) { LIMU_CUDA_TRY(cudaStreamSynchronize(s)); cudaFree(p); }
        p = np; bytes = nb;
        return LIMU_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
    template <class T> T *as() const { return static_cast<T *>(p); }
};

// Device-side status word shared by kernels of one context.
// SMs a cooperative launch of the pipelined path leaves free for k_gate threads (voxelize.cu) that may be resident while its CTAs are placed (one gate per
// context stream; a few more for other contexts on the same GPU)
constexpr int GATE_SLACK_SMS = 4;

struct DevStatus {
    int key_range;    // some voxel index fell outside the packed key range
    int table_full;   // an insert probe ran through the whole table
    int pad[2];       // pad[0]: a ring index exceeded num_scan_lines in the constant-rotation preprocessing path
};

}  // namespace limu

// Multi-GPU exchange state of one rank (comm.cu).
struct limu_comm {
    int rank = 0, nranks = 1;
    double *mbox_local = nullptr;          // [4][8][24] doubles in this rank's memory (IPC-exported)
    double *mbox_peer[8] = {};             // peer mappings of every rank's mailbox ([rank] == mbox_local)
    bool peer_opened[8] = {};
    unsigned long long stamp_base = 0;     // exchanges completed so far; advances identically on every rank
    int *d_error = nullptr;                // device flag: a peer timed out
    double *d_state = nullptr;             // NCCL baseline: [0..19] sums, [24..30] E, [32..38] T_icp, [40] done, [41] iter
    void *nccl_lib = nullptr, *nccl_comm = nullptr;
};

struct limu_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    limu::DevStatus *d_status = nullptr;   // device
    limu::DevStatus *h_status = nullptr;   // pinned host mirror
    // scratch pools reused by the stateless entry points and the pipeline
    limu::DevBuf in0, in1, out0, out1, out2, tmp0, tmp1, tmp2, tmp3, tmp4, tmp5;
    limu::DevBuf ll_rows;      // per-CTA rows of the stand-alone ICP entry points (stamped words, see registration.cu)
    void *h_pinned = nullptr;  // small pinned staging area for scalars / poses / counts
    size_t h_pinned_bytes = 0;
    limu::DevBuf d_small;      // small device staging area (poses, counts, partial sums)
    limu_comm *comm = nullptr;
    unsigned int ll_seq = 1;   // stamp counter of the row exchange in the registration kernels (31 bits, advances with every launch)
    // optional per-stage event timing (limu_ctx_set_profiling)
    bool profiling = false;
    cudaEvent_t ev[LIMU_NUM_STAGES][2] = {};
    bool ev_used[LIMU_NUM_STAGES] = {};
    double stage_ms[LIMU_NUM_STAGES] = {};
    int64_t profiled_frames = 0;
};

namespace limu {
// Bind the calling thread to the context's device (cheap when already current).
inline int bind(limu_ctx *c) {
    if (!c) { set_error("null context"); return LIMU_ERR_INVALID; }
    LIMU_CUDA_TRY(cudaSetDevice(c->device));
    return LIMU_OK;
}
int check_status(limu_ctx *c);   // sync + read DevStatus; maps flags to limu_status and clears them
int status_to_error(limu_ctx *c, const DevStatus &s);   // the same for a status word the caller already copied to the host
// out <- T * in for n points (n read from *n_dev when given); pose7 is DEVICE memory.
int transform_device(limu_ctx *c, const double *pose_dev, const double *in, double *out, int64_t n_max, const int *n_dev);
// H2D helpers on the context's stream.
int stage_in(limu_ctx *c, DevBuf &buf, const void *host, size_t bytes);
int stage_small(limu_ctx *c, const double *host, int count, int off, double **dev);
inline int prof_begin(limu_ctx *c, int stage) {
    if (c->profiling) { LIMU_CUDA_TRY(cudaEventRecord(c->ev[stage][0], c->stream)); c->ev_used[stage] = true; }
    return LIMU_OK;
}
inline int prof_end(limu_ctx *c, int stage) {
    if (c->profiling) LIMU_CUDA_TRY(cudaEventRecord(c->ev[stage][1], c->stream));
    return LIMU_OK;
}
int prof_collect(limu_ctx *c);   // after a stream sync: fold the recorded stage times into the accumulators
inline int div_up(int64_t a, int64_t b) { return static_cast<int>((a + b - 1) / b); }
}  // namespace limu
