// api.cu -- context, error state and host-side helpers of the C ABI (include/limu_cuda.h).
#include <stdarg.h>

#include <vector>

#include "common.cuh"

namespace limu {

static thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

void release_ctx_scratch(limu_ctx *c);
void release_pre_scratch(limu_ctx *c);

int status_to_error(limu_ctx *c, const DevStatus &s);

int check_status(limu_ctx *c) {
    LIMU_CUDA_TRY(cudaMemcpyAsync(c->h_status, c->d_status, sizeof(DevStatus), cudaMemcpyDeviceToHost, c->stream));
    LIMU_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return status_to_error(c, *c->h_status);
}

// Map a status word that is already on the host to a limu_status; clears the device word when something was flagged.
int status_to_error(limu_ctx *c, const DevStatus &s) {
    if (s.key_range || s.table_full || s.pad[0]) {
        LIMU_CUDA_TRY(cudaMemsetAsync(c->d_status, 0, sizeof(DevStatus), c->stream));
        if (s.pad[0]) { set_error("a point's ring index is >= num_scan_lines (the reference indexes its per-ring state out of bounds here)"); return LIMU_ERR_INVALID; }
        if (s.key_range) { set_error("voxel index outside the packed key range (|index| >= 2^20): point too far for this voxel size"); return LIMU_ERR_KEY_RANGE; }
        set_error("voxel hash table full");
        return LIMU_ERR_MAP_FULL;
    }
    return LIMU_OK;
}

int prof_collect(limu_ctx *c) {
    if (!c->profiling) return LIMU_OK;
    for (int s = 0; s < LIMU_NUM_STAGES; ++s) {
        if (!c->ev_used[s]) continue;
        float ms = 0.f;
        LIMU_CUDA_TRY(cudaEventElapsedTime(&ms, c->ev[s][0], c->ev[s][1]));
        c->stage_ms[s] += (double)ms;
        c->ev_used[s] = false;
    }
    ++c->profiled_frames;
    return LIMU_OK;
}

}  // namespace limu

using namespace limu;

extern "C" {

int limu_ctx_set_profiling(limu_ctx *c, int enabled) {
    LIMU_TRY(bind(c));
    if (enabled && !c->ev[0][0])
        for (int s = 0; s < LIMU_NUM_STAGES; ++s) for (int k = 0; k < 2; ++k) LIMU_CUDA_TRY(cudaEventCreate(&c->ev[s][k]));
    c->profiling = enabled != 0;
    for (int s = 0; s < LIMU_NUM_STAGES; ++s) { c->stage_ms[s] = 0.0; c->ev_used[s] = false; }
    c->profiled_frames = 0;
    return LIMU_OK;
}
int limu_ctx_get_profile(limu_ctx *c, double ms[LIMU_NUM_STAGES], int64_t frames[1]) {
    LIMU_REQUIRE(c && ms && frames, "limu_ctx_get_profile: null argument");
    for (int s = 0; s < LIMU_NUM_STAGES; ++s) ms[s] = c->stage_ms[s];
    frames[0] = c->profiled_frames;
    return LIMU_OK;
}

const char *limu_last_error(void) { return g_err; }
int limu_abi_version(void) { return LIMU_ABI_VERSION; }
#ifndef LIMU_SOURCE_HASH
#define LIMU_SOURCE_HASH "unknown"
#endif
const char *limu_source_hash(void) { return LIMU_SOURCE_HASH; }
uint64_t limu_kernel_launches(void) { return g_launches.load(); }

int limu_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

void *limu_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
        set_error("cudaHostAlloc(%zu) failed: %s", bytes, cudaGetErrorString(cudaGetLastError()));
        return nullptr;
    }
    return p;
}
void limu_host_free(void *p) { if (p) cudaFreeHost(p); }

int limu_ctx_create(int device, limu_ctx **out) {
    LIMU_REQUIRE(out, "limu_ctx_create: out is null");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        set_error("no usable CUDA device (%s); liblimu_cuda has no CPU fallback", e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        cudaGetLastError();
        return LIMU_ERR_CUDA;
    }
    LIMU_REQUIRE(device >= 0 && device < n, "limu_ctx_create: device index out of range");
    LIMU_CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    LIMU_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        set_error("device %d is sm_%d%d; liblimu_cuda is built for sm_100a (B200) only", device, prop.major, prop.minor);
        return LIMU_ERR_CUDA;
    }
    int coop = 0;
    LIMU_CUDA_TRY(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device));
    if (!coop) { set_error("device %d does not support cooperative launch", device); return LIMU_ERR_CUDA; }
    limu_ctx *c = new limu_ctx;
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    LIMU_CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    LIMU_CUDA_TRY(cudaHostAlloc(&c->h_status, sizeof(DevStatus), cudaHostAllocDefault));
    c->h_pinned_bytes = 8192;
    LIMU_CUDA_TRY(cudaHostAlloc(&c->h_pinned, c->h_pinned_bytes, cudaHostAllocDefault));
    LIMU_TRY(c->d_small.reserve(8192));
    LIMU_CUDA_TRY(cudaMemset(c->d_small.p, 0, 8192));
    // the device status word lives inside the small area, right behind the pipeline's per-scan result block ([32..52]), so one
    // device-to-host copy per scan brings back counts, pose, loop statistics AND the status
    c->d_status = reinterpret_cast<DevStatus *>(c->d_small.as<double>() + 53);
    *out = c;
    return LIMU_OK;
}

void limu_ctx_destroy(limu_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    limu_comm_destroy(c);
    release_ctx_scratch(c);
    release_pre_scratch(c);
    limu::DevBuf *bufs[] = {&c->in0, &c->in1, &c->out0, &c->out1, &c->out2, &c->tmp0, &c->tmp1, &c->tmp2, &c->tmp3, &c->tmp4, &c->tmp5, &c->ll_rows, &c->d_small};
    for (auto *b : bufs) b->release();
    for (int s = 0; s < LIMU_NUM_STAGES; ++s) for (int k = 0; k < 2; ++k) if (c->ev[s][k]) cudaEventDestroy(c->ev[s][k]);
    cudaFreeHost(c->h_status);
    cudaFreeHost(c->h_pinned);
    cudaStreamDestroy(c->stream);
    delete c;
}

int limu_ctx_sync(limu_ctx *c) {
    LIMU_TRY(bind(c));
    LIMU_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return LIMU_OK;
}
void *limu_ctx_stream(limu_ctx *c) { return c ? (void *)c->stream : nullptr; }

void limu_se3_exp(const double x[6], double pose_out[7]) { pose_store(se3_exp(x), pose_out); }
void limu_se3_log(const double pose[7], double x_out[6]) { se3_log(pose_load(pose), x_out); }
void limu_se3_mul(const double a[7], const double b[7], double out[7]) { pose_store(mul(pose_load(a), pose_load(b)), out); }
void limu_se3_inverse(const double a[7], double out[7]) { pose_store(inverse(pose_load(a)), out); }

}  // extern "C"
