// comm.cu -- multi-GPU plumbing for the point-sharded ICP (BASELINE configs[4]).
// One process per GPU. Each rank allocates a small mailbox in its own HBM and exports it with CUDA IPC; after the
// handles have been exchanged (torch.distributed / MPI / anything -- the bytes are opaque) every rank maps every
// peer's mailbox, and k_icp_persistent stores its per-iteration row of 20 doubles straight into the peers' memory
// over NVLink (registration.cu). NCCL is loaded lazily with dlopen and is only used by the un-fused baseline.
#include <dlfcn.h>

#include "common.cuh"

namespace limu {

constexpr int MBOX_DOUBLES = 4 * 8 * 24;   // MBOX_RING x MBOX_MAX_RANKS x MBOX_ROW (registration.cu)

struct NcclId128 { char b[128]; };   // ncclUniqueId is passed BY VALUE to ncclCommInitRank (nccl.h: char internal[128])
struct NcclApi {
    int (*GetUniqueId)(void *) = nullptr;
    int (*CommInitRank)(void **, int, NcclId128, int) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;
static void *g_nccl_lib = nullptr;

static int load_nccl() {
    if (g_nccl_lib) return LIMU_OK;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) { g_nccl_lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (g_nccl_lib) break; }
    if (!g_nccl_lib) { set_error("cannot load libnccl.so.2: %s", dlerror()); return LIMU_ERR_COMM; }
    *(void **)&g_nccl.GetUniqueId = dlsym(g_nccl_lib, "ncclGetUniqueId");
    *(void **)&g_nccl.CommInitRank = dlsym(g_nccl_lib, "ncclCommInitRank");
    *(void **)&g_nccl.AllReduce = dlsym(g_nccl_lib, "ncclAllReduce");
    *(void **)&g_nccl.CommDestroy = dlsym(g_nccl_lib, "ncclCommDestroy");
    *(void **)&g_nccl.GetErrorString = dlsym(g_nccl_lib, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce) { set_error("libnccl lacks the expected symbols"); return LIMU_ERR_COMM; }
    return LIMU_OK;
}

int comm_nccl_allreduce_sum_f64(limu_ctx *c, double *buf, size_t count) {
    limu_comm *cm = c->comm;
    const int rc = g_nccl.AllReduce(buf, buf, count, /* ncclFloat64 */ 8, /* ncclSum */ 0, cm->nccl_comm, c->stream);
    if (rc != 0) { set_error("ncclAllReduce failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?"); return LIMU_ERR_COMM; }
    return LIMU_OK;
}

}  // namespace limu

using namespace limu;

extern "C" {

int limu_comm_create(limu_ctx *c, int rank, int nranks, unsigned char ipc_handle_out[64]) {
    LIMU_TRY(bind(c));
    LIMU_REQUIRE(nranks >= 1 && nranks <= 8 && rank >= 0 && rank < nranks && ipc_handle_out, "limu_comm_create: 1 <= nranks <= 8, 0 <= rank < nranks");
    LIMU_REQUIRE(!c->comm, "limu_comm_create: this context already has a communicator");
    limu_comm *cm = new limu_comm;
    cm->rank = rank; cm->nranks = nranks;
    LIMU_CUDA_TRY(cudaMalloc(&cm->mbox_local, MBOX_DOUBLES * sizeof(double)));
    LIMU_CUDA_TRY(cudaMemset(cm->mbox_local, 0, MBOX_DOUBLES * sizeof(double)));
    LIMU_CUDA_TRY(cudaMalloc(&cm->d_error, sizeof(int)));
    LIMU_CUDA_TRY(cudaMemset(cm->d_error, 0, sizeof(int)));
    LIMU_CUDA_TRY(cudaMalloc(&cm->d_state, 64 * sizeof(double)));
    LIMU_CUDA_TRY(cudaMemset(cm->d_state, 0, 64 * sizeof(double)));
    cm->mbox_peer[rank] = cm->mbox_local;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    LIMU_CUDA_TRY(cudaIpcGetMemHandle(&h, cm->mbox_local));
    memcpy(ipc_handle_out, &h, 64);
    c->comm = cm;
    return LIMU_OK;
}

int limu_comm_connect(limu_ctx *c, const unsigned char *all_handles) {
    LIMU_TRY(bind(c));
    LIMU_REQUIRE(c->comm && all_handles, "limu_comm_connect: create the communicator first");
    limu_comm *cm = c->comm;
    for (int r = 0; r < cm->nranks; ++r) {
        if (r == cm->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, all_handles + 64 * r, 64);
        void *p = nullptr;
        LIMU_CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        cm->mbox_peer[r] = static_cast<double *>(p);
        cm->peer_opened[r] = true;
    }
    return LIMU_OK;
}

void limu_comm_destroy(limu_ctx *c) {
    if (!c || !c->comm) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    limu_comm *cm = c->comm;
    if (cm->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(cm->nccl_comm);
    for (int r = 0; r < cm->nranks; ++r) if (cm->peer_opened[r]) cudaIpcCloseMemHandle(cm->mbox_peer[r]);
    cudaFree(cm->mbox_local); cudaFree(cm->d_error); cudaFree(cm->d_state);
    delete cm;
    c->comm = nullptr;
}

int limu_comm_nccl_unique_id(unsigned char id_out[128]) {
    LIMU_TRY(load_nccl());
    LIMU_REQUIRE(id_out, "limu_comm_nccl_unique_id: null");
    const int rc = g_nccl.GetUniqueId(id_out);
    if (rc != 0) { set_error("ncclGetUniqueId failed (%d)", rc); return LIMU_ERR_COMM; }
    return LIMU_OK;
}

int limu_comm_nccl_init(limu_ctx *c, const unsigned char id[128]) {
    LIMU_TRY(bind(c));
    LIMU_REQUIRE(c->comm && id, "limu_comm_nccl_init: create the communicator first");
    LIMU_TRY(load_nccl());
    NcclId128 u;
    memcpy(u.b, id, 128);
    const int rc = g_nccl.CommInitRank(&c->comm->nccl_comm, c->comm->nranks, u, c->comm->rank);
    if (rc != 0) { set_error("ncclCommInitRank failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?"); return LIMU_ERR_COMM; }
    return LIMU_OK;
}

}  // extern "C"
