// voxel_map.cu -- GPU-resident voxel hash map: build, capped ordered insertion, eviction, queries, dump.
// Replaces lidar::VoxelHashMap / lidar::VoxelBlock (L/src/sensors/lidar/helpers/voxel_hash_map.cpp,
// voxel_block.cpp). Layout and lookup rules: voxel_map.cuh.
#include <math.h>

#include <algorithm>
#include <vector>

#include "compact.cuh"
#include "voxel_map.cuh"

namespace limu {

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------
static __global__ void k_map_clear(MapView m, int64_t C) {
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < C; s += (int64_t)gridDim.x * blockDim.x)
        *slot_at(m, (unsigned int)s) = Slot{KEY_EMPTY, META_NONE};
}

static __global__ void __launch_bounds__(256) k_insert_claim(MapView m, const double *__restrict__ xyz, int64_t n_max, const int *n_dev,
                                                            unsigned long long birth_base, unsigned int *__restrict__ pslot,
                                                            unsigned long long *counters, DevStatus *st) {
    const int64_t n = n_dev ? (int64_t)*n_dev : n_max;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool claimed = false;
    unsigned int slot = PEND_NONE;
    if (i < n) pslot[i] = slot = insert_claim_one(m, V3{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]}, (unsigned int)i, birth_base, st, &claimed);
    insert_account(claimed, slot, counters, m);
}

static __global__ void __launch_bounds__(256) k_insert_place(MapView m, const double *__restrict__ xyz, int64_t n_max, const int *n_dev,
                                                            const unsigned int *__restrict__ pslot) {
    const int64_t n = n_dev ? (int64_t)*n_dev : n_max;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    insert_place_one(m, V3{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]}, (unsigned int)i, pslot[i]);
}

// Walks the dense live list (one entry per voxel ever created since the last rebuild) instead of the C table slots.
static __global__ void __launch_bounds__(256) k_remove_far(MapView m, const double *__restrict__ origin, double max_distance, unsigned long long *counters) {
    const int64_t used = (int64_t)counters[3];
    const double ox = origin[0], oy = origin[1], oz = origin[2];
    const int ovx = vox_index(m, ox), ovy = vox_index(m, oy), ovz = vox_index(m, oz);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < used; i += (int64_t)gridDim.x * blockDim.x)
        remove_far_entry(m, i, __ldcg(m.live_key + i), ovx, ovy, ovz, ox, oy, oz, max_distance, counters);
}

// Move every live voxel of `old` into the (cleared) table `nw`.
static __global__ void __launch_bounds__(256) k_rehash(MapView old, int64_t oldC, MapView nw, unsigned long long *new_counters) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= oldC) return;
    const unsigned long long key = slot_at(old, (unsigned int)s)->key;
    if (key >= KEY_TOMB) return;
    unsigned int t = slot_of(key, nw.shift);
    for (;;) {
        const unsigned long long cur = atomicCAS(&slot_at(nw, t)->key, KEY_EMPTY, key);
        if (cur == KEY_EMPTY) break;
        t = (t + 1) & nw.mask;
    }
    {
        const unsigned long long at = atomicAdd(&new_counters[3], 1ull);   // the rebuilt list holds live voxels only
        nw.live[at] = t;
        nw.live_key[at] = key;
    }
    const unsigned long long meta = slot_at(old, (unsigned int)s)->meta;
    const int count = meta_count(meta);
    slot_at(nw, t)->meta = meta;
    const double *src = voxel_rows(old, (unsigned int)s);
    double *dst = voxel_rows(nw, t);
    for (int r = 0; r < count; ++r) { dst[r] = src[r]; dst[nw.capp + r] = src[old.capp + r]; dst[2 * nw.capp + r] = src[2 * old.capp + r]; }
}

// Both walk the dense live list (counters[3] entries; erased voxels read KEY_TOMB) instead of the C block headers.
static __global__ void k_sum_counts(MapView m, const unsigned long long *counters, unsigned long long *out /* [0]=voxels [1]=points */) {
    const int64_t used = (int64_t)counters[3];
    unsigned long long v = 0, p = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < used; i += (int64_t)gridDim.x * blockDim.x) {
        const Slot *sl = slot_at(m, m.live[i]);
        if (sl->key < KEY_TOMB) { ++v; p += (unsigned long long)meta_count(sl->meta); }
    }
    for (int o = 16; o > 0; o >>= 1) { v += __shfl_down_sync(0xFFFFFFFFu, v, o); p += __shfl_down_sync(0xFFFFFFFFu, p, o); }
    if ((threadIdx.x & 31) == 0 && (v | p)) { atomicAdd(&out[0], v); atomicAdd(&out[1], p); }
}

// Live slots -> (birth, slot) pairs, unordered; the host sorts by birth to recover creation order.
static __global__ void k_collect_live(MapView m, const unsigned long long *counters, unsigned long long *pairs, unsigned long long *n_out) {
    const int64_t used = (int64_t)counters[3];
    for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x; i0 < used; i0 += (int64_t)gridDim.x * blockDim.x) {   // whole warps stay converged for the ballot
        const int64_t i = i0 + threadIdx.x;
        unsigned int s = 0;
        bool live = false;
        if (i < used) { s = m.live[i]; live = slot_at(m, s)->key < KEY_TOMB; }
        const unsigned bal = __ballot_sync(0xFFFFFFFFu, live);
        unsigned long long base = 0;
        const int lane = threadIdx.x & 31;
        if (lane == 0 && bal) base = atomicAdd(n_out, (unsigned long long)__popc(bal));
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (live) {
            const unsigned long long j = base + __popc(bal & ((1u << lane) - 1u));
            pairs[2 * j] = meta_birth(slot_at(m, s)->meta);
            pairs[2 * j + 1] = (unsigned long long)s;
        }
    }
}

static __global__ void k_gather_voxels(MapView m, const unsigned int *order, int64_t nv, int *keys, int *counts, double *pts) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nv) return;
    const unsigned int s = order[j];
    int x, y, z;
    unpack_key(slot_at(m, s)->key, x, y, z);
    keys[3 * j] = x; keys[3 * j + 1] = y; keys[3 * j + 2] = z;
    const int c = meta_count(slot_at(m, s)->meta);
    counts[j] = c;
    const double *b = voxel_rows(m, s);
    for (int r = 0; r < c; ++r) {   // back to array-of-structs for the host
        double *o = pts + ((size_t)j * m.cap + r) * 3;
        o[0] = b[r]; o[1] = b[m.capp + r]; o[2] = b[2 * m.capp + r];
    }
}

// get_closest_neighbour for a batch (voxel_hash_map.cpp:64-102); flag = within max_correspondance (:120).
static __global__ void __launch_bounds__(256) k_closest(MapView m, const double *__restrict__ xyz, int64_t n_max, const int *n_dev, double max_sq,
                                                       double *__restrict__ out_xyz, int *__restrict__ out_key, int *__restrict__ out_rank,
                                                       unsigned char *__restrict__ flags, int nn27) {
    const int64_t n = n_dev ? (int64_t)*n_dev : n_max;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const V3 p{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
    const Nearest r = nn27 ? map_closest27(m, p) : map_closest(m, p);
    if (out_xyz) { out_xyz[3 * i] = r.x; out_xyz[3 * i + 1] = r.y; out_xyz[3 * i + 2] = r.z; }
    if (out_key) {
        int x = INT32_MIN, y = INT32_MIN, z = INT32_MIN;
        if (r.slot >= 0 && r.rank >= 0) unpack_key(slot_at(m, (unsigned int)r.slot)->key, x, y, z);
        out_key[3 * i] = x; out_key[3 * i + 1] = y; out_key[3 * i + 2] = z;
    }
    if (out_rank) out_rank[i] = r.rank;
    if (flags) flags[i] = sqnorm3(r.x - p.x, r.y - p.y, r.z - p.z) < max_sq ? 1 : 0;   // (found - point).squaredNorm()
}

static __global__ void k_transform(const double *__restrict__ pose7, const double *__restrict__ in, double *__restrict__ out, int64_t n_max, const int *n_dev) {
    const int64_t n = n_dev ? (int64_t)*n_dev : n_max;
    const Pose T = pose_load(pose7);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const V3 q = apply(T, V3{in[3 * i], in[3 * i + 1], in[3 * i + 2]});
        out[3 * i] = q.x; out[3 * i + 1] = q.y; out[3 * i + 2] = q.z;
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int64_t next_pow2(int64_t v) { int64_t p = 1024; while (p < v) p <<= 1; return p; }

int transform_device(limu_ctx *c, const double *pose_dev, const double *in, double *out, int64_t n_max, const int *n_dev) {
    if (n_max <= 0) return LIMU_OK;
    const int blocks = std::min<int64_t>(div_up(n_max, 256), (int64_t)c->sm_count * 16);
    k_transform<<<blocks, 256, 0, c->stream>>>(pose_dev, in, out, n_max, n_dev);
    LIMU_LAUNCHED();
    return LIMU_OK;
}

int map_alloc(limu_map *m, int64_t C) {
    limu_ctx *c = m->ctx;
    m->blk.release(); m->pend.release(); m->live.release();
    LIMU_TRY(m->blk.reserve((size_t)C * limu::block_stride(m->cap) * 8));
    LIMU_TRY(m->live.reserve((size_t)C * 12));   // C slot indices (4 B) followed by their C keys (8 B)
    LIMU_TRY(m->pend.reserve((size_t)C * m->cap * 4));
    m->capacity = C;
    const int blocks = std::min<int64_t>(div_up(C, 256), (int64_t)c->sm_count * 32);
    k_map_clear<<<blocks, 256, 0, c->stream>>>(m->view(), C);
    LIMU_LAUNCHED();
    LIMU_CUDA_TRY(cudaMemsetAsync(m->pend.p, 0xFF, (size_t)C * m->cap * 4, c->stream));
    LIMU_TRY(m->counters.reserve(8 * sizeof(unsigned long long)));
    LIMU_CUDA_TRY(cudaMemsetAsync(m->counters.p, 0, 8 * sizeof(unsigned long long), c->stream));
    m->used_upper = 0;
    return LIMU_OK;
}

}  // namespace limu

limu::MapView limu_map::view() const {
    limu::MapView v;
    v.blk = blk.as<double>();
    v.pend = pend.as<unsigned int>();
    v.live = live.as<unsigned int>();
    v.live_key = reinterpret_cast<unsigned long long *>(v.live + capacity);   // (capacity is a power of two >= 1024: 8-byte aligned)
    v.mask = (unsigned int)(capacity - 1);
    int lg = 0;
    while ((int64_t(1) << lg) < capacity) ++lg;
    v.shift = 64 - lg;
    v.cap = cap;
    v.capp = limu::cap_padded(cap);
    v.stride = limu::block_stride(cap);
    v.vox = vox_size;
    {   // power of two? (mantissa exactly 0.5)
        int e = 0;
        v.inv_vox = (vox_size > 0.0 && frexp(vox_size, &e) == 0.5) ? 1.0 / vox_size : 0.0;
    }
    return v;
}

namespace limu {

static int read_counters(limu_map *m, unsigned long long out[4]) {
    limu_ctx *c = m->ctx;
    unsigned long long *h = static_cast<unsigned long long *>(c->h_pinned);
    LIMU_CUDA_TRY(cudaMemcpyAsync(h, m->counters.p, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    LIMU_CUDA_TRY(cudaStreamSynchronize(c->stream));
    for (int i = 0; i < 4; ++i) out[i] = h[i];
    return LIMU_OK;
}

// Keep load (live + tombstones + incoming) <= 1/2: rebuild into a table sized for the live voxels.
int map_maybe_grow(limu_map *m, int64_t incoming) {
    if ((m->used_upper + incoming) * 2 <= m->capacity) return LIMU_OK;
    unsigned long long cnt[4];
    LIMU_TRY(read_counters(m, cnt));
    const int64_t live = (int64_t)cnt[0], used = (int64_t)cnt[3];
    m->used_upper = used;
    if ((used + incoming) * 2 <= m->capacity) return LIMU_OK;
    LIMU_TRY(map_wait_readers(m));   // (the old table is freed below)
    limu_ctx *c = m->ctx;
    const int64_t newC = std::max<int64_t>(next_pow2((live + incoming) * 4), 1024);
    limu_map old = *m;  // shallow: keeps the old buffers alive
    m->blk = DevBuf(); m->pend = DevBuf(); m->live = DevBuf(); m->counters = DevBuf();
    int st = map_alloc(m, newC);
    if (st != LIMU_OK) {
        m->blk.release(); m->pend.release(); m->live.release(); m->counters.release();
        m->blk = old.blk; m->pend = old.pend; m->live = old.live; m->counters = old.counters;
        m->capacity = old.capacity;
        set_error("voxel map cannot grow to %lld slots", (long long)newC);
        return LIMU_ERR_MAP_FULL;
    }
    k_rehash<<<div_up(old.capacity, 256), 256, 0, c->stream>>>(old.view(), old.capacity, m->view(), m->counters.as<unsigned long long>());
    LIMU_LAUNCHED();
    unsigned long long *h = static_cast<unsigned long long *>(c->h_pinned);
    h[0] = (unsigned long long)live; h[1] = 0; h[2] = 0;   // [3] (used slots = length of the live list) was counted by k_rehash itself
    LIMU_CUDA_TRY(cudaMemcpyAsync(m->counters.p, h, 3 * sizeof(unsigned long long), cudaMemcpyHostToDevice, c->stream));
    LIMU_CUDA_TRY(cudaStreamSynchronize(c->stream));
    old.blk.release(); old.pend.release(); old.live.release(); old.counters.release();
    old.pslot = DevBuf(); old.world = DevBuf();  // still owned by *m
    m->used_upper = live;
    return LIMU_OK;
}

int map_insert_device(limu_map *m, const double *xyz_dev, int64_t n, const int *n_dev) {
    if (n <= 0) return LIMU_OK;
    limu_ctx *c = m->ctx;
    LIMU_TRY(map_wait_readers(m));
    LIMU_TRY(map_maybe_grow(m, n));
    LIMU_TRY(m->pslot.reserve((size_t)n * 4, c->stream));
    const MapView v = m->view();
    const int blocks = div_up(n, 256);
    k_insert_claim<<<blocks, 256, 0, c->stream>>>(v, xyz_dev, n, n_dev, m->birth_base, m->pslot.as<unsigned int>(),
                                                  m->counters.as<unsigned long long>(), c->d_status);
    LIMU_LAUNCHED();
    k_insert_place<<<blocks, 256, 0, c->stream>>>(v, xyz_dev, n, n_dev, m->pslot.as<unsigned int>());
    LIMU_LAUNCHED();
    m->birth_base += (uint64_t)n;
    m->used_upper += n;
    ++m->mutations;
    return LIMU_OK;
}

int map_remove_far_device(limu_map *m, const double *origin_dev3) {
    limu_ctx *c = m->ctx;
    LIMU_TRY(map_wait_readers(m));
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(div_up(std::max<int64_t>(m->used_upper, 1), 256), (int64_t)c->sm_count * 8));
    k_remove_far<<<blocks, 256, 0, c->stream>>>(m->view(), origin_dev3, m->max_distance, m->counters.as<unsigned long long>());
    LIMU_LAUNCHED();
    ++m->mutations;
    return LIMU_OK;
}

int stage_in(limu_ctx *c, DevBuf &buf, const void *host, size_t bytes) {
    LIMU_TRY(buf.reserve(bytes ? bytes : 8, c->stream));
    if (bytes) LIMU_CUDA_TRY(cudaMemcpyAsync(buf.p, host, bytes, cudaMemcpyHostToDevice, c->stream));
    return LIMU_OK;
}

// Copy a few doubles to the context's small device area at element offset `off`; returns device pointer.
int stage_small(limu_ctx *c, const double *host, int count, int off, double **dev) {
    double *h = static_cast<double *>(c->h_pinned) + 64 + off;
    for (int i = 0; i < count; ++i) h[i] = host[i];
    double *d = c->d_small.as<double>() + 64 + off;
    LIMU_CUDA_TRY(cudaMemcpyAsync(d, h, sizeof(double) * count, cudaMemcpyHostToDevice, c->stream));
    *dev = d;
    return LIMU_OK;
}

}  // namespace limu

using namespace limu;

extern "C" {

int limu_map_create(limu_ctx *c, double vox_size, double max_distance, int cap, int64_t capacity_voxels, limu_map **out) {
    LIMU_TRY(bind(c));
    LIMU_REQUIRE(out, "limu_map_create: out is null");
    LIMU_REQUIRE(vox_size > 0 && cap >= 1 && cap <= 4096, "limu_map_create: vox_size must be > 0 and 1 <= max_points_per_voxel <= 4096");
    limu_map *m = new limu_map;
    m->ctx = c; m->vox_size = vox_size; m->max_distance = max_distance; m->cap = cap;
    const int64_t C = next_pow2(std::max<int64_t>(capacity_voxels, 512) * 2);
    int st = map_alloc(m, C);
    if (st != LIMU_OK) { limu_map_destroy(m); return st; }
    *out = m;
    return LIMU_OK;
}

void limu_map_destroy(limu_map *m) {
    if (!m) return;
    cudaSetDevice(m->ctx->device);
    cudaStreamSynchronize(m->ctx->stream);
    m->blk.release(); m->pend.release(); m->live.release(); m->counters.release();
    m->pslot.release(); m->world.release();
    delete m;
}

int limu_map_clear(limu_map *m) {
    LIMU_REQUIRE(m, "limu_map_clear: null map");
    LIMU_TRY(bind(m->ctx));
    limu_ctx *c = m->ctx;
    LIMU_TRY(map_wait_readers(m));
    const int blocks = std::min<int64_t>(div_up(m->capacity, 256), (int64_t)c->sm_count * 32);
    k_map_clear<<<blocks, 256, 0, c->stream>>>(m->view(), m->capacity);
    LIMU_LAUNCHED();
    LIMU_CUDA_TRY(cudaMemsetAsync(m->pend.p, 0xFF, (size_t)m->capacity * m->cap * 4, c->stream));
    LIMU_CUDA_TRY(cudaMemsetAsync(m->counters.p, 0, 8 * sizeof(unsigned long long), c->stream));
    m->used_upper = 0;
    ++m->mutations;
    return LIMU_OK;
}

int limu_map_size(limu_map *m, int64_t *n_voxels, int64_t *n_points) {
    LIMU_REQUIRE(m, "limu_map_size: null map");
    LIMU_TRY(bind(m->ctx));
    limu_ctx *c = m->ctx;
    unsigned long long *d = m->counters.as<unsigned long long>() + 4;
    LIMU_CUDA_TRY(cudaMemsetAsync(d, 0, 2 * sizeof(unsigned long long), c->stream));
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(div_up(std::max<int64_t>(m->used_upper, 1), 256), (int64_t)c->sm_count * 8));
    k_sum_counts<<<blocks, 256, 0, c->stream>>>(m->view(), m->counters.as<unsigned long long>(), d);
    LIMU_LAUNCHED();
    unsigned long long *h = static_cast<unsigned long long *>(c->h_pinned);
    LIMU_CUDA_TRY(cudaMemcpyAsync(h, d, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    LIMU_CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (n_voxels) *n_voxels = (int64_t)h[0];
    if (n_points) *n_points = (int64_t)h[1];
    return LIMU_OK;
}

int limu_map_empty(limu_map *m, int *out) {
    int64_t nv = 0;
    LIMU_TRY(limu_map_size(m, &nv, nullptr));
    if (out) *out = nv == 0;
    return LIMU_OK;
}

int limu_map_insert_dev(limu_map *m, const double *xyz_dev, int64_t n) {
    LIMU_REQUIRE(m && (xyz_dev || n == 0) && n >= 0, "limu_map_insert_dev: bad arguments");
    LIMU_TRY(bind(m->ctx));
    LIMU_TRY(map_insert_device(m, xyz_dev, n, nullptr));
    return check_status(m->ctx);
}

int limu_map_insert(limu_map *m, const double *xyz, int64_t n) {
    LIMU_REQUIRE(m && (xyz || n == 0) && n >= 0, "limu_map_insert: bad arguments");
    LIMU_TRY(bind(m->ctx));
    if (n == 0) return LIMU_OK;
    LIMU_TRY(stage_in(m->ctx, m->world, xyz, (size_t)n * 24));
    LIMU_TRY(map_insert_device(m, m->world.as<double>(), n, nullptr));
    return check_status(m->ctx);
}

int limu_map_remove_far(limu_map *m, const double origin[3]) {
    LIMU_REQUIRE(m && origin, "limu_map_remove_far: bad arguments");
    LIMU_TRY(bind(m->ctx));
    double *d;
    LIMU_TRY(stage_small(m->ctx, origin, 3, 0, &d));
    LIMU_TRY(map_remove_far_device(m, d));
    return check_status(m->ctx);
}

int limu_map_update_origin(limu_map *m, const double *xyz, int64_t n, const double origin[3]) {
    LIMU_REQUIRE(m && origin && (xyz || n == 0) && n >= 0, "limu_map_update_origin: bad arguments");
    LIMU_TRY(bind(m->ctx));
    if (n > 0) {
        LIMU_TRY(stage_in(m->ctx, m->world, xyz, (size_t)n * 24));
        LIMU_TRY(map_insert_device(m, m->world.as<double>(), n, nullptr));
    }
    double *d;
    LIMU_TRY(stage_small(m->ctx, origin, 3, 0, &d));
    LIMU_TRY(map_remove_far_device(m, d));
    return check_status(m->ctx);
}

int limu_map_update(limu_map *m, const double *xyz, int64_t n, const double pose[7]) {
    LIMU_REQUIRE(m && pose && (xyz || n == 0) && n >= 0, "limu_map_update: bad arguments");
    LIMU_TRY(bind(m->ctx));
    limu_ctx *c = m->ctx;
    double *dpose;
    LIMU_TRY(stage_small(c, pose, 7, 0, &dpose));
    if (n > 0) {   // transform_points prints and returns on an empty vector (calculation_helpers.cpp:123-127)
        LIMU_TRY(stage_in(c, m->world, xyz, (size_t)n * 24));
        LIMU_TRY(transform_device(c, dpose, m->world.as<double>(), m->world.as<double>(), n, nullptr));
        LIMU_TRY(map_insert_device(m, m->world.as<double>(), n, nullptr));
    } else {
        printf("[INFO] utils::transform_points the points vector is empty\n");
    }
    LIMU_TRY(map_remove_far_device(m, dpose + 4));
    return check_status(c);
}

int limu_map_closest(limu_map *m, const double *xyz, int64_t n, double *out_xyz, int32_t *out_key, int32_t *out_rank) {
    return limu_map_closest_ex(m, xyz, n, LIMU_ICP_REFERENCE, out_xyz, out_key, out_rank);
}

int limu_map_closest_ex(limu_map *m, const double *xyz, int64_t n, int32_t icp_mode, double *out_xyz, int32_t *out_key, int32_t *out_rank) {
    LIMU_REQUIRE(m && (xyz || n == 0) && n >= 0 && (out_xyz || n == 0) && (icp_mode & ~LIMU_ICP_NN27) == 0, "limu_map_closest: bad arguments");
    LIMU_TRY(bind(m->ctx));
    if (n == 0) return LIMU_OK;
    limu_ctx *c = m->ctx;
    LIMU_TRY(stage_in(c, c->in0, xyz, (size_t)n * 24));
    LIMU_TRY(c->out0.reserve((size_t)n * 24, c->stream));
    LIMU_TRY(c->out1.reserve((size_t)n * 12, c->stream));
    LIMU_TRY(c->out2.reserve((size_t)n * 4, c->stream));
    k_closest<<<div_up(n, 256), 256, 0, c->stream>>>(m->view(), c->in0.as<double>(), n, nullptr, 0.0, c->out0.as<double>(),
                                                     out_key ? c->out1.as<int>() : nullptr, out_rank ? c->out2.as<int>() : nullptr, nullptr,
                                                     (icp_mode & LIMU_ICP_NN27) ? 1 : 0);
    LIMU_LAUNCHED();
    LIMU_CUDA_TRY(cudaMemcpyAsync(out_xyz, c->out0.p, (size_t)n * 24, cudaMemcpyDeviceToHost, c->stream));
    if (out_key) LIMU_CUDA_TRY(cudaMemcpyAsync(out_key, c->out1.p, (size_t)n * 12, cudaMemcpyDeviceToHost, c->stream));
    if (out_rank) LIMU_CUDA_TRY(cudaMemcpyAsync(out_rank, c->out2.p, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
    return check_status(c);
}

int limu_map_correspondences(limu_map *m, const double *xyz, int64_t n, double max_correspondance, double *src, double *tgt,
                             int64_t *out_idx, int64_t *n_out) {
    LIMU_REQUIRE(m && (xyz || n == 0) && n >= 0 && n_out, "limu_map_correspondences: bad arguments");
    LIMU_TRY(bind(m->ctx));
    *n_out = 0;
    if (n == 0) return LIMU_OK;
    limu_ctx *c = m->ctx;
    LIMU_TRY(stage_in(c, c->in0, xyz, (size_t)n * 24));
    LIMU_TRY(c->out0.reserve((size_t)n * 24, c->stream));   // nearest per query
    LIMU_TRY(c->tmp0.reserve((size_t)n, c->stream));        // flags
    LIMU_TRY(c->tmp1.reserve((size_t)div_up(n, COMPACT_BLOCK) * 4 + 16, c->stream));
    LIMU_TRY(c->tmp2.reserve((size_t)n * 4 + 16, c->stream));   // survivor indices
    LIMU_TRY(c->out1.reserve((size_t)n * 24, c->stream));   // compact src
    LIMU_TRY(c->out2.reserve((size_t)n * 24, c->stream));   // compact tgt
    LIMU_TRY(c->tmp3.reserve((size_t)n * 8, c->stream));    // idx64
    int *count_dev = reinterpret_cast<int *>(c->d_small.as<double>());
    const double max_sq = max_correspondance * max_correspondance;   // voxel_hash_map.cpp:112
    k_closest<<<div_up(n, 256), 256, 0, c->stream>>>(m->view(), c->in0.as<double>(), n, nullptr, max_sq, c->out0.as<double>(), nullptr, nullptr,
                                                     c->tmp0.as<unsigned char>(), 0);
    LIMU_LAUNCHED();
    LIMU_TRY(compact_flags(c, c->tmp0.as<unsigned char>(), n, nullptr, c->tmp1.as<int>(), c->tmp2.as<int>(), count_dev));
    const int gb = std::min<int64_t>(div_up(n, 256), (int64_t)c->sm_count * 8);
    k_gather_points<<<gb, 256, 0, c->stream>>>(c->in0.as<double>(), c->tmp2.as<int>(), count_dev, c->out1.as<double>(), c->tmp3.as<int64_t>());
    LIMU_LAUNCHED();
    k_gather_points<<<gb, 256, 0, c->stream>>>(c->out0.as<double>(), c->tmp2.as<int>(), count_dev, c->out2.as<double>(), nullptr);
    LIMU_LAUNCHED();
    int *h = static_cast<int *>(c->h_pinned);
    LIMU_CUDA_TRY(cudaMemcpyAsync(h, count_dev, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    LIMU_CUDA_TRY(cudaStreamSynchronize(c->stream));
    const int64_t k = h[0];
    *n_out = k;
    if (k > 0) {
        if (src) LIMU_CUDA_TRY(cudaMemcpyAsync(src, c->out1.p, (size_t)k * 24, cudaMemcpyDeviceToHost, c->stream));
        if (tgt) LIMU_CUDA_TRY(cudaMemcpyAsync(tgt, c->out2.p, (size_t)k * 24, cudaMemcpyDeviceToHost, c->stream));
        if (out_idx) LIMU_CUDA_TRY(cudaMemcpyAsync(out_idx, c->tmp3.p, (size_t)k * 8, cudaMemcpyDeviceToHost, c->stream));
    }
    return check_status(c);
}

int limu_map_dump(limu_map *m, int32_t *keys, int32_t *counts, double *pts, int64_t max_voxels, int64_t max_points, int64_t *n_voxels,
                  int64_t *n_points) {
    LIMU_REQUIRE(m, "limu_map_dump: null map");
    LIMU_TRY(bind(m->ctx));
    limu_ctx *c = m->ctx;
    int64_t nv = 0, np = 0;
    LIMU_TRY(limu_map_size(m, &nv, &np));
    if (n_voxels) *n_voxels = nv;
    if (n_points) *n_points = np;
    if ((!keys && !counts && !pts) || nv == 0) return LIMU_OK;
    // 1. collect (birth, slot), sort by birth on the host -> creation order
    LIMU_TRY(c->tmp0.reserve((size_t)nv * 16 + 16, c->stream));
    unsigned long long *cnt = m->counters.as<unsigned long long>() + 6;
    LIMU_CUDA_TRY(cudaMemsetAsync(cnt, 0, 8, c->stream));
    k_collect_live<<<(int)std::max<int64_t>(1, std::min<int64_t>(div_up(std::max<int64_t>(m->used_upper, 1), 256), (int64_t)c->sm_count * 8)), 256, 0, c->stream>>>(
        m->view(), m->counters.as<unsigned long long>(), c->tmp0.as<unsigned long long>(), cnt);
    LIMU_LAUNCHED();
    std::vector<unsigned long long> pairs((size_t)nv * 2);
    LIMU_CUDA_TRY(cudaMemcpyAsync(pairs.data(), c->tmp0.p, (size_t)nv * 16, cudaMemcpyDeviceToHost, c->stream));
    LIMU_CUDA_TRY(cudaStreamSynchronize(c->stream));
    std::vector<std::pair<unsigned long long, unsigned int>> ord((size_t)nv);
    for (int64_t j = 0; j < nv; ++j) ord[j] = {pairs[2 * j], (unsigned int)pairs[2 * j + 1]};
    std::sort(ord.begin(), ord.end());
    std::vector<unsigned int> order((size_t)nv);
    for (int64_t j = 0; j < nv; ++j) order[j] = ord[j].second;
    // 2. gather voxel contents in that order
    LIMU_TRY(c->tmp1.reserve((size_t)nv * 4, c->stream));
    LIMU_TRY(c->out0.reserve((size_t)nv * 12, c->stream));
    LIMU_TRY(c->out1.reserve((size_t)nv * 4, c->stream));
    LIMU_TRY(c->out2.reserve((size_t)nv * m->cap * 24, c->stream));
    LIMU_CUDA_TRY(cudaMemcpyAsync(c->tmp1.p, order.data(), (size_t)nv * 4, cudaMemcpyHostToDevice, c->stream));
    k_gather_voxels<<<div_up(nv, 128), 128, 0, c->stream>>>(m->view(), c->tmp1.as<unsigned int>(), nv, c->out0.as<int>(), c->out1.as<int>(),
                                                          c->out2.as<double>());
    LIMU_LAUNCHED();
    std::vector<int> hk((size_t)nv * 3), hc((size_t)nv);
    std::vector<double> hp((size_t)nv * m->cap * 3);
    LIMU_CUDA_TRY(cudaMemcpyAsync(hk.data(), c->out0.p, hk.size() * 4, cudaMemcpyDeviceToHost, c->stream));
    LIMU_CUDA_TRY(cudaMemcpyAsync(hc.data(), c->out1.p, hc.size() * 4, cudaMemcpyDeviceToHost, c->stream));
    LIMU_CUDA_TRY(cudaMemcpyAsync(hp.data(), c->out2.p, hp.size() * 8, cudaMemcpyDeviceToHost, c->stream));
    LIMU_CUDA_TRY(cudaStreamSynchronize(c->stream));
    int64_t w = 0;
    for (int64_t j = 0; j < nv; ++j) {
        if (j < max_voxels) {
            if (keys) { keys[3 * j] = hk[3 * j]; keys[3 * j + 1] = hk[3 * j + 1]; keys[3 * j + 2] = hk[3 * j + 2]; }
            if (counts) counts[j] = hc[j];
        }
        for (int r = 0; r < hc[j]; ++r, ++w)
            if (pts && w < max_points) memcpy(pts + 3 * w, hp.data() + ((size_t)j * m->cap + r) * 3, 24);
    }
    return LIMU_OK;
}

int limu_map_pointcloud(limu_map *m, double *out_xyz, int64_t max_points, int64_t *n_out) {
    int64_t nv = 0, np = 0;
    LIMU_TRY(limu_map_dump(m, nullptr, nullptr, out_xyz, 0, out_xyz ? max_points : 0, &nv, &np));
    if (n_out) *n_out = np;
    return LIMU_OK;
}

}  // extern "C"
