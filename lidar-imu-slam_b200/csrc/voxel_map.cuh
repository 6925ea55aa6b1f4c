// voxel_map.cuh -- device view of the GPU-resident voxel hash map and the lookups kernels inline.
//
// Replaces lidar::VoxelHashMap + lidar::VoxelBlock (L/include/limu/sensors/lidar/helpers/
// voxel_hash_map.hpp:14-48, voxel_block.hpp:13-56): tsl::robin_map<Voxel, VoxelBlock> whose blocks hold
// one heap node per point becomes ONE open-addressing table in HBM:
//
//   slots[C]        16 B  {u64 packed (i,j,k) key, u64 meta = birth << 13 | count}   C = power of two, load <= 0.5
//                          birth = creation sequence of the voxel (= the reference container's address order,
//                          used only by the 27-cell fallback tie-break, voxel_hash_map.cpp:81-101)
//   pts[C*stride]   24 B  per point, slot-indexed, structure-of-arrays INSIDE a voxel: voxel s owns three rows
//                          x[capp] y[capp] z[capp] at pts + s*3*capp (capp = cap rounded up to 4 doubles = one 32 B
//                          sector), so eight lanes reading ranks r..r+7 of one row fetch 64 contiguous bytes
//   pend[C*cap]      4 B  per point slot: insert scratch (sorted pending input indices), all-ones at rest
//
// One 16-byte load answers "is this my voxel, how many points does it hold, and how old is it"; the points
// of a voxel are contiguous (capp*24 B), so a query touches 1 + 3*ceil(8*count/32) sectors.
#pragma once
#include "common.cuh"

namespace limu {

constexpr unsigned long long KEY_EMPTY = 0xFFFFFFFFFFFFFFFFull;
constexpr unsigned long long KEY_TOMB = 0xFFFFFFFFFFFFFFFEull;
constexpr unsigned long long META_NONE = 0xFFFFFFFFFFFFFFFFull;
constexpr int META_COUNT_BITS = 13;                       // max_points_per_voxel <= 4096 < 2^13
constexpr unsigned long long META_COUNT_MASK = (1ull << META_COUNT_BITS) - 1ull;
constexpr unsigned int PEND_NONE = 0xFFFFFFFFu;
constexpr int KEY_BIAS = 1 << 20;  // voxel indices in (-2^20, 2^20) pack into 3 x 21 bits

struct __align__(16) Slot {
    unsigned long long key;
    unsigned long long meta;   // birth << 13 | count
};
__host__ __device__ __forceinline__ int meta_count(unsigned long long meta) { return (int)(meta & META_COUNT_MASK); }
__host__ __device__ __forceinline__ unsigned long long meta_birth(unsigned long long meta) { return meta >> META_COUNT_BITS; }

struct MapView {
    Slot *slots;
    double *pts;
    unsigned int *pend;
    unsigned int mask;   // C - 1
    int shift;           // 64 - log2(C)
    int cap;
    int capp;            // cap rounded up to a multiple of 4: row length of the per-voxel SoA block
    int stride;          // doubles per voxel block: 3*capp rounded up to 16 (= one 128 B L2 line), so a block never straddles an extra line
    double vox;
};
__host__ __device__ __forceinline__ int cap_padded(int cap) { return (cap + 3) & ~3; }
__host__ __device__ __forceinline__ int block_stride(int cap) { return (3 * cap_padded(cap) + 15) & ~15; }
// row pointers of voxel `slot`: x = base, y = base + capp, z = base + 2*capp
__device__ __forceinline__ double *voxel_rows(const MapView &m, unsigned int slot) { return m.pts + (size_t)slot * (size_t)m.stride; }

// utils::get_vox_index, calculation_helpers.cpp:142-147: IEEE double division, truncation toward zero.
__device__ __forceinline__ int vox_index(double p, double v) { return __double2int_rz(p / v); }

__device__ __forceinline__ bool key_in_range(int x, int y, int z) {
    return (unsigned)(x + KEY_BIAS - 1) < (unsigned)(2 * KEY_BIAS - 1) && (unsigned)(y + KEY_BIAS - 1) < (unsigned)(2 * KEY_BIAS - 1) &&
           (unsigned)(z + KEY_BIAS - 1) < (unsigned)(2 * KEY_BIAS - 1);
}
__device__ __forceinline__ unsigned long long pack_key(int x, int y, int z) {
    return ((unsigned long long)(unsigned)(x + KEY_BIAS) << 42) | ((unsigned long long)(unsigned)(y + KEY_BIAS) << 21) |
           (unsigned long long)(unsigned)(z + KEY_BIAS);
}
__host__ __device__ __forceinline__ void unpack_key(unsigned long long k, int &x, int &y, int &z) {
    x = (int)((k >> 42) & 0x1FFFFF) - KEY_BIAS;
    y = (int)((k >> 21) & 0x1FFFFF) - KEY_BIAS;
    z = (int)(k & 0x1FFFFF) - KEY_BIAS;
}
// Fibonacci hashing of the packed key; the top log2(C) bits index the table.
__device__ __forceinline__ unsigned int slot_of(unsigned long long key, int shift) {
    return (unsigned int)((key * 0x9E3779B97F4A7C15ull) >> shift);
}

__device__ __forceinline__ ulonglong2 load_slot(const Slot *s) { return __ldg(reinterpret_cast<const ulonglong2 *>(s)); }

// Read-only lookup. Returns slot index or -1; count of the voxel in *count.
__device__ __forceinline__ int map_find(const MapView &m, unsigned long long key, int *count) {
    unsigned int s = slot_of(key, m.shift);
    for (;;) {
        const ulonglong2 v = load_slot(m.slots + s);
        if (v.x == key) { *count = meta_count(v.y); return (int)s; }
        if (v.x == KEY_EMPTY) return -1;
        s = (s + 1) & m.mask;
    }
}

// 27-neighbourhood offsets by |delta|^2 class.
__device__ constexpr signed char NB_CORNERS[8][3] = {{-1, -1, -1}, {-1, -1, 1}, {-1, 1, -1}, {-1, 1, 1}, {1, -1, -1}, {1, -1, 1}, {1, 1, -1}, {1, 1, 1}};
__device__ constexpr signed char NB_EDGES[12][3] = {{-1, -1, 0}, {-1, 1, 0}, {1, -1, 0}, {1, 1, 0}, {-1, 0, -1}, {-1, 0, 1},
                                                    {1, 0, -1},  {1, 0, 1},  {0, -1, -1}, {0, -1, 1}, {0, 1, -1}, {0, 1, 1}};
// faces are padded to 8 entries (two chunks of four); the padding repeats real faces, which is harmless for a maximum
__device__ constexpr signed char NB_FACES[8][3] = {{-1, 0, 0}, {1, 0, 0}, {0, -1, 0}, {0, 1, 0}, {0, 0, -1}, {0, 0, 1}, {0, 0, -1}, {0, 0, 1}};

struct Nearest {
    double x, y, z;   // matched map point, or (0,0,0) when nothing was found
    int slot;         // table slot of the matched voxel, -1 if none
    int rank;         // position of the match inside its voxel, -1 if none
    int ncand;        // points compared
    int own;          // 1 if the query's own voxel was present
};

// VoxelBlock::get_closest_point, voxel_block.cpp:87-105: linear scan, strict '<', first minimum wins.
// One-thread-per-query form: candidates are fetched four at a time (one 32 B sector per row) and compared in order.
__device__ __forceinline__ void block_closest(const MapView &m, int slot, int count, const V3 &p, Nearest &r) {
    const double *bx = voxel_rows(m, (unsigned int)slot), *by = bx + m.capp, *bz = by + m.capp;
    double best = 1.7976931348623157e308;
    r.rank = -1;
    for (int base = 0; base < count; base += 4) {   // rows are padded to a multiple of 4, so the loads stay in bounds
        double cx[4], cy[4], cz[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { cx[k] = __ldg(bx + base + k); cy[k] = __ldg(by + base + k); cz[k] = __ldg(bz + base + k); }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const double d = sqnorm3(p.x - cx[k], p.y - cy[k], p.z - cz[k]);
            if (base + k < count && d < best) { best = d; r.rank = base + k; r.x = cx[k]; r.y = cy[k]; r.z = cz[k]; }
        }
    }
    r.ncand = count;
}

// Step 1 of the lookup: WHICH voxel answers the query (rules (a)/(b) below). Returns its slot (or -1), its count,
// and whether it is the query's own voxel.
__device__ __forceinline__ int map_locate(const MapView &m, const V3 &p, int *count_out, int *own_out) {
    const int kx = vox_index(p.x, m.vox), ky = vox_index(p.y, m.vox), kz = vox_index(p.z, m.vox);
    *own_out = 0;
    *count_out = 0;
    if (key_in_range(kx, ky, kz)) {
        int count = 0;
        const int s = map_find(m, pack_key(kx, ky, kz), &count);
        if (s >= 0) { *own_out = 1; *count_out = count; return s; }
    }
    // Fallback (b): the winner is the occupied neighbour with the largest (|delta|^2, birth). Visit the
    // three distance classes in decreasing |delta|^2 -- 8 corners (3), 12 edges (2), 6 faces (1) -- and stop
    // after the first class with an occupied cell. Cells are probed four at a time: the four 16-byte home-slot
    // loads are issued before any is consumed, and each answers key, birth and count at once.
    int best_slot = -1;
    unsigned long long best_meta = 0;
#define LIMU_PROBE_CHUNK(TABLE, FIRST)                                                                      \
    {                                                                                                       \
        ulonglong2 got[4];                                                                                  \
        _Pragma("unroll") for (int c = 0; c < 4; ++c) {                                                     \
            const int x = kx + TABLE[FIRST + c][0], y = ky + TABLE[FIRST + c][1], z = kz + TABLE[FIRST + c][2]; \
            got[c] = key_in_range(x, y, z) ? load_slot(m.slots + slot_of(pack_key(x, y, z), m.shift)) : make_ulonglong2(KEY_EMPTY, 0ull); \
        }                                                                                                   \
        _Pragma("unroll") for (int c = 0; c < 4; ++c) {                                                     \
            ulonglong2 v = got[c];                                                                          \
            if (v.x != KEY_EMPTY) {                                                                         \
                const unsigned long long want = pack_key(kx + TABLE[FIRST + c][0], ky + TABLE[FIRST + c][1], kz + TABLE[FIRST + c][2]); \
                unsigned int s = slot_of(want, m.shift);                                                    \
                while (v.x != want && v.x != KEY_EMPTY) { s = (s + 1) & m.mask; v = load_slot(m.slots + s); } \
                if (v.x == want && (best_slot < 0 || v.y > best_meta)) { best_meta = v.y; best_slot = (int)s; } \
            }                                                                                               \
        }                                                                                                   \
    }
    LIMU_PROBE_CHUNK(NB_CORNERS, 0)
    LIMU_PROBE_CHUNK(NB_CORNERS, 4)
    if (best_slot < 0) {
        LIMU_PROBE_CHUNK(NB_EDGES, 0)
        LIMU_PROBE_CHUNK(NB_EDGES, 4)
        LIMU_PROBE_CHUNK(NB_EDGES, 8)
    }
    if (best_slot < 0) {
        LIMU_PROBE_CHUNK(NB_FACES, 0)
        LIMU_PROBE_CHUNK(NB_FACES, 4)
    }
#undef LIMU_PROBE_CHUNK
    if (best_slot >= 0) *count_out = meta_count(best_meta);   // births are unique, so comparing meta compared birth
    return best_slot;
}

// VoxelHashMap::get_closest_neighbour, voxel_hash_map.cpp:64-102:
//  (a) own voxel present            -> closest point inside it only (:71-73)
//  (b) else the top of a max-heap on (|delta index|^2, block address) over the occupied cells of the
//      27-neighbourhood (:76-96,101) = the FARTHEST occupied cell, ties to the later-created voxel
//  (c) nothing                      -> (0,0,0) (:98-99)
__device__ __forceinline__ Nearest map_closest(const MapView &m, const V3 &p) {
    Nearest r;
    r.x = r.y = r.z = 0.0; r.rank = -1; r.ncand = 0;
    int count;
    r.slot = map_locate(m, p, &count, &r.own);
    if (r.slot >= 0) {
        block_closest(m, r.slot, count, p, r);
        if (r.rank < 0) { r.x = r.y = r.z = 0.0; }
    }
    return r;
}

}  // namespace limu

// Host-side object behind the C handle.
struct limu_map {
    limu_ctx *ctx = nullptr;
    double vox_size = 1.0, max_distance = 100.0;
    int cap = 10;
    int64_t capacity = 0;          // C (slots), power of two
    limu::DevBuf slots, pts, pend;
    limu::DevBuf counters;         // device: [0] n_live voxels, [1] n_tomb, [2] n_points, [3] n_used (live + tomb)
    uint64_t birth_base = 0;       // creation sequence offset of the next insert batch
    int64_t used_upper = 0;        // host upper bound on live + tomb slots
    limu::DevBuf pslot;            // per-point slot scratch of the current insert batch
    limu::DevBuf world;            // transformed copy for update(points, pose)
    limu::MapView view() const;
};

namespace limu {
int map_alloc(limu_map *m, int64_t capacity_slots);
int map_insert_device(limu_map *m, const double *xyz_dev, int64_t n, const int *n_dev /* optional device count */);
int map_remove_far_device(limu_map *m, const double *origin_dev3);
int map_maybe_grow(limu_map *m, int64_t incoming);
}  // namespace limu
