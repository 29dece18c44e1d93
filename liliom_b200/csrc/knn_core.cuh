// Lane-group-cooperative exact 5-NN over the dense cell grid (shared by the scan-to-map kernel
// and the backend correspondence kernels).  See grid_knn.cu for the design notes.
//
// Candidate ordering is (fp32 squared distance, map index).  Both live in one 64-bit key
//     key = (bits(d) << 32) | index             (d >= +0, so the bit pattern orders like the value)
// so a single unsigned compare implements FLANN's distance order plus our index tie-break with
// no branches and no extra loads.
// Loops are deliberately NOT unrolled: the first version of this kernel was 11.7k SASS
// instructions and spent 47 % of its issue slots in stall_no_inst (instruction-cache misses,
// profiles/r01_knn_v1_ncu.txt).
#pragma once
#include "ctx.cuh"
#include "dev_math.cuh"

namespace lili {

__device__ __forceinline__ int cell_coord(float v, float inv_cell) { return (int)floorf(v * inv_cell); }

constexpr int kBlock = 256;
constexpr int kWarps = kBlock / 32;

__device__ __forceinline__ unsigned block_hash(int bx, int by, int bz) {
    unsigned h = (unsigned)bx * 73856093u ^ (unsigned)by * 19349663u ^ (unsigned)bz * 83492791u;
    h ^= h >> 15; h *= 0x2c1b3c6du; h ^= h >> 12;
    return h;
}
__device__ __forceinline__ int owner_of(float x, float y, float z, int nranks) {
    int bx = (int)floorf(x * 0.0625f), by = (int)floorf(y * 0.0625f), bz = (int)floorf(z * 0.0625f);
    return (int)(block_hash(bx, by, bz) % (unsigned)nranks);
}

typedef unsigned long long u64;

// Sorted (ascending) list of the five best keys.  The low word of a key is the neighbour's index in
// the un-sorted map array (map_download order), which is also where phase B fetches its coordinates
// from, so no separate payload has to ride through the compare-exchanges.
struct Top5 { u64 k0, k1, k2, k3, k4; };

__device__ __forceinline__ void top5_init(Top5& t) { t.k0 = t.k1 = t.k2 = t.k3 = t.k4 = ~0ull; }
__device__ __forceinline__ float top5_dist(u64 k) { return __uint_as_float((unsigned)(k >> 32)); }
__device__ __forceinline__ int top5_index(u64 k) { return k == ~0ull ? -1 : (int)(unsigned)(k & 0xffffffffull); }

__device__ __forceinline__ void cswap(u64& a, u64& b) {   // afterwards a <= b
    const bool s = b < a;
    const u64 lo = s ? b : a, hi = s ? a : b;
    a = lo; b = hi;
}

__device__ __forceinline__ void top5_insert(Top5& t, u64 k) {
    if (k < t.k4) {
        t.k4 = k;
        cswap(t.k3, t.k4);
        cswap(t.k2, t.k3);
        cswap(t.k1, t.k2);
        cswap(t.k0, t.k1);
    }
}

__device__ __forceinline__ u64 make_key(float sx, float sy, float sz, const float4& m) {
    // FLANN L2_Simple<float>: ((dx*dx) + dy*dy) + dz*dz with no contraction (exact-op intrinsics)
    const float dx = fsubx(sx, m.x), dy = fsubx(sy, m.y), dz = fsubx(sz, m.z);
    const float d = faddx(faddx(fmulx(dx, dx), fmulx(dy, dy)), fmulx(dz, dz));
    return ((u64)__float_as_uint(d) << 32) | (u64)(unsigned)__float_as_int(m.w);
}

// LANES (1, 2, 4, 8 or 16) lanes — mask `gmask`, lane-in-group `sub` — search the 3x3x3 cell block around
// (sx,sy,sz).  The three x-adjacent cells of a (y,z) row are ONE contiguous run of the cell-sorted
// map, so the block is 9 runs; lane `sub` walks rows sub, sub+LANES, ...  Each lane pulls its run in
// batches of 8 independent 16-byte loads (memory-level parallelism without occupancy), ranks them into
// a private sorted top-5, and the group merges with 5 REDUX.MIN rounds.  With LANES == 1 a thread owns a
// whole query: consecutive queries are spatial neighbours (VoxelGrid output order), so the warp's
// loads hit the same cells in L1.  On return every lane of the group holds the merged top-5.
// `cand` accumulates the number of map points this lane examined.
template <int LANES, int BATCH = 8>
__device__ __forceinline__ void group_knn5(float sx, float sy, float sz, const float4* __restrict__ map,
                                           const int* __restrict__ cell_start, const GridDesc& g, int sub, unsigned gmask,
                                           Top5& top, unsigned long long& cand, long long* dbg = nullptr) {
    const int cx = cell_coord(sx, g.inv_cell) - g.org[0];
    const int cy = cell_coord(sy, g.inv_cell) - g.org[1];
    const int cz = cell_coord(sz, g.inv_cell) - g.org[2];
    const int x0 = max(cx - 1, 0), x1 = min(cx + 1, g.dim[0] - 1);
    auto row_range = [&](int row, int& b, int& e) {
        b = 0; e = 0;
        if (row < 9 && x0 <= x1) {
            const int y = cy + (row % 3) - 1, z = cz + (row / 3) - 1;
            if (y >= 0 && y < g.dim[1] && z >= 0 && z < g.dim[2]) {
                const int base = (z * g.dim[1] + y) * g.dim[0];
                b = __ldg(cell_start + base + x0);
                e = __ldg(cell_start + base + x1 + 1);
            }
        }
    };
    int b, e;
    row_range(sub, b, e);
    if (dbg) dbg[9] = clock64() + (long long)(b & 0);
#pragma unroll 1
    for (int row = sub; row < 9; row += LANES) {
        int nb, ne;
        row_range(row + LANES, nb, ne);          // next run's bounds are in flight while this run is ranked
        cand += (unsigned long long)(e - b);
#pragma unroll 1
        for (int p0 = b; p0 < e; p0 += BATCH) {
            float4 c[BATCH];
#pragma unroll
            for (int i = 0; i < BATCH; ++i) if (p0 + i < e) c[i] = __ldg(map + p0 + i);
#pragma unroll
            for (int i = 0; i < BATCH; ++i) if (p0 + i < e) top5_insert(top, make_key(sx, sy, sz, c[i]));
        }
        b = nb; e = ne;
    }
    if (dbg) dbg[10] = clock64() + (long long)(top.k0 & 0);
    if (LANES > 1) {
        // merge: 5 rounds of "group-wide minimum of the list heads, winner pops".  The minimum is an xor
        // butterfly of 64-bit keys (REDUX.MIN on a partial lane mask measured ~350 cycles per call on
        // B200, the shuffle butterfly ~35 cycles per step).  Keys are unique — a map point lives in
        // exactly one lane's list — so exactly one lane pops per round.
        Top5 res;
#define LILI_MERGE_ROUND(KJ)                                                              \
        {                                                                                 \
            u64 mn = top.k0;                                                              \
            _Pragma("unroll")                                                             \
            for (int o = 1; o < LANES; o <<= 1) { const u64 other = __shfl_xor_sync(gmask, mn, o); mn = other < mn ? other : mn; } \
            KJ = mn;                                                                      \
            if (top.k0 == mn && mn != ~0ull) { top.k0 = top.k1; top.k1 = top.k2; top.k2 = top.k3; top.k3 = top.k4; top.k4 = ~0ull; } \
        }
        LILI_MERGE_ROUND(res.k0)
        LILI_MERGE_ROUND(res.k1)
        LILI_MERGE_ROUND(res.k2)
        LILI_MERGE_ROUND(res.k3)
        LILI_MERGE_ROUND(res.k4)
#undef LILI_MERGE_ROUND
        top = res;
    }
}


// ---- flat variant for the latency-bound small-scan shape (LANES == 16: two queries per warp) -------------------
// group_knn5 gives lane r the whole run r, so a query costs max-run-length ranking steps and lanes 9..15 idle
// (measured on B200, 1.7k queries: 5.3k of a 13k-cycle pass in "candidates + rank").  Here the 9 run bounds are
// shared through shuffles and the concatenated candidate list is dealt round-robin: lane `sub` ranks candidates
// sub, sub+LANES, ... — ceil(T/LANES) steps instead of the longest run, one batch of independent 16-byte loads per
// lane for T <= 8*LANES, consecutive lanes reading consecutive map points.  The candidate SET is unchanged, so the
// merged top-5 is bit-identical to group_knn5's.
// `cc` (optional, per-thread slots in shared memory): a lane's batch is kept across the GN iterations of the
// persistent kernel; while the transformed query stays in the same cell the bounds + candidate loads (two dependent
// L2 round trips) are skipped.  cc_tag holds the cell the cache was filled for (tag.w < 0: empty / not cacheable).
constexpr int kFlatBatch = 8;

template <int LANES>
__device__ __forceinline__ void group_knn5_flat(float sx, float sy, float sz, const float4* __restrict__ map,
                                                const int* __restrict__ cell_start, const GridDesc& g, int sub, unsigned gmask,
                                                Top5& top, unsigned long long& cand, float4* cc, int cc_stride, int4* cc_tag,
                                                float4* nb_out = nullptr, int* nb_flag = nullptr) {
    static_assert(LANES >= 16, "one lane per (y,z) row of the 3x3x3 block");
    const int cx = cell_coord(sx, g.inv_cell) - g.org[0];
    const int cy = cell_coord(sy, g.inv_cell) - g.org[1];
    const int cz = cell_coord(sz, g.inv_cell) - g.org[2];
    float4 c[kFlatBatch];
    int nmine = 0, T = 0;
    int rb[9], pre[10];
    const bool reuse = cc_tag && cc_tag->w >= 0 && cc_tag->x == cx && cc_tag->y == cy && cc_tag->z == cz;   // uniform in the group
    if (reuse) {
        nmine = cc_tag->w;
#pragma unroll
        for (int i = 0; i < kFlatBatch; ++i) if (i < nmine) c[i] = cc[i * cc_stride];
    } else {
        const int x0 = max(cx - 1, 0), x1 = min(cx + 1, g.dim[0] - 1);
        int b = 0, len = 0;
        if (sub < 9 && x0 <= x1) {
            const int y = cy + (sub % 3) - 1, z = cz + (sub / 3) - 1;
            if (y >= 0 && y < g.dim[1] && z >= 0 && z < g.dim[2]) {
                const int base = (z * g.dim[1] + y) * g.dim[0];
                b = __ldg(cell_start + base + x0);
                len = __ldg(cell_start + base + x1 + 1) - b;
            }
        }
        pre[0] = 0;
#pragma unroll
        for (int r = 0; r < 9; ++r) {
            rb[r] = __shfl_sync(gmask, b, r, LANES);
            pre[r + 1] = pre[r] + __shfl_sync(gmask, len, r, LANES);
        }
        T = pre[9];
    }
#pragma unroll 1
    for (int j0 = sub;; j0 += LANES * kFlatBatch) {
        if (!reuse) {
            nmine = 0;
#pragma unroll
            for (int i = 0; i < kFlatBatch; ++i) {
                const int j = j0 + i * LANES;
                if (j < T) {
                    int pos = rb[0] + j;
#pragma unroll
                    for (int r = 1; r < 9; ++r) if (j >= pre[r]) pos = rb[r] + (j - pre[r]);   // last row whose prefix <= j
                    c[i] = __ldg(map + pos);
                    nmine = i + 1;
                }
            }
            if (cc_tag && j0 == sub) {      // first batch: cacheable when it is also the only one
                if (T <= LANES * kFlatBatch) {
#pragma unroll
                    for (int i = 0; i < kFlatBatch; ++i) if (i < nmine) cc[i * cc_stride] = c[i];
                    *cc_tag = make_int4(cx, cy, cz, nmine);
                } else cc_tag->w = -1;
            }
        }
#pragma unroll
        for (int i = 0; i < kFlatBatch; ++i) if (i < nmine) top5_insert(top, make_key(sx, sy, sz, c[i]));
        cand += (unsigned long long)nmine;
        if (reuse || j0 - sub + LANES * kFlatBatch >= T) break;     // group-uniform: T and reuse are
    }
    // merge (same as group_knn5)
    Top5 res;
#define LILI_MERGE_ROUND(KJ)                                                              \
    {                                                                                     \
        u64 mn = top.k0;                                                                  \
        _Pragma("unroll")                                                                 \
        for (int o = 1; o < LANES; o <<= 1) { const u64 other = __shfl_xor_sync(gmask, mn, o); mn = other < mn ? other : mn; } \
        KJ = mn;                                                                          \
        if (top.k0 == mn && mn != ~0ull) { top.k0 = top.k1; top.k1 = top.k2; top.k2 = top.k3; top.k3 = top.k4; top.k4 = ~0ull; } \
    }
    LILI_MERGE_ROUND(res.k0)
    LILI_MERGE_ROUND(res.k1)
    LILI_MERGE_ROUND(res.k2)
    LILI_MERGE_ROUND(res.k3)
    LILI_MERGE_ROUND(res.k4)
#undef LILI_MERGE_ROUND
    top = res;
    // Hand the winners' coordinates to the plane fit through shared memory: the lane that loaded a winning candidate
    // still holds it in registers (single-batch case), so the fit does not have to fetch the 5 neighbours from L2 again
    // (the cell-sorted copy and map_download order hold the same coordinates).  nb_flag = 1 tells phase B to use them.
    if (nb_out) {
        const bool single = reuse || T <= LANES * kFlatBatch;      // group-uniform
        if (single) {
#pragma unroll
            for (int i = 0; i < kFlatBatch; ++i) {
                if (i < nmine) {
                    const u64 k = make_key(sx, sy, sz, c[i]);
                    const int j = k == res.k0 ? 0 : k == res.k1 ? 1 : k == res.k2 ? 2 : k == res.k3 ? 3 : k == res.k4 ? 4 : -1;
                    if (j >= 0) nb_out[j] = c[i];
                }
            }
        }
        if (sub == 0) *nb_flag = single ? 1 : 0;
    }
}

}  // namespace lili
