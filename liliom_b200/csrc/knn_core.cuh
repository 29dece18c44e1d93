// Lane-group-cooperative exact 5-NN over the dense cell grid (shared by the scan-to-map kernel
// and the backend correspondence kernels).  See grid_knn.cu for the design notes.
//
// Candidate ordering is (fp32 squared distance, original map index).  Both live in one 64-bit key
//     key = (bits(d) << 32) | orig_index        (d >= +0, so the bit pattern orders like the value)
// so a single unsigned compare implements FLANN's distance order plus our index tie-break with
// no branches and no extra loads; the position in the cell-sorted array travels as a payload.
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
    int bx = (int)floorf(x * 0.125f), by = (int)floorf(y * 0.125f), bz = (int)floorf(z * 0.125f);
    return (int)(block_hash(bx, by, bz) % (unsigned)nranks);
}

typedef unsigned long long u64;

struct Top5 {
    u64 k0, k1, k2, k3, k4;   // ascending
    int p0, p1, p2, p3, p4;   // positions in the cell-sorted map (-1 = empty)
};

__device__ __forceinline__ void top5_init(Top5& t) {
    t.k0 = t.k1 = t.k2 = t.k3 = t.k4 = ~0ull;
    t.p0 = t.p1 = t.p2 = t.p3 = t.p4 = -1;
}
__device__ __forceinline__ float top5_dist(u64 k) { return __uint_as_float((unsigned)(k >> 32)); }
__device__ __forceinline__ int top5_orig(u64 k) { return (int)(unsigned)(k & 0xffffffffull); }

__device__ __forceinline__ void cswap(u64& ka, int& pa, u64& kb, int& pb) {   // afterwards ka <= kb
    const bool s = kb < ka;
    const u64 lo = s ? kb : ka, hi = s ? ka : kb;
    const int plo = s ? pb : pa, phi = s ? pa : pb;
    ka = lo; kb = hi; pa = plo; pb = phi;
}

__device__ __forceinline__ void top5_insert(Top5& t, u64 k, int p) {
    if (k < t.k4) {
        t.k4 = k; t.p4 = p;
        cswap(t.k3, t.p3, t.k4, t.p4);
        cswap(t.k2, t.p2, t.k3, t.p3);
        cswap(t.k1, t.p1, t.k2, t.p2);
        cswap(t.k0, t.p0, t.k1, t.p1);
    }
}

__device__ __forceinline__ u64 make_key(float sx, float sy, float sz, const float4& m) {
    // FLANN L2_Simple<float>: ((dx*dx) + dy*dy) + dz*dz with no contraction (exact-op intrinsics)
    const float dx = fsubx(sx, m.x), dy = fsubx(sy, m.y), dz = fsubx(sz, m.z);
    const float d = faddx(faddx(fmulx(dx, dx), fmulx(dy, dy)), fmulx(dz, dz));
    return ((u64)__float_as_uint(d) << 32) | (u64)(unsigned)__float_as_int(m.w);
}

// LANES (8, 16 or 32) lanes with mask `gmask` and lane-in-group `sub` search the 3x3x3 cell block
// around (sx,sy,sz).  The three x-adjacent cells of a (y,z) row are ONE contiguous run of the
// cell-sorted map, so the block is 9 coalesced runs.  On return every lane of the group holds the
// same sorted top-5.  `cand` (lane sub==0 only) accumulates the number of map points examined.
template <int LANES>
__device__ __forceinline__ void group_knn5(float sx, float sy, float sz, const float4* __restrict__ map,
                                           const int* __restrict__ cell_start, const GridDesc& g, int sub, unsigned gmask,
                                           Top5& top, unsigned long long& cand) {
    const int cx = cell_coord(sx, g.inv_cell) - g.org[0];
    const int cy = cell_coord(sy, g.inv_cell) - g.org[1];
    const int cz = cell_coord(sz, g.inv_cell) - g.org[2];
    const int x0 = max(cx - 1, 0), x1 = min(cx + 1, g.dim[0] - 1);
    // lane r (< 9) fetches the [begin,end) of row r; with 8 lanes, lane 0 also fetches row 8
    int rs = 0, re = 0, rs8 = 0, re8 = 0;
    if (x0 <= x1) {
        if (sub < 9) {
            const int y = cy + (sub % 3) - 1, z = cz + (sub / 3) - 1;
            if (y >= 0 && y < g.dim[1] && z >= 0 && z < g.dim[2]) {
                const int base = (z * g.dim[1] + y) * g.dim[0];
                rs = __ldg(cell_start + base + x0);
                re = __ldg(cell_start + base + x1 + 1);
            }
        }
        if (LANES == 8 && sub == 0) {
            const int y = cy + 1, z = cz + 1;
            if (y >= 0 && y < g.dim[1] && z >= 0 && z < g.dim[2]) {
                const int base = (z * g.dim[1] + y) * g.dim[0];
                rs8 = __ldg(cell_start + base + x0);
                re8 = __ldg(cell_start + base + x1 + 1);
            }
        }
    }
    // software-pipelined walk over the 9 runs: the first candidate of run r+1 is in flight while run r is ranked
    int b = __shfl_sync(gmask, rs, 0, LANES), e = __shfl_sync(gmask, re, 0, LANES);
    int p = b + sub;
    float4 cur = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p < e) cur = __ldg(map + p);
#pragma unroll 1
    for (int row = 0; row < 9; ++row) {
        int nb = 0, ne = 0;
        if (row < 8) {
            const int nr = row + 1;
            if (LANES == 8 && nr == 8) { nb = __shfl_sync(gmask, rs8, 0, LANES); ne = __shfl_sync(gmask, re8, 0, LANES); }
            else { nb = __shfl_sync(gmask, rs, nr, LANES); ne = __shfl_sync(gmask, re, nr, LANES); }
        }
        const int np = nb + sub;
        float4 nxt = make_float4(0.f, 0.f, 0.f, 0.f);
        if (np < ne) nxt = __ldg(map + np);
        if (sub == 0) cand += (unsigned long long)(e - b);
        if (p < e) {
            top5_insert(top, make_key(sx, sy, sz, cur), p);
            for (int pp = p + LANES; pp < e; pp += LANES) {       // runs longer than LANES points (rare)
                const float4 m = __ldg(map + pp);
                top5_insert(top, make_key(sx, sy, sz, m), pp);
            }
        }
        b = nb; e = ne; p = np; cur = nxt;
    }
    // merge: 5 rounds of "group-wide minimum of the list heads, winner pops" with REDUX.MIN on the two
    // key halves (keys are unique: a map point lives in exactly one lane's list), ~25 instructions a round
    Top5 res;
#define LILI_MERGE_ROUND(KJ, PJ)                                                        \
    {                                                                                    \
        const unsigned hi = (unsigned)(top.k0 >> 32), lo = (unsigned)top.k0;             \
        const unsigned mhi = __reduce_min_sync(gmask, hi);                               \
        const unsigned mlo = __reduce_min_sync(gmask, hi == mhi ? lo : 0xffffffffu);     \
        const bool win = (hi == mhi) && (lo == mlo);                                     \
        const unsigned bal = __ballot_sync(gmask, win);                                  \
        PJ = __shfl_sync(gmask, top.p0, __ffs(bal) - 1);                                 \
        KJ = ((u64)mhi << 32) | (u64)mlo;                                                \
        if (win) {                                                                       \
            top.k0 = top.k1; top.k1 = top.k2; top.k2 = top.k3; top.k3 = top.k4; top.k4 = ~0ull; \
            top.p0 = top.p1; top.p1 = top.p2; top.p2 = top.p3; top.p3 = top.p4; top.p4 = -1;    \
        }                                                                                \
    }
    LILI_MERGE_ROUND(res.k0, res.p0)
    LILI_MERGE_ROUND(res.k1, res.p1)
    LILI_MERGE_ROUND(res.k2, res.p2)
    LILI_MERGE_ROUND(res.k3, res.p3)
    LILI_MERGE_ROUND(res.k4, res.p4)
#undef LILI_MERGE_ROUND
    top = res;
}

}  // namespace lili
