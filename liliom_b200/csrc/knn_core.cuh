// Octet-cooperative exact 5-NN over the dense cell grid (shared by the scan-to-map kernel and
// the backend correspondence kernels).  See grid_knn.cu for the design notes.
#pragma once
#include "ctx.cuh"
#include "dev_math.cuh"

namespace lili {

__device__ __forceinline__ int cell_coord(float v, float inv_cell) { return (int)floorf(v * inv_cell); }

constexpr int kLanes = 8;               // lanes per query
constexpr int kBlock = 128;
constexpr int kWarps = kBlock / 32;

struct Slot { int pos[5]; float sx, sy, sz; };   // 32 B

__device__ __forceinline__ unsigned block_hash(int bx, int by, int bz) {
    unsigned h = (unsigned)bx * 73856093u ^ (unsigned)by * 19349663u ^ (unsigned)bz * 83492791u;
    h ^= h >> 15; h *= 0x2c1b3c6du; h ^= h >> 12;
    return h;
}
__device__ __forceinline__ int owner_of(float x, float y, float z, int nranks) {
    int bx = (int)floorf(x * 0.125f), by = (int)floorf(y * 0.125f), bz = (int)floorf(z * 0.125f);
    return (int)(block_hash(bx, by, bz) % (unsigned)nranks);
}

struct Top5 {
    float d0, d1, d2, d3, d4;
    int p0, p1, p2, p3, p4;
};

// strict ordering (distance, then original index) — ties are broken by the index stored in .w
__device__ __forceinline__ bool cand_less(float da, int pa, float db, int pb, const float4* __restrict__ map) {
    if (da < db) return true;
    if (da > db) return false;
    if (pb < 0) return true;
    if (pa < 0) return false;
    return __float_as_int(map[pa].w) < __float_as_int(map[pb].w);
}

__device__ __forceinline__ void top5_insert(Top5& t, float d, int p, const float4* __restrict__ map) {
    if (!cand_less(d, p, t.d4, t.p4, map)) return;
    t.d4 = d; t.p4 = p;
    if (cand_less(t.d4, t.p4, t.d3, t.p3, map)) { float a = t.d3; int b = t.p3; t.d3 = t.d4; t.p3 = t.p4; t.d4 = a; t.p4 = b; } else return;
    if (cand_less(t.d3, t.p3, t.d2, t.p2, map)) { float a = t.d2; int b = t.p2; t.d2 = t.d3; t.p2 = t.p3; t.d3 = a; t.p3 = b; } else return;
    if (cand_less(t.d2, t.p2, t.d1, t.p1, map)) { float a = t.d1; int b = t.p1; t.d1 = t.d2; t.p1 = t.p2; t.d2 = a; t.p2 = b; } else return;
    if (cand_less(t.d1, t.p1, t.d0, t.p0, map)) { float a = t.d0; int b = t.p0; t.d0 = t.d1; t.p0 = t.p1; t.d1 = a; t.p1 = b; }
}


// 8 lanes (mask `omask`, lane-in-octet `sub`) search the 3x3x3 cell block around (sx,sy,sz).
// On return every lane of the octet holds the same sorted top-5 (distance, position in `map`).
// `cand` (lane sub==0 only) accumulates the number of map points examined.
__device__ __forceinline__ void octet_knn5(float sx, float sy, float sz, const float4* __restrict__ map,
                                           const int* __restrict__ cell_start, const GridDesc& g, int sub, unsigned omask,
                                           Top5& top, unsigned long long& cand) {
    const int cx = cell_coord(sx, g.inv_cell) - g.org[0];
    const int cy = cell_coord(sy, g.inv_cell) - g.org[1];
    const int cz = cell_coord(sz, g.inv_cell) - g.org[2];
    const int x0 = max(cx - 1, 0), x1 = min(cx + 1, g.dim[0] - 1);
    // row ranges: lane `sub` owns row `sub` (0..7); sub 0 also owns row 8.  The three x-adjacent
    // cells of a row are one contiguous run of the cell-sorted map.
    int rs = 0, re = 0, rs8 = 0, re8 = 0;
    if (x0 <= x1) {
        {
            int y = cy + (sub % 3) - 1, z = cz + (sub / 3) - 1;
            if (y >= 0 && y < g.dim[1] && z >= 0 && z < g.dim[2]) {
                int base = (z * g.dim[1] + y) * g.dim[0];
                rs = __ldg(cell_start + base + x0);
                re = __ldg(cell_start + base + x1 + 1);
            }
        }
        if (sub == 0) {
            int y = cy + 1, z = cz + 1;
            if (y >= 0 && y < g.dim[1] && z >= 0 && z < g.dim[2]) {
                int base = (z * g.dim[1] + y) * g.dim[0];
                rs8 = __ldg(cell_start + base + x0);
                re8 = __ldg(cell_start + base + x1 + 1);
            }
        }
    }
#pragma unroll
    for (int row = 0; row < 9; ++row) {
        int b, e;
        if (row < 8) { b = __shfl_sync(omask, rs, row, kLanes); e = __shfl_sync(omask, re, row, kLanes); }
        else { b = __shfl_sync(omask, rs8, 0, kLanes); e = __shfl_sync(omask, re8, 0, kLanes); }
        if (sub == 0) cand += (unsigned long long)(e - b);
        for (int p = b + sub; p < e; p += kLanes) {
            float4 m = __ldg(map + p);
            // FLANN L2_Simple<float>: ((dx*dx) + dy*dy) + dz*dz, no contraction
            float dx = fsubx(sx, m.x), dy = fsubx(sy, m.y), dz = fsubx(sz, m.z);
            float d = faddx(faddx(fmulx(dx, dx), fmulx(dy, dy)), fmulx(dz, dz));
            top5_insert(top, d, p, map);
        }
    }
    // merge the 8 private lists (butterfly inside the octet)
#pragma unroll
    for (int o = 1; o < kLanes; o <<= 1) {
        float e0 = __shfl_xor_sync(omask, top.d0, o), e1 = __shfl_xor_sync(omask, top.d1, o),
              e2 = __shfl_xor_sync(omask, top.d2, o), e3 = __shfl_xor_sync(omask, top.d3, o),
              e4 = __shfl_xor_sync(omask, top.d4, o);
        int g0 = __shfl_xor_sync(omask, top.p0, o), g1 = __shfl_xor_sync(omask, top.p1, o),
            g2 = __shfl_xor_sync(omask, top.p2, o), g3 = __shfl_xor_sync(omask, top.p3, o),
            g4 = __shfl_xor_sync(omask, top.p4, o);
        if (g0 >= 0) top5_insert(top, e0, g0, map);
        if (g1 >= 0) top5_insert(top, e1, g1, map);
        if (g2 >= 0) top5_insert(top, e2, g2, map);
        if (g3 >= 0) top5_insert(top, e3, g3, map);
        if (g4 >= 0) top5_insert(top, e4, g4, map);
    }
}

}  // namespace lili
