// Scan-to-map on sm_100a: dense 1 m cell grid over the down-sampled map (stand-in for the
// per-scan FLANN kd-tree, L/src/LidarOdometry.cpp:490) and the fused
//   transform -> exact 5-NN -> 5x3 QR plane -> gates -> residual + SE(3) Jacobian row ->
//   Huber -> 21+6 reduction -> 6x6 solve -> pose update
// kernel replacing findCorrespondingSurfFeatures + the Ceres problem/solve
// (L/src/LidarOdometry.cpp:352-413, 506-561; L/include/factors/LidarKeyframeFactor.h:111-139).
//
// HBM-bound integer/float gather work: no dense contraction, so no tensor cores.  Design:
//   * map points stored as float4 {x,y,z,orig_index} sorted by cell id (z,y,x order) so the
//     3 x-adjacent cells of a row are ONE contiguous run: 9 coalesced runs per query;
//   * 8 lanes ("octet") cooperate on one query: lanes stride through each run with 16-byte
//     loads, keep a private sorted top-5 in registers, then merge with 3 xor-shuffles;
//   * a warp finishes 4*R queries, parks the 5 neighbour slots in shared memory, and then every
//     lane runs the fp64 plane fit / Jacobian of a DIFFERENT query (no redundant fp64 issue);
//   * 29 fp64 partial sums per block -> fixed-order cross-block sum by the last block (ticket),
//     which also solves the 6x6 system and updates the pose in HBM: one launch per iteration,
//     the pose never visits the host inside the loop, results are run-to-run deterministic.
// Exactness: a query's cell block covers its full `cell`-metre ball and cells are aligned to
// the integer lattice (floorf(x * 2^-k) is exact), so any point outside the 27 cells is at
// fp32 distance >= cell >= sqrt(knn_max_sqdist): the accepted 5-NN sets equal the kd-tree's.
#include "ctx.cuh"
#include "dev_math.cuh"
#include "knn_core.cuh"
#include <cstdlib>

namespace lili {

// ------------------------------------------------------------------ small utilities
__global__ void k_repack_f4(const unsigned char* __restrict__ in, int n_max, const int* __restrict__ d_n, int stride, float4* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = d_n ? min(*d_n, n_max) : n_max;
    if (i >= n) return;
    float4 v = *reinterpret_cast<const float4*>(in + (size_t)i * stride);
    v.w = __int_as_float(i);
    out[i] = v;
}

int repack_to_f4(liliom_ctx* c, const void* d_in, int n, int stride, float4* d_out, const int* d_n) {
    if (n <= 0) return LILIOM_OK;
    k_repack_f4<<<cdiv(n, 256), 256, 0, c->stream>>>((const unsigned char*)d_in, n, d_n, stride, d_out);
    return launch_check(c, "k_repack_f4");
}

__device__ __forceinline__ int f2ord(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

// mm[0..2] = min xyz (ordered-int encoding), mm[3..5] = max
__global__ void k_minmax_f4(const float4* __restrict__ p, int n, int* __restrict__ mm) {
    int lo[3] = {INT_MAX, INT_MAX, INT_MAX}, hi[3] = {INT_MIN, INT_MIN, INT_MIN};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float4 v = p[i];
        int a = f2ord(v.x), b = f2ord(v.y), cz = f2ord(v.z);
        lo[0] = min(lo[0], a); hi[0] = max(hi[0], a);
        lo[1] = min(lo[1], b); hi[1] = max(hi[1], b);
        lo[2] = min(lo[2], cz); hi[2] = max(hi[2], cz);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[k] = min(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
            hi[k] = max(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
        }
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { atomicMin(&mm[k], lo[k]); atomicMax(&mm[3 + k], hi[k]); }
    }
}

__global__ void k_init_minmax(int* mm) {
    if (threadIdx.x < 3) mm[threadIdx.x] = INT_MAX;
    else if (threadIdx.x < 6) mm[threadIdx.x] = INT_MIN;
}

// `escaped` (optional): set when a point lies outside the grid's box — only possible when the box was handed in by the caller
// (grid_build's host_box); the key is then clamped into range for memory safety and the caller rebuilds with the exact box.
__global__ void k_cell_keys(const float4* __restrict__ p, int n, GridDesc g, uint32_t* __restrict__ keys, int* __restrict__ vals,
                            int* __restrict__ escaped) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 v = p[i];
    int cx = cell_coord(v.x, g.inv_cell) - g.org[0];
    int cy = cell_coord(v.y, g.inv_cell) - g.org[1];
    int cz = cell_coord(v.z, g.inv_cell) - g.org[2];
    if (escaped && ((unsigned)cx >= (unsigned)g.dim[0] || (unsigned)cy >= (unsigned)g.dim[1] || (unsigned)cz >= (unsigned)g.dim[2])) {
        *escaped = 1;
        cx = min(max(cx, 0), g.dim[0] - 1); cy = min(max(cy, 0), g.dim[1] - 1); cz = min(max(cz, 0), g.dim[2] - 1);
    }
    keys[i] = (uint32_t)((cz * g.dim[1] + cy) * g.dim[0] + cx);
    vals[i] = i;
}

__global__ void k_gather_sorted(const float4* __restrict__ p, const int* __restrict__ vals, int n, float4* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int src = vals[i];
    float4 v = p[src];
    out[i] = make_float4(v.x, v.y, v.z, __int_as_float(src));   // w = index into the un-sorted map array
}

// cell_start[c] = first sorted position whose key >= c = number of points with key < c = upper bound of cell c-1.
// Run ENDS are scattered (cell_start[key+1] = position after the run) into a zeroed table and a running maximum fills the
// empty cells: two streaming passes over the table.  (The first version filled every gap with one warp per sorted
// position: 724 us for a 10 M-point map, profiles/r02_stream_1gpu_launches.txt; this is ~60 us.)
__global__ void k_cell_run_ends(const uint32_t* __restrict__ keys, int n, int* __restrict__ cell_start) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n) return;
    const uint32_t k = keys[w];
    if (w == n - 1 || keys[w + 1] != k) cell_start[(size_t)k + 1] = w + 1;
}

int grid_build(liliom_ctx* c, int m, const int* host_box) {
    c->map_ready = false;
    c->map_n = m;
    if (m <= 0) {   // empty (shard of the) map: a 1-cell grid with no points, so that collective callers still run every launch
        GridDesc g{};
        g.inv_cell = 1.0f; g.dim[0] = g.dim[1] = g.dim[2] = 1; g.ncells = 1;
        c->grid = g;
        LILI_CUDA(c, c->cell_start.ensure(4 * sizeof(int)));
        LILI_CUDA(c, cudaMemsetAsync(c->cell_start.p, 0, 4 * sizeof(int), c->stream));
        LILI_CUDA(c, c->map_sorted.ensure(sizeof(float4)));
        LILI_CUDA(c, c->map_xyzw.ensure(sizeof(float4)));
        c->map_n = 0;
        c->map_ready = true;
        return LILIOM_OK;
    }
    float4* pts = c->map_xyzw.as<float4>();
    int h[6];
    // A box handed in by the caller contains the raw points.  A voxel centroid (a sequential fp32 mean of such points) can leave
    // it by rounding, in which case its cell may not exist: k_cell_keys reports that, and the grid is rebuilt from the exact
    // min/max below.  (One cell of padding instead would be free of the re-run but can push the cell count over a power of
    // 256 and cost every rebuild a whole extra radix pass: 2^24 -> 17.7 M cells on the 10 M-point bench map, +90 us.)
    int* escaped = nullptr;
    if (host_box) {
        for (int k = 0; k < 6; ++k) h[k] = host_box[k];
        if (getenv("LILIOM_TEST_SHRINK_BOX")) h[3] = h[0];      // test hook: a box that is too small in x -> the escape path below must run
        LILI_CUDA(c, c->vg_minmax.ensure(8 * sizeof(int)));
        escaped = c->vg_minmax.as<int>() + 7;
        LILI_CUDA(c, cudaMemsetAsync(escaped, 0, sizeof(int), c->stream));
    } else {
        LILI_CUDA(c, c->vg_minmax.ensure(8 * sizeof(int)));
        int* mm = c->vg_minmax.as<int>();
        k_init_minmax<<<1, 32, 0, c->stream>>>(mm);
        LILI_TRY(launch_check(c, "k_init_minmax"));
        k_minmax_f4<<<min(cdiv(m, 256), c->sm_count * 8), 256, 0, c->stream>>>(pts, m, mm);
        LILI_TRY(launch_check(c, "k_minmax_f4"));
        LILI_CUDA(c, cudaMemcpyAsync(h, mm, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
        LILI_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    auto dec = [](int i) { int j = i >= 0 ? i : i ^ 0x7fffffff; float f; memcpy(&f, &j, 4); return f; };
    // cell size: smallest power of two >= sqrt(knn_max_sqdist) (1.0 for the reference's gate)
    float cell = 1.0f;
    while ((double)cell * (double)cell < c->prm.knn_max_sqdist) cell *= 2.0f;
    GridDesc g;
    g.inv_cell = 1.0f / cell;
    long long nc = 1;
    for (int k = 0; k < 3; ++k) {
        float lo = dec(h[k]), hi = dec(h[3 + k]);
        if (!(std::isfinite(lo) && std::isfinite(hi))) { c->last_error = "non-finite map point"; return LILIOM_E_ARG; }
        int clo = (int)floorf(lo * g.inv_cell), chi = (int)floorf(hi * g.inv_cell);
        g.org[k] = clo;
        g.dim[k] = chi - clo + 1;
        nc *= g.dim[k];
    }
    if (nc > (1LL << 29)) { c->last_error = "map extent needs more than 2^29 cells"; return LILIOM_E_GRID; }
    g.ncells = (int)nc;
    c->grid = g;
    LILI_CUDA(c, c->grid_keys.ensure((size_t)m * 4));
    LILI_CUDA(c, c->grid_keys2.ensure((size_t)m * 4));
    LILI_CUDA(c, c->grid_vals.ensure((size_t)m * 4));
    LILI_CUDA(c, c->grid_vals2.ensure((size_t)m * 4));
    LILI_CUDA(c, c->map_sorted.ensure((size_t)m * sizeof(float4)));
    LILI_CUDA(c, c->cell_start.ensure(((size_t)g.ncells + 2) * 4));
    k_cell_keys<<<cdiv(m, 256), 256, 0, c->stream>>>(pts, m, g, c->grid_keys.as<uint32_t>(), c->grid_vals.as<int>(), escaped);
    LILI_TRY(launch_check(c, "k_cell_keys"));
    int bits = 1;
    while ((1LL << bits) < nc) ++bits;
    LILI_TRY(sort_pairs_u32(c, c->grid_keys.as<uint32_t>(), c->grid_keys2.as<uint32_t>(), c->grid_vals.as<int>(),
                            c->grid_vals2.as<int>(), m, bits));
    k_gather_sorted<<<cdiv(m, 256), 256, 0, c->stream>>>(pts, c->grid_vals2.as<int>(), m, c->map_sorted.as<float4>());
    LILI_TRY(launch_check(c, "k_gather_sorted"));
    LILI_CUDA(c, cudaMemsetAsync(c->cell_start.p, 0, ((size_t)g.ncells + 2) * 4, c->stream));
    k_cell_run_ends<<<cdiv(m, 256), 256, 0, c->stream>>>(c->grid_keys2.as<uint32_t>(), m, c->cell_start.as<int>());
    LILI_TRY(launch_check(c, "k_cell_run_ends"));
    LILI_TRY(inclusive_max_scan_i32(c, c->cell_start.as<int>(), g.ncells + 1));
    if (escaped) {
        int* hp = reinterpret_cast<int*>(c->h_pin) + 1048;
        LILI_CUDA(c, cudaMemcpyAsync(hp, escaped, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        LILI_CUDA(c, cudaStreamSynchronize(c->stream));
        if (hp[0]) return grid_build(c, m, nullptr);
    }
    c->map_ready = true;
    return LILIOM_OK;
}

// ------------------------------------------------------------------ the hot kernel
struct KnnArgs {
    const float4* feats; int n; const int* n_dev;   // n = host upper bound, n_dev = optional device-side count
    const float4* map; const float4* map_orig; const int* cell_start; GridDesc g;   // cell-sorted / download-order map
    const double* pose;                 // 7 doubles in HBM
    double max_sqd, plane_thres, w_gate, huber_a;
    unsigned char* valid; float4* plane; int* nn_idx; float* nn_sqd;   // optional outputs
    double* partials; double* neq; unsigned int* ticket;
    double* stats;                      // this iteration's stats slot (kStatsDoubles) or null
    double* pose_out;                   // where the updated pose goes (== pose for GN)
    unsigned long long* cand_total;     // instrumentation
    int update_pose;                    // 1: last block performs the GN step
    int rounds;                         // a warp handles (32/LANES)*rounds queries per task
    int nranks, rank;                   // multi-GPU ownership filter (16 m block hash)
    long long* dbg;                     // optional phase timestamps (LILIOM_DEBUG_TIMING)
    int interleave;                     // 1: deal tasks to warps block-cyclically (load balance for small scans)
    float tau0;                         // largest fp32 distance inside the gate (knn_gate_tau(max_sqd)): search pruning threshold
    float4* qstate;                     // per query {transformed position, fifth distance} of the previous pass of this call (see coherence_tau)
    int use_state;                      // 1: qstate holds the previous pass (per-iteration launches; the persistent kernel sets it per pass)
    float inv_block;                    // 1 / shard block edge (multi-GPU ownership)
    unsigned long long* queries_total;  // instrumentation: queries processed (device-side count), one add per pass
};

// per-thread instrumentation word: examined candidates in the low 40 bits, searched queries above (summed per block)
constexpr int kCandBits = 40;
constexpr unsigned long long kCandMask = (1ull << kCandBits) - 1ull;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Temporal coherence of the GN passes: a query whose five neighbours lay within r = sqrt(d5) at its previous position p'
// still has those five map points within r + |p - p'| of its new position p, so its new fifth distance is at most
// (r + |p - p'|)^2.  Starting the search with that bound instead of the gate skips most cells outright and leaves little to
// rank — from the second pass of a scan on the poses move by centimetres.  The bound is evaluated with upward roundings and
// a 1e-5 relative margin (the candidate distance expression carries ~3 fp32 roundings, ~2e-7), so every point of the true
// 5-NN set passes the filter: the sets found, their order and the accept decisions are exactly those of the full search.
__device__ __forceinline__ float coherence_tau(const float4& st, float sx, float sy, float sz, float tau0) {
    if (!(st.w <= tau0)) return tau0;                    // no five neighbours inside the gate last time (or NaN): nothing to exploit
    const float dx = fsubx(sx, st.x), dy = fsubx(sy, st.y), dz = fsubx(sz, st.z);
    const float mv = __fsqrt_ru(__fadd_ru(__fadd_ru(__fmul_ru(dx, dx), __fmul_ru(dy, dy)), __fmul_ru(dz, dz)));
    const float r = __fadd_ru(__fsqrt_ru(st.w), mv);
    return fminf(tau0, __fmul_ru(__fmul_ru(r, r), 1.00001f));
}

// Per-query scratch a warp keeps in shared memory between the phases.
struct Slot {
    int   idx[5];            // neighbour indices into map_orig; idx[0] < 0: no 5-NN inside the radius
    float sx, sy, sz;        // transformed query (fp32, as the kd-tree saw it)
    float fx, fy, fz;        // body-frame query (phase B re-derives R p from it)
    int   pad;               // 48 bytes
};
struct Row { double J[6]; double r; double half_rho; };   // robustified Jacobian row, residual, rho/2

// scalar k of the 29-vector is  sum_i row.J[kA[k]] * row.X[kB[k]]  (X = J for k<21, r for 21..26)
__constant__ unsigned char kPairA[27] = {0,0,0,0,0,0, 1,1,1,1,1, 2,2,2,2, 3,3,3, 4,4, 5, 0,1,2,3,4,5};
__constant__ unsigned char kPairB[27] = {0,1,2,3,4,5, 1,2,3,4,5, 2,3,4,5, 3,4,5, 4,5, 5, 6,6,6,6,6,6};

__device__ __forceinline__ double q_t7(const Q4& q, const D3& t, int k) {
    return k == 0 ? q.w : k == 1 ? q.x : k == 2 ? q.y : k == 3 ? q.z : k == 4 ? t.x : k == 5 ? t.y : t.z;
}

struct KnnSmem {
    Slot slots[kWarps][32];
    Row rows[kWarps][32];
    unsigned char rvalid[kWarps][32];
    double red[kWarps][kNormEq];
    double xch[kMaxPeers][kNormEq];   // fused multi-GPU exchange: every rank's 29 sums before they are added in rank order
    unsigned long long red_cand[kWarps];
    double pose[8];
    int is_last;
    int peer_lost;               // fused exchange: a peer did not publish within the wait bound -> the pose is poisoned (NaN), the host reports it
};

#define LILI_STAMP(i) do { if (a.dbg && blockIdx.x == 0 && threadIdx.x == 0) a.dbg[i] = clock64(); } while (0)
#define LILI_STAMP_LAST(i) do { if (a.dbg && threadIdx.x == 0) a.dbg[i] = clock64(); } while (0)

// One pass over this block's share of the queries at pose (q,t): phases A (search), B (fit + row), C (lane k
// accumulates scalar k).  On return lane k < 29 of every warp holds its partial of scalar k in `acc`.
template <int LANES, bool TMA = false>
__device__ __forceinline__ void knn_phases(const KnnArgs& a, const Q4& q, const D3& t, const int n_q, KnnSmem& S,
                                           double& acc, unsigned long long& cand, const bool use_state, const float4* fpre = nullptr,
                                           unsigned* stage_phase = nullptr) {
    constexpr int GROUPS = 32 / LANES;                 // queries a warp searches concurrently
    extern __shared__ __align__(16) unsigned char dyn_smem[];      // LANES == 1: [kRunCap][kBlock] int4 run lists (thread_knn5)
    int4* runs = reinterpret_cast<int4*>(dyn_smem) + threadIdx.x;
    StageTile* stage = nullptr;                                    // TMA: [kWarps][2] staging tiles (bulk-copy path of the 16-lane search)
    if constexpr (TMA) stage = reinterpret_cast<StageTile*>(dyn_smem) + (threadIdx.x >> 5) * 2 + ((threadIdx.x & 31) >> 4);
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int sub = lane & (LANES - 1);
    const int grp = lane / LANES;
    const unsigned gmask = (LANES == 32) ? 0xffffffffu : (((1u << LANES) - 1u) << (grp * LANES));
    // lane k (< 29) owns scalar k of the normal equations: no per-lane 29-vector, no final shuffle tree
    const int pa = lane < 27 ? kPairA[lane] : 0, pb = lane < 27 ? kPairB[lane] : 0;

    const int per_task = GROUPS * a.rounds;            // <= 32
    const int ntasks = (n_q + per_task - 1) / per_task;
    // queries arrive in voxel order (spatially sorted): dealing neighbouring tasks to different blocks evens out the
    // per-block work (dense vs sparse regions) at the grid barrier; large scans keep neighbours together for L1 reuse
    const int gw = a.interleave ? warp * (int)gridDim.x + (int)blockIdx.x : (int)blockIdx.x * kWarps + warp;
    const int nw = gridDim.x * kWarps;

#pragma unroll 1
    for (int task = gw; task < ntasks; task += nw) {
        // ---------------- phase A: cooperative exact 5-NN, one query per lane group per round
#pragma unroll 1
        for (int r = 0; r < a.rounds; ++r) {
            const int slot = r * GROUPS + grp;
            const int qi = task * per_task + slot;
            Top5 top;
            top5_init(top);
            float sx = 0.f, sy = 0.f, sz = 0.f;
            float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
            bool live = qi < n_q;
            if (live) {
                f = fpre ? *fpre : a.feats[qi];   // (persistent kernel: the query stays in registers across iterations)
                const D3 pw = qrot_x(q, D3{(double)f.x, (double)f.y, (double)f.z});             // L/src/LidarOdometry.cpp:230-231
                sx = (float)addx(pw.x, t.x); sy = (float)addx(pw.y, t.y); sz = (float)addx(pw.z, t.z);   // :236-238
                if (a.nranks > 1 && owner_of(sx, sy, sz, a.nranks, a.inv_block) != a.rank) live = false;
            }
            // `live` is uniform inside a lane group and the shuffles are masked per group
            LILI_STAMP(8);
            float tau_q = a.tau0;
            if (live && use_state) tau_q = coherence_tau(__ldcg(a.qstate + qi), sx, sy, sz, a.tau0);
            if (live) {
                if (sub == 0) cand += 1ull << kCandBits;
                if constexpr (LANES == 1) thread_knn5(sx, sy, sz, a.map, a.cell_start, a.g, tau_q, runs, kBlock, top, cand);
                else group_knn5<LANES>(sx, sy, sz, a.map, a.cell_start, a.g, sub, gmask, tau_q, top, cand, (a.dbg && blockIdx.x == 0 && threadIdx.x == 0) ? a.dbg : nullptr,
                                       stage, stage_phase);
            }
            LILI_STAMP(11);
            if (sub == 0) {
                Slot& s = S.slots[warp][slot];
                const bool ok = live && top.k4 != ~0ull && ((double)top5_dist(top.k4) < a.max_sqd);  // :365
                // this pass's position and fifth distance for the next pass (a query this rank does not own keeps no bound)
                if (qi < n_q) a.qstate[qi] = make_float4(sx, sy, sz, ok ? top5_dist(top.k4) : __int_as_float(0x7f800000));
                s.idx[0] = ok ? top5_index(top.k0) : -1; s.idx[1] = top5_index(top.k1); s.idx[2] = top5_index(top.k2);
                s.idx[3] = top5_index(top.k3); s.idx[4] = top5_index(top.k4);
                s.sx = sx; s.sy = sy; s.sz = sz;
                s.fx = f.x; s.fy = f.y; s.fz = f.z;
                s.pad = 0;
                if (live && a.nn_idx) {
                    int* o = a.nn_idx + (size_t)qi * 5;
                    const u64 kk[5] = {top.k0, top.k1, top.k2, top.k3, top.k4};
#pragma unroll
                    for (int j = 0; j < 5; ++j) { const int li = top5_index(kk[j]); o[j] = li >= 0 ? __float_as_int(a.map_orig[li].w) : -1; }
                }
                if (live && a.nn_sqd) {
                    float* o = a.nn_sqd + (size_t)qi * 5;
                    o[0] = top5_dist(top.k0); o[1] = top5_dist(top.k1); o[2] = top5_dist(top.k2); o[3] = top5_dist(top.k3); o[4] = top5_dist(top.k4);
                }
            }
        }
        __syncwarp();
        LILI_STAMP(1);
        // ---------------- phase B: one lane per query — plane fit, gates, residual, Jacobian row
        bool ok = false;
        if (lane < per_task) {
            const int qi = task * per_task + lane;
            const Slot s = S.slots[warp][lane];
            float pl0 = 0.f, pl1 = 0.f, pl2 = 0.f, pl3 = 0.f;
            if (qi < n_q && s.idx[0] >= 0) {
                float4 m[5];
#pragma unroll
                for (int j = 0; j < 5; ++j) m[j] = __ldg(a.map_orig + s.idx[j]);                                     // :369-371
                double nv[3];
                if (!plane_fit5_fast(m, nv)) plane_fit5_qr(m, nv);                                                    // :375 (see dev_math.cuh)
                const double n2 = nv[0] * nv[0] + nv[1] * nv[1] + nv[2] * nv[2];
                const double nn = sqrt(n2);
                const double normInverse = 1.0 / nn;                                                                  // :376
                if (n2 > 0) { nv[0] /= nn; nv[1] /= nn; nv[2] /= nn; }                                                // :377
                bool planeValid = true;
#pragma unroll
                for (int j = 0; j < 5; ++j)                                                                            // :386-393
                    if (fabs(nv[0] * m[j].x + nv[1] * m[j].y + nv[2] * m[j].z + normInverse) > a.plane_thres) planeValid = false;
                if (planeValid) {
                    // :397-398 with the reference's mixed widths; exact ops keep the fp32 roundings stable
                    const float pd = (float)addx(addx(addx(mulx(nv[0], (double)s.sx), mulx(nv[1], (double)s.sy)), mulx(nv[2], (double)s.sz)), normInverse);
                    const float rng = __fsqrt_rn(__fsqrt_rn(faddx(faddx(fmulx(s.sx, s.sx), fmulx(s.sy, s.sy)), fmulx(s.sz, s.sz))));
                    const float weight = (float)subx(1.0, mulx(0.9, (double)fabsf(pd)) / (double)rng);
                    if ((double)weight > a.w_gate) {                                                                  // :400
                        pl0 = (float)mulx((double)weight, nv[0]);                                                     // :402-405
                        pl1 = (float)mulx((double)weight, nv[1]);
                        pl2 = (float)mulx((double)weight, nv[2]);
                        pl3 = (float)mulx((double)weight, normInverse);
                        ok = true;
                    }
                }
            }
            if (qi < n_q) {
                if (a.valid) a.valid[qi] = ok ? 1 : 0;
                if (a.plane) a.plane[qi] = make_float4(pl0, pl1, pl2, pl3);
            }
            if (ok) {
                // LidarPlaneNormIncreFactor (LidarKeyframeFactor.h:118-128) in closed form:
                //   r = n~ . (q*p + t) + d~ ;  row = [ 2 (R p x n~)^T , n~^T ]  
                const D3 rp = qrot_x(q, D3{(double)s.fx, (double)s.fy, (double)s.fz});
                const double nx = pl0, ny = pl1, nz = pl2;
                double r = nx * (rp.x + t.x) + ny * (rp.y + t.y) + nz * (rp.z + t.z) + (double)pl3;
                // ceres::HuberLoss(a) + Corrector (rho'' <= 0 branch): scale row and residual by sqrt(rho')
                const double s2 = r * r, b2 = a.huber_a * a.huber_a;
                double rho0 = s2, rho1 = 1.0;
                if (s2 > b2) { const double rt = sqrt(s2); rho0 = 2.0 * a.huber_a * rt - b2; rho1 = fmax(DBL_MIN, a.huber_a / rt); }
                const double sr = sqrt(rho1);
                Row& R = S.rows[warp][lane];
                R.J[0] = sr * 2.0 * (rp.y * nz - rp.z * ny);
                R.J[1] = sr * 2.0 * (rp.z * nx - rp.x * nz);
                R.J[2] = sr * 2.0 * (rp.x * ny - rp.y * nx);
                R.J[3] = sr * nx; R.J[4] = sr * ny; R.J[5] = sr * nz;
                R.r = sr * r;
                R.half_rho = 0.5 * rho0;
            }
            S.rvalid[warp][lane] = ok ? 1 : 0;
        }
        __syncwarp();
        LILI_STAMP(2);
        // ---------------- phase C: lane k accumulates scalar k over the task's rows (fixed slot order)
        // (four interleaved partial sums, folded in a fixed order: a 32-row task is a chain of 8 dependent fp64 adds instead of 32)
        if (lane < kNormEq) {
            double p4[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 1
            for (int s0 = 0; s0 < per_task; s0 += 4) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int sl = s0 + u;
                    if (sl < per_task && S.rvalid[warp][sl]) {
                        const double* R = reinterpret_cast<const double*>(&S.rows[warp][sl]);
                        p4[u] += lane < 27 ? R[pa] * R[pb] : lane == 27 ? R[7] : 1.0;
                    }
                }
            }
            acc += (p4[0] + p4[1]) + (p4[2] + p4[3]);
        }
        __syncwarp();
    }

}

// Block partial -> partials[29][G] (scalar-major) + candidate counter.
__device__ __forceinline__ void write_block_partials(const KnnArgs& a, KnnSmem& S, double acc, unsigned long long cand) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane < kNormEq) S.red[warp][lane] = acc;
    {
        unsigned long long v = cand;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) S.red_cand[warp] = v;
    }
    __syncthreads();
    const int G = gridDim.x;
    if (threadIdx.x < kNormEq) {
        double v = 0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) v += S.red[w][threadIdx.x];
        a.partials[(size_t)threadIdx.x * G + blockIdx.x] = v;          // scalar-major: [29][G]
    }
    if (threadIdx.x == 0 && a.cand_total) {     // instrumentation (liliom_set_kernel_timing): examined candidates | searched queries
        unsigned long long v = 0;
        for (int w = 0; w < kWarps; ++w) v += S.red_cand[w];
        if (v & kCandMask) atomicAdd(a.cand_total, v & kCandMask);
        if (v >> kCandBits) atomicAdd(a.queries_total, v >> kCandBits);
    }
}

// Fixed-order sum over the G block partials into S.red[0][0..28]: 8 lanes per scalar.  This is exposed latency, so
// every load of a lane is issued before the first add: ONE L2 round trip for up to 8*kRedLoads = 160 blocks (the
// persistent kernel never has more than one block per SM); larger grids take another trip per 160 blocks.
constexpr int kRedLoads = 20;
__device__ __forceinline__ void reduce_partials(const KnnArgs& a, KnnSmem& S) {
    const int G = gridDim.x;
    const int sc = threadIdx.x >> 3, l8 = threadIdx.x & 7;
    double v = 0.0;
    if (sc < kNormEq) {
        const double* src = a.partials + (size_t)sc * G;
#pragma unroll 1
        for (int base = l8; base < G; base += 8 * kRedLoads) {
            double w[kRedLoads];
#pragma unroll
            for (int j = 0; j < kRedLoads; ++j) w[j] = (base + 8 * j < G) ? __ldcg(src + base + 8 * j) : 0.0;
            // fixed pairwise tree (deterministic, short dependency chain)
#pragma unroll
            for (int j = 0; j < 10; ++j) w[j] += w[j + 10];
#pragma unroll
            for (int j = 0; j < 5; ++j) w[j] += w[j + 5];
            v += ((w[0] + w[1]) + (w[2] + w[3])) + w[4];
        }
    }
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    if (l8 == 0 && sc < kNormEq) S.red[0][sc] = v;
    __syncthreads();
}

// ---- fused multi-GPU exchange (persistent kernel; liliom_comm_peer_attach) -----------------------------------------------
// Every 8-byte word carries the pass's epoch in its upper half (a double travels as {epoch|lo32, epoch|hi32}): an aligned 8-byte
// store is single-copy atomic, so a reader that sees the epoch sees the data — no fence and no flag word.  Stores and loads are
// system scope (the writer is another GPU over NVLink).  Two buffers alternate by epoch parity: a rank can publish pass e+1 only
// after it has read every rank's pass e, i.e. after every block of every rank has finished reading pass e-1 (they pass their own
// grid barrier of pass e before their block 0 publishes).  Epochs grow monotonically across launches and the buffers are zeroed
// once, so a stale word never matches.  On entry S.red[0][0..28] holds this rank's sums (identical in every block); on return it
// holds the sums over all ranks, added in rank order (bit-identical on every rank).
// `writer`: the block that publishes (block 0 of the persistent kernel; the last block of the per-iteration kernel, the only caller there).
__device__ __forceinline__ void peer_exchange(const PeerArgs& pa, unsigned int epoch, KnnSmem& S, bool writer) {
    const size_t base = (size_t)(epoch & 1u) * kMaxPeers * 32;
    const bool already_lost = S.peer_lost != 0;     // block-uniform (written before the previous exchange's closing barrier): no second wait
    if (writer && threadIdx.x < kNormEq) {
        const double v = S.red[0][threadIdx.x];
        const u64 w0 = ((u64)epoch << 32) | (u64)(unsigned)__double2loint(v);
        const u64 w1 = ((u64)epoch << 32) | (u64)(unsigned)__double2hiint(v);
        for (int p = 0; p < pa.nranks; ++p) {
            ulonglong2* dst = pa.buf[p] + base + (size_t)pa.rank * 32 + threadIdx.x;
            asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(dst), "l"(w0), "l"(w1) : "memory");
        }
    }
    for (int t = threadIdx.x; t < pa.nranks * kNormEq; t += blockDim.x) {
        const int r = t / kNormEq, k = t - r * kNormEq;
        const ulonglong2* src = pa.buf[pa.rank] + base + (size_t)r * 32 + k;
        u64 w0 = 0, w1 = 0;
        unsigned int spins = 0;
        bool lost = already_lost;
        while (!already_lost) {
            asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(src) : "memory");
            if ((unsigned int)(w0 >> 32) == epoch && (unsigned int)(w1 >> 32) == epoch) break;
            if (++spins > (1u << 23)) { lost = true; S.peer_lost = 1; break; }      // ~6 s: a lost peer must end in a NaN pose, never in a hung GPU
            __nanosleep(20);
        }
        S.xch[r][k] = lost ? __longlong_as_double(0x7ff8000000000000ll) : __hiloint2double((int)(unsigned int)w1, (int)(unsigned int)w0);
    }
    __syncthreads();
    if (threadIdx.x < kNormEq) {
        double v = 0.0;
        for (int r = 0; r < pa.nranks; ++r) v += S.xch[r][threadIdx.x];
        S.red[0][threadIdx.x] = v;
    }
    __syncthreads();
}

// Inside one GPU only block 0 talks to the peers: it republishes the sums over all ranks (and the loss flag) for the other
// blocks, which wait on one epoch word instead of polling nranks x 29 system-scope words each (148 blocks doing that cost
// ~20 us per pass on 2 GPUs, profiles/r02_bench_2gpu_first.json).
__device__ __forceinline__ void peer_local_publish(const PeerArgs& pa, unsigned int epoch, KnnSmem& S) {
    double* dst = pa.lsum + (size_t)(epoch & 1u) * 32;
    if (threadIdx.x < kNormEq) dst[threadIdx.x] = S.red[0][threadIdx.x];
    if (threadIdx.x == kNormEq) dst[kNormEq] = S.peer_lost ? 1.0 : 0.0;
    __syncthreads();
    if (threadIdx.x == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(pa.lflag + (epoch & 1u)), "r"(epoch) : "memory");
}
__device__ __forceinline__ void peer_local_wait(const PeerArgs& pa, unsigned int epoch, KnnSmem& S) {
    if (threadIdx.x == 0) {
        unsigned int v;
        do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(pa.lflag + (epoch & 1u)) : "memory"); } while (v != epoch);
    }
    __syncthreads();
    const double* src = pa.lsum + (size_t)(epoch & 1u) * 32;
    if (threadIdx.x < kNormEq) S.red[0][threadIdx.x] = __ldcg(src + threadIdx.x);
    if (threadIdx.x == kNormEq && __ldcg(src + kNormEq) != 0.0) S.peer_lost = 1;
    __syncthreads();
}

// One thread: 6x6 solve of the reduced normal equations in S.red[0], Plus, sign-unify -> xn.
__device__ __forceinline__ void gn_step(const KnnSmem& S, const Q4& q, const D3& t, double xn[7]) {
    double s[kNormEq];
#pragma unroll
    for (int k = 0; k < kNormEq; ++k) s[k] = S.red[0][k];
    double x[7], nb[6], d[6];
#pragma unroll
    for (int k = 0; k < 7; ++k) x[k] = q_t7(q, t, k);
#pragma unroll
    for (int k = 0; k < 6; ++k) nb[k] = -s[21 + k];
    const bool solved = gn_safe_step(s, nb, d);
    if (s[28] > 0.0 && solved) pose_plus(x, d, xn);
    else {
#pragma unroll
        for (int k = 0; k < 7; ++k) xn[k] = x[k];
    }
    if (xn[0] < 0) { xn[0] = -xn[0]; xn[1] = -xn[1]; xn[2] = -xn[2]; xn[3] = -xn[3]; }   // :539-549
    if (S.peer_lost) {
#pragma unroll
        for (int k = 0; k < 7; ++k) xn[k] = __longlong_as_double(0x7ff8000000000000ll);
    }
}

__device__ __forceinline__ void write_neq_stats(const KnnArgs& a, const KnnSmem& S, double* stats, bool with_stats) {
    if (threadIdx.x < kNormEq) {
        const double v = S.red[0][threadIdx.x];
        a.neq[threadIdx.x] = v;
        if (with_stats && stats) {
            if (threadIdx.x < 27) stats[3 + threadIdx.x] = v;
            else if (threadIdx.x == 27) stats[2] = v;
            else { stats[0] = v; stats[1] = 1.0; }
        }
    }
}


// (Search and fit as two kernels — the search alone runs at 64 registers and twice the resident warps — was measured twice,
// with the exhaustive search of round 1 and with the pruned one: 62.3 vs 57.6 us per pass at 128k queries.  The search
// alone takes 32 us at 47 % lane utilisation; what limits the one-thread-per-query shape is divergence between the lanes'
// candidate lists, not latency hiding.  profiles/r02_knn_dense_split_ab.txt, r02_knn_search_only_ncu.txt.)
// Blocks per SM of the one-thread-per-query instance: that shape is bound by latency per issued instruction (8.8 cycles at
// ~4 warps per scheduler, profiles/r02_knn_dense_v1_ncu.txt), so it trades registers (spills in the fp64 fit) for resident warps.
#ifndef LILI_KNN1_MINBLOCKS
#define LILI_KNN1_MINBLOCKS 2
#endif
template <int LANES>
__global__ void __launch_bounds__(kBlock, LANES == 1 ? LILI_KNN1_MINBLOCKS : 2) k_knn_plane(KnnArgs a, const __grid_constant__ PeerArgs pa) {
    __shared__ __align__(16) KnnSmem S;
    if (threadIdx.x == 0) S.peer_lost = 0;
    LILI_STAMP(0);
    const Q4 q{a.pose[0], a.pose[1], a.pose[2], a.pose[3]};
    const D3 t{a.pose[4], a.pose[5], a.pose[6]};
    const int n_q = a.n_dev ? min(*a.n_dev, a.n) : a.n;
    double acc = 0.0;
    unsigned long long cand = 0;
    knn_phases<LANES>(a, q, t, n_q, S, acc, cand, a.use_state != 0);
    LILI_STAMP(3);
    write_block_partials(a, S, acc, cand);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int tk = atomicAdd(a.ticket, 1u);
        S.is_last = (tk == gridDim.x - 1);
    }
    __syncthreads();
    LILI_STAMP(4);
    if (!S.is_last) return;
    __threadfence();
    LILI_STAMP_LAST(5);
    // ---------------- last block: fixed-order sum over blocks, then the 6x6 step
    reduce_partials(a, S);
    if (pa.enabled) peer_exchange(pa, pa.epoch0, S, true);      // multi-GPU, one launch per iteration: the last block trades sums with the peers
    LILI_STAMP_LAST(6);
    if (threadIdx.x == 0) {
        *a.ticket = 0;
        if (a.update_pose) {
            double xn[7];
            gn_step(S, q, t, xn);
#pragma unroll
            for (int k = 0; k < 7; ++k) a.pose_out[k] = xn[k];
            if (pa.enabled) a.pose_out[7] = S.peer_lost ? 1.0 : 0.0;
            if (a.stats) {
#pragma unroll
                for (int k = 0; k < 7; ++k) a.stats[30 + k] = xn[k];
            }
        }
    }
    LILI_STAMP_LAST(7);
    write_neq_stats(a, S, a.stats, a.update_pose != 0);
}

// All GN iterations in ONE cooperative launch (single GPU, GN mode): per iteration the blocks meet at one
// grid barrier after publishing their partials; then EVERY block sums the partials and solves the 6x6
// system redundantly (bit-identical), so no second barrier or broadcast is needed.  Removes the launch
// gap, the drain and the ticket round trip of the per-iteration kernel (~6 us of ~18 per iteration).
// The persistent kernel never runs more than one block per SM (grid <= sm_count), so it takes a 176-register cap instead
// of the 128 of __launch_bounds__(256, 2): the spills (200-600 B per thread in the fp64 fit) disappear, while a
// 64-register block of the Preprocessing node's cooperative kernel still fits beside it (176*256 + 64*256 <= 65536
// registers), which the two-node e2e leg relies on.  Measured on B200 (1.5k-query scans, A/B in one process,
// tools/ab_variants.py): 12.70 -> 12.32 us per pass; with the release-only barrier below 11.04 us.
// (__maxnreg__ and __launch_bounds__ cannot be combined; the block size is fixed by the host code.)
#ifndef LILI_GN_MAXNREG
#define LILI_GN_MAXNREG 176
#endif
#define LILI_GN_BOUNDS __maxnreg__(LILI_GN_MAXNREG)
template <int LANES, bool TMA = false>
__global__ void LILI_GN_BOUNDS k_gn_persistent(KnnArgs a, int iters, unsigned int* bar, double* stats_base, unsigned int bar_base,
                                                             int sync_mode, const __grid_constant__ PeerArgs pa, const __grid_constant__ GnIo io) {
    __shared__ __align__(16) KnnSmem S;
    unsigned stage_phase = 0;
    if constexpr (TMA) {       // one mbarrier per lane group, 16 arrivals (every lane of the group, with or without a run to stage)
        extern __shared__ __align__(16) unsigned char dyn_smem_k[];
        StageTile* tile = reinterpret_cast<StageTile*>(dyn_smem_k) + (threadIdx.x >> 5) * 2 + ((threadIdx.x & 31) >> 4);
        if ((threadIdx.x & 15) == 0) mbar_init(&tile->bar, 16u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const int n_q = a.n_dev ? min(*a.n_dev, a.n) : a.n;
    if (threadIdx.x == 32) S.peer_lost = 0;
    if (threadIdx.x < 7) S.pose[threadIdx.x] = io.use_pose0 ? io.pose0[threadIdx.x] : a.pose[threadIdx.x];
    if (a.dbg && blockIdx.x == 0 && threadIdx.x == 0) a.dbg[30] = (long long)globaltimer_ns();
    __syncthreads();
    const unsigned int G = gridDim.x;
    // bar_base = arrivals of all previous launches on this context (tracked by the host, advanced by iters*G per launch)
    // one task per warp and one round per task (the small-scan shape): each lane group serves the same
    // query in every iteration, so its body-frame point is loaded once and kept in registers
    float4 f_keep = make_float4(0.f, 0.f, 0.f, 0.f);
    bool keep = false;
    {
        constexpr int GROUPS = 32 / LANES;
        const int per_task = GROUPS * a.rounds;
        const int ntasks = (n_q + per_task - 1) / per_task;
        const int gw = a.interleave ? (threadIdx.x >> 5) * (int)gridDim.x + (int)blockIdx.x : (int)blockIdx.x * kWarps + (threadIdx.x >> 5);
        const int nw = gridDim.x * kWarps;
        if (a.rounds == 1 && ntasks <= nw) {
            keep = true;
            const int qi = gw * per_task + ((threadIdx.x & 31) / LANES);
            if (qi < n_q) f_keep = a.feats[qi];
        }
    }
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        const Q4 q{S.pose[0], S.pose[1], S.pose[2], S.pose[3]};
        const D3 t{S.pose[4], S.pose[5], S.pose[6]};
        double acc = 0.0;
        unsigned long long cand = 0;
        const bool stamp = a.dbg && blockIdx.x == 0 && threadIdx.x == 0 && it == 2;
        if (stamp) a.dbg[16] = clock64();
        knn_phases<LANES, TMA>(a, q, t, n_q, S, acc, cand, it > 0, keep ? &f_keep : nullptr, &stage_phase);
        if (stamp) a.dbg[17] = clock64();
        // ---- grid barrier + cross-block sum.  Mode 3 (default; measured 12.70 -> 11.99 us per pass against mode 0): one release by
        // thread 0 after the block barrier (cumulative over bar.sync, as in cooperative groups' grid.sync; SASS: MEMBAR.ALL.GPU +
        // RED, no L1 invalidate) and NO acquire fence after the relaxed poll.  An acquire would only add an L1 invalidation:
        // everything this kernel reads through L1 (map, cell table, features) is immutable for the launch, and the partials are
        // read with L2-scope loads (__ldcg) issued after the poll's control dependency and a bar.sync.  The PTX model formally
        // asks for the acquire: mode 1 polls with ld.acquire (+5 % per pass, profiles/r02_ab_gn_switches.txt), mode 0 fences both
        // sides; tests pin all three to the same bits (test_gpu_variants.py, test_persistent_barrier_stress_all_sync_modes).
        // Multi-GPU (fused exchange): the other blocks of this GPU arrive but do not wait here — they wait for block 0's
        // republished sums over all ranks.
        const bool follower = pa.enabled && blockIdx.x != 0;
        write_block_partials(a, S, acc, cand);
        if (sync_mode == 0) __threadfence();
        __syncthreads();
        if (stamp) a.dbg[18] = clock64();
        if (threadIdx.x == 0) {
            const unsigned int target = bar_base + (unsigned int)(it + 1) * G;
            unsigned int v;
            if (sync_mode == 0) {
                atomicAdd(bar, 1u);
                if (!follower) { while ((int)(*reinterpret_cast<volatile unsigned int*>(bar) - target) < 0) { } __threadfence(); }
            } else {
                asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(bar), "r"(1u) : "memory");
                if (!follower) {
                    if (sync_mode == 1) { do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory"); } while ((int)(v - target) < 0); }
                    else { do { asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory"); } while ((int)(v - target) < 0); }
                }
            }
        }
        if (!follower) {
            __syncthreads();
            if (stamp) a.dbg[19] = clock64();
            reduce_partials(a, S);
            if (pa.enabled) {       // block 0: this rank's sums -> sums over all ranks, then on to the other blocks
                peer_exchange(pa, pa.epoch0 + (unsigned int)it, S, true);
                peer_local_publish(pa, pa.epoch0 + (unsigned int)it, S);
            }
        } else peer_local_wait(pa, pa.epoch0 + (unsigned int)it, S);
        if (stamp) a.dbg[20] = clock64();
        double* stats = stats_base ? stats_base + (size_t)it * kStatsDoubles : nullptr;
        if (threadIdx.x == 0) {
            double xn[7];
            gn_step(S, q, t, xn);
            if (stamp) a.dbg[21] = clock64();
#pragma unroll
            for (int k = 0; k < 7; ++k) S.pose[k] = xn[k];
            if (blockIdx.x == 0) {
#pragma unroll
                for (int k = 0; k < 7; ++k) a.pose_out[k] = xn[k];
                if (pa.enabled) a.pose_out[7] = S.peer_lost ? 1.0 : 0.0;      // explicit loss flag for the host (the pose is NaN as well)
                if (stats) {
#pragma unroll
                    for (int k = 0; k < 7; ++k) stats[30 + k] = xn[k];
                }
            }
        }
        if (blockIdx.x == 0) write_neq_stats(a, S, stats, true);
        if (blockIdx.x == 0 && it == iters - 1 && io.host_out) {      // results straight into the host's pinned block
            if (threadIdx.x == 0) {
#pragma unroll
                for (int k = 0; k < 7; ++k) io.host_out[k] = S.pose[k];
                io.host_out[7] = (pa.enabled && S.peer_lost) ? 1.0 : 0.0;
                if (a.n_dev) *reinterpret_cast<int*>(io.host_out + 40) = *a.n_dev;
            }
            if (threadIdx.x >= 32 && threadIdx.x < 32 + io.n_vgp_words) reinterpret_cast<int*>(io.host_out + 48)[threadIdx.x - 32] = io.vgp[threadIdx.x - 32];
            __threadfence_system();
        }
        __syncthreads();
        // the partials of this iteration may only be overwritten after every block has summed them: the next
        // barrier is behind the next write, so alternate between two partial buffers
        a.partials = (it & 1) ? a.partials - (size_t)kNormEq * G : a.partials + (size_t)kNormEq * G;
    }
    if (a.dbg && blockIdx.x == 0 && threadIdx.x == 0) a.dbg[31] = (long long)globaltimer_ns();
}

// GN step for the multi-GPU path: runs after the all-reduce of neq[29], identically on every rank.
__global__ void k_gn_update(const double* __restrict__ neq, double* __restrict__ pose, double* __restrict__ stats) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double s[kNormEq], x[7], xn[7], nb[6], d[6];
    for (int k = 0; k < kNormEq; ++k) s[k] = neq[k];
    for (int k = 0; k < 7; ++k) x[k] = pose[k];
    for (int k = 0; k < 6; ++k) nb[k] = -s[21 + k];
    if (s[28] > 0.0 && gn_safe_step(s, nb, d)) pose_plus(x, d, xn);
    else for (int k = 0; k < 7; ++k) xn[k] = x[k];
    if (xn[0] < 0) { xn[0] = -xn[0]; xn[1] = -xn[1]; xn[2] = -xn[2]; xn[3] = -xn[3]; }
    for (int k = 0; k < 7; ++k) pose[k] = xn[k];
    if (stats) {
        stats[0] = s[28]; stats[1] = 1.0; stats[2] = s[27];
        for (int k = 0; k < 27; ++k) stats[3 + k] = s[k];
        for (int k = 0; k < 7; ++k) stats[30 + k] = xn[k];
    }
}

// ------------------------------------------------------------------ Ceres-faithful LM on frozen correspondences
// One block: every LM iteration is a pass over the correspondences (cost + 27 scalars at the
// candidate), a block reduction, and the trust-region bookkeeping of Ceres 2.0's
// TrustRegionMinimizer + LevenbergMarquardtStrategy on thread 0 (Jacobi scaling fixed at
// iteration 0, D = sqrt(clamp(diag)/radius), rho-based radius update), with the dense QR on
// [J; D] replaced by its normal equations (J^T J + D^2) y = J^T r in fp64 (6x6 LDL^T).
struct LmArgs {
    const float4* feats; const unsigned char* valid; const float4* plane; int n; const int* n_dev;
    double* pose;         // in/out
    double* stats;        // slot of this outer iteration
    double huber_a;
    int max_num_iter;
    const double* neq0;   // 29 scalars at the linearisation pose (from k_knn_plane)
    int nranks; void* unused;
};

constexpr int kLmBlock = 512;

__device__ void lm_eval(const LmArgs& a, const double x[7], double out[kNormEq], double (*red)[kNormEq]) {
    Q4 q{x[0], x[1], x[2], x[3]};
    double acc[kNormEq];
#pragma unroll
    for (int k = 0; k < kNormEq; ++k) acc[k] = 0.0;
    const int n_c = a.n_dev ? min(*a.n_dev, a.n) : a.n;
    for (int i = threadIdx.x; i < n_c; i += blockDim.x) {
        if (!a.valid[i]) continue;
        float4 f = a.feats[i];
        float4 pl = a.plane[i];
        D3 rp = qrot_x(q, D3{(double)f.x, (double)f.y, (double)f.z});
        double nx = pl.x, ny = pl.y, nz = pl.z;
        double r = nx * (rp.x + x[4]) + ny * (rp.y + x[5]) + nz * (rp.z + x[6]) + (double)pl.w;
        double J[6];
        J[0] = 2.0 * (rp.y * nz - rp.z * ny);
        J[1] = 2.0 * (rp.z * nx - rp.x * nz);
        J[2] = 2.0 * (rp.x * ny - rp.y * nx);
        J[3] = nx; J[4] = ny; J[5] = nz;
        double s2 = r * r, rho0, rho1;
        const double b2 = a.huber_a * a.huber_a;
        if (s2 > b2) { double rt = sqrt(s2); rho0 = 2.0 * a.huber_a * rt - b2; rho1 = fmax(DBL_MIN, a.huber_a / rt); }
        else { rho0 = s2; rho1 = 1.0; }
        double sr = sqrt(rho1);
#pragma unroll
        for (int k = 0; k < 6; ++k) J[k] *= sr;
        r *= sr;
        int k = 0;
#pragma unroll
        for (int i2 = 0; i2 < 6; ++i2)
#pragma unroll
            for (int j = i2; j < 6; ++j) acc[k++] += J[i2] * J[j];
#pragma unroll
        for (int i2 = 0; i2 < 6; ++i2) acc[21 + i2] += J[i2] * r;
        acc[27] += 0.5 * rho0;
        acc[28] += 1.0;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < kNormEq; ++k) {
        double v = warp_sum(acc[k]);
        if (lane == 0) red[warp][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < kNormEq) {
        double v = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += red[w][threadIdx.x];
        red[0][threadIdx.x] = v;
    }
    __syncthreads();
    for (int k = 0; k < kNormEq; ++k) out[k] = red[0][k];
    __syncthreads();
}

__global__ void __launch_bounds__(kLmBlock) k_lm_solve(LmArgs a) {
    __shared__ double red[kLmBlock / 32][kNormEq];
    __shared__ double sh_x[7];
    __shared__ int sh_cmd;   // 0 = stop, 1 = evaluate candidate in sh_x
    double x[7];
    for (int k = 0; k < 7; ++k) x[k] = a.pose[k];
    double S[kNormEq];
    lm_eval(a, x, S, red);          // iteration 0: cost, gradient, J^T J at x
    // --- thread-0 state (replicated arithmetic is avoided: only thread 0 decides) ---
    double radius = 1e4, decrease_factor = 2.0;
    bool reuse_diagonal = false, last_successful = false;
    double diagonal[6], scaling[6];
    double x_cost = S[27];
    int iteration = 0, invalid = 0;
    if (threadIdx.x == 0) {
        int kk = 0;
        for (int i = 0; i < 6; ++i) for (int j = i; j < 6; ++j) { if (i == j) scaling[i] = 1.0 / (1.0 + sqrt(S[kk])); ++kk; }
    }
    const double count0 = S[28];
    double step_keep[6];
    double model_cost_change = 0;
    for (;;) {
        if (threadIdx.x == 0) {
            int cmd = 1;
            double xc[7];
            for (;;) {   // loop over invalid steps without evaluation
                if (count0 <= 0.0) { cmd = 0; break; }
                if (iteration >= a.max_num_iter) { cmd = 0; break; }
                if (last_successful) {
                    double gmax = 0;
                    for (int k = 0; k < 6; ++k) gmax = fmax(gmax, fabs(S[21 + k]));
                    if (gmax <= 1e-10) { cmd = 0; break; }
                }
                if (radius < 1e-32) { cmd = 0; break; }
                ++iteration;
                last_successful = false;
                // scaled normal equations  Hs = S H S, gs = S g
                double Hs[21], gs[6];
                int kk = 0;
                for (int i = 0; i < 6; ++i) for (int j = i; j < 6; ++j) { Hs[kk] = S[kk] * scaling[i] * scaling[j]; ++kk; }
                for (int i = 0; i < 6; ++i) gs[i] = S[21 + i] * scaling[i];
                if (!reuse_diagonal) {
                    kk = 0;
                    for (int i = 0; i < 6; ++i) for (int j = i; j < 6; ++j) { if (i == j) diagonal[i] = fmin(fmax(Hs[kk], 1e-6), 1e32); ++kk; }
                }
                double Ha[21];
                kk = 0;
                for (int i = 0; i < 6; ++i) for (int j = i; j < 6; ++j) { Ha[kk] = Hs[kk] + (i == j ? diagonal[i] / radius : 0.0); ++kk; }
                double y[6];
                bool ok = solve6_ldlt(Ha, gs, y);
                reuse_diagonal = true;
                double step[6];
                model_cost_change = 0;
                if (ok) {
                    for (int k = 0; k < 6; ++k) step[k] = -y[k];
                    // -(step.gs + 1/2 step^T Hs step)
                    double Hfull[6][6];
                    kk = 0;
                    for (int i = 0; i < 6; ++i) for (int j = i; j < 6; ++j) { Hfull[i][j] = Hs[kk]; Hfull[j][i] = Hs[kk]; ++kk; }
                    double sg = 0, shs = 0;
                    for (int i = 0; i < 6; ++i) {
                        sg += step[i] * gs[i];
                        double hv = 0;
                        for (int j = 0; j < 6; ++j) hv += Hfull[i][j] * step[j];
                        shs += step[i] * hv;
                    }
                    model_cost_change = -(sg + 0.5 * shs);
                }
                if (!ok || !(model_cost_change > 0.0)) {
                    if (++invalid >= 5) { cmd = 0; break; }
                    radius *= 0.5; reuse_diagonal = false;
                    continue;
                }
                invalid = 0;
                for (int k = 0; k < 6; ++k) step_keep[k] = step[k] * scaling[k];
                pose_plus(x, step_keep, xc);
                for (int k = 0; k < 7; ++k) sh_x[k] = xc[k];
                cmd = 1;
                break;
            }
            sh_cmd = cmd;
        }
        __syncthreads();
        if (sh_cmd == 0) break;
        double xc[7];
        for (int k = 0; k < 7; ++k) xc[k] = sh_x[k];
        double Sc[kNormEq];
        lm_eval(a, xc, Sc, red);
        int stop = 0;
        if (threadIdx.x == 0) {
            double candidate_cost = Sc[27];
            double xn = 0, sn = 0;
            for (int k = 0; k < 7; ++k) { xn += x[k] * x[k]; sn += (x[k] - xc[k]) * (x[k] - xc[k]); }
            xn = sqrt(xn); sn = sqrt(sn);
            double cost_change = x_cost - candidate_cost;
            if (sn <= 1e-8 * (xn + 1e-8)) stop = 1;                                   // parameter tolerance
            else if (fabs(cost_change) <= 1e-6 * x_cost) stop = 1;                    // function tolerance
            else {
                double relative_decrease = cost_change / model_cost_change;
                if (relative_decrease > 1e-3) {
                    for (int k = 0; k < 7; ++k) x[k] = xc[k];
                    for (int k = 0; k < kNormEq; ++k) S[k] = Sc[k];
                    x_cost = candidate_cost;
                    last_successful = true;
                    double qd = 2.0 * relative_decrease - 1.0;
                    radius = radius / fmax(1.0 / 3.0, 1.0 - qd * qd * qd);
                    radius = fmin(1e16, radius);
                    decrease_factor = 2.0;
                    reuse_diagonal = false;
                } else {
                    radius = radius / decrease_factor;
                    decrease_factor *= 2.0;
                    reuse_diagonal = true;
                }
            }
            sh_cmd = stop ? 0 : 1;
        }
        __syncthreads();
        if (sh_cmd == 0) break;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (x[0] < 0) { x[0] = -x[0]; x[1] = -x[1]; x[2] = -x[2]; x[3] = -x[3]; }     // :539-549
        for (int k = 0; k < 7; ++k) a.pose[k] = x[k];
        if (a.stats) {
            a.stats[0] = a.neq0[28]; a.stats[1] = (double)iteration; a.stats[2] = a.neq0[27];
            for (int k = 0; k < 27; ++k) a.stats[3 + k] = a.neq0[k];
            for (int k = 0; k < 7; ++k) a.stats[30 + k] = x[k];
        }
    }
}

// ------------------------------------------------------------------ instrumentation: the C-bar of SURVEY.md §8(d)
// Points in the full 3x3x3 cell block of every (owned) query at pose (q,t): the candidate count the algorithmic-bytes figure
// B_q = 16 + 27*8 + 16*C-bar is defined on.  The search kernels examine fewer (pruning, knn_core.cuh) and count those
// separately (liliom_counters::knn_candidates).  out[0] += queries, out[1] += block points.
__global__ void k_block27_count(const float4* __restrict__ feats, int n, Q4 q, D3 t, const int* __restrict__ cell_start, GridDesc g,
                                int nranks, int rank, float inv_block, unsigned long long* __restrict__ out) {
    unsigned long long cnt = 0, nq = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 f = feats[i];
        const D3 pw = qrot_x(q, D3{(double)f.x, (double)f.y, (double)f.z});
        const float sx = (float)addx(pw.x, t.x), sy = (float)addx(pw.y, t.y), sz = (float)addx(pw.z, t.z);
        if (nranks > 1 && owner_of(sx, sy, sz, nranks, inv_block) != rank) continue;
        ++nq;
        const int cx = cell_coord(sx, g.inv_cell) - g.org[0], cy = cell_coord(sy, g.inv_cell) - g.org[1], cz = cell_coord(sz, g.inv_cell) - g.org[2];
        const int x0 = max(cx - 1, 0), x1 = min(cx + 1, g.dim[0] - 1);
        if (x0 > x1) continue;
        for (int row = 0; row < 9; ++row) {
            const int y = cy + (row % 3) - 1, z = cz + (row / 3) - 1;
            if (y < 0 || y >= g.dim[1] || z < 0 || z >= g.dim[2]) continue;
            const int base = (z * g.dim[1] + y) * g.dim[0];
            cnt += (unsigned long long)(__ldg(cell_start + base + x1 + 1) - __ldg(cell_start + base + x0));
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { cnt += __shfl_xor_sync(0xffffffffu, cnt, o); nq += __shfl_xor_sync(0xffffffffu, nq, o); }
    if ((threadIdx.x & 31) == 0) { if (nq) atomicAdd(out, nq); if (cnt) atomicAdd(out + 1, cnt); }
}

int block27_stats(liliom_ctx* c, const double pose7[7], unsigned long long out[2]) {
    out[0] = out[1] = 0;
    if (!c->map_ready) return LILIOM_E_NOMAP;
    const int n = c->n_feats_actual;
    if (n <= 0) return LILIOM_OK;
    LILI_CUDA(c, c->lm_state.ensure(64 * sizeof(long long)));
    unsigned long long* d = c->lm_state.as<unsigned long long>();
    LILI_CUDA(c, cudaMemsetAsync(d, 0, 16, c->stream));
    const Q4 q{pose7[0], pose7[1], pose7[2], pose7[3]};
    const D3 t{pose7[4], pose7[5], pose7[6]};
    k_block27_count<<<min(cdiv(n, 256), c->sm_count * 4), 256, 0, c->stream>>>(c->feats.as<float4>(), n, q, t, c->cell_start.as<int>(), c->grid,
                                                                              c->nranks, c->rank, c->shard_inv_block, d);
    LILI_TRY(launch_check(c, "k_block27_count"));
    unsigned long long* hp = reinterpret_cast<unsigned long long*>(c->h_pin) + 56;
    LILI_CUDA(c, cudaMemcpyAsync(hp, d, 16, cudaMemcpyDeviceToHost, c->stream));
    LILI_CUDA(c, cudaStreamSynchronize(c->stream));
    out[0] = hp[0]; out[1] = hp[1];
    return LILIOM_OK;
}

// ------------------------------------------------------------------ host orchestration
typedef int (*nccl_allreduce_fn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
int nccl_allreduce_sum_f64(liliom_ctx* c, double* buf, int count);   // comm.cu

// Work decomposition.  LANES lanes cooperate on one query; a warp finishes (32/LANES)*rounds queries
// per task before the fp64 phase.  Large query sets: one thread per query (every issue slot ranks 32
// candidates, neighbouring threads share cells in L1).  Small sets (a down-sampled 24k-point sweep
// leaves 1-3k queries) are latency bound: split each query's 9 runs over 4 or 8 lanes so that the
// whole GPU is busy and the per-warp critical path is short.
static void pick_shape(int n, int sm_count, int& lanes, int& rounds) {
    const long long w8 = (long long)sm_count * 8;      // warps for ~8 per SM
    rounds = 1;
    if ((long long)n >= w8 * 32) { lanes = 1; return; }
    // 9.5k..38k queries: two lanes per query (measured at 16k queries, profiles/r02_knn_lanes_sweep.txt: 20.8 us per pass with
    // 2 lanes, 22.1 with 1, 27.2 with 4, 33.5 with 8)
    if ((long long)n >= w8 * 8) { lanes = 2; return; }
    lanes = ((long long)n * 16 / 32 <= w8 * 2) ? 16 : 8;   // tiny scans: one run per lane (9 of 16 lanes)
}

int s2m_run(liliom_ctx* c, double pose7[7], int match_cnt, int max_num_iter, int mode, liliom_iter_stats* stats,
            bool want_corr, double out29[29]) {
    if (!c->map_ready) return LILIOM_E_NOMAP;
    if (c->map_n_global < 10) return LILIOM_E_FEWMAP;          // L/src/LidarOdometry.cpp:485-488
    const int n = c->n_feats;
    if (match_cnt < 0) return LILIOM_E_ARG;
    // the per-iteration stats come back through the pinned block: check its capacity before anything is enqueued
    if (stats && match_cnt > 0 && 64 * sizeof(double) + (size_t)match_cnt * kStatsDoubles * sizeof(double) > c->h_pin_bytes) return LILIOM_E_CAPACITY;
    const int iters = match_cnt;
    // kernel timing (liliom_set_kernel_timing): an event pair + the device-side query/candidate counts on every time_every-th call
    const bool timed = c->time_kernels && (c->time_calls++ % (unsigned)c->time_every) == 0;
    const int n_est = c->d_nfeats ? (c->last_n_feats > 0 ? min(c->last_n_feats, n) : n) : n;   // device-side count: predict from the last scan
    int lanes = 8, rounds = 1;
    pick_shape(n_est, c->sm_count, lanes, rounds);
    if (c->force_lanes) { lanes = c->force_lanes; rounds = c->force_rounds > 0 ? c->force_rounds : 1; }
    if ((32 / lanes) * rounds > 32) rounds = lanes;   // at most 32 queries per warp task
    int per_task = (32 / lanes) * rounds;
    int ntasks = cdiv(c->d_nfeats ? max(n_est + n_est / 4, 1) : n, per_task);
    int grid = min(max(cdiv(ntasks, kWarps), 1), c->sm_count * 2);
    // A 16-lane scan that needs more than one block per SM (1.9k-3.8k queries, e.g. a down-sampled HDL-64E sweep) would
    // leave the single-launch path.  Eight lanes per query (four queries per warp, one round) keep it there; measured on the
    // configs[2] workload, 3.1k queries: 13.7 us per pass, against 16.1 for two rounds of 16 lanes and 21 for per-iteration
    // launches (profiles/r02_ab_lanes_3k.txt).
    if (lanes == 16 && rounds == 1 && !c->force_lanes && grid > c->sm_count && mode == LILIOM_MODE_GN && !want_corr) {
        const int ntasks8 = cdiv(c->d_nfeats ? max(n_est + n_est / 4, 1) : n, 4);
        const int grid8 = max(cdiv(ntasks8, kWarps), 1);
        if (grid8 <= c->sm_count) { lanes = 8; per_task = 4; ntasks = ntasks8; grid = grid8; }
    }

    LILI_CUDA(c, c->pose_dev.ensure(16 * sizeof(double)));
    // two buffers (the persistent kernel alternates); sized for the largest grid so that a bigger scan never reallocates mid-stream
    LILI_CUDA(c, c->partials.ensure((size_t)2 * max(grid, c->sm_count * 2) * kNormEq * sizeof(double)));
    LILI_CUDA(c, c->neq.ensure(32 * sizeof(double)));
    LILI_CUDA(c, c->stats_dev.ensure((size_t)(iters + 1) * kStatsDoubles * sizeof(double)));
    if (!c->counter.p) {
        LILI_CUDA(c, c->counter.ensure(64));
        LILI_CUDA(c, cudaMemsetAsync(c->counter.p, 0, 64, c->stream));
    }
    const bool need_corr = want_corr || mode == LILIOM_MODE_CERES;
    if (need_corr) {
        LILI_CUDA(c, c->corr_valid.ensure((size_t)n + 16));
        LILI_CUDA(c, c->corr_plane.ensure((size_t)n * sizeof(float4) + 16));
    }
    if (want_corr) {
        LILI_CUDA(c, c->nn_idx.ensure((size_t)n * 5 * sizeof(int) + 16));
        LILI_CUDA(c, c->nn_sqd.ensure((size_t)n * 5 * sizeof(float) + 16));
        LILI_CUDA(c, cudaMemsetAsync(c->nn_idx.p, 0xff, (size_t)n * 5 * sizeof(int), c->stream));
        LILI_CUDA(c, cudaMemsetAsync(c->nn_sqd.p, 0x7f, (size_t)n * 5 * sizeof(float), c->stream));
    }
    KnnArgs a{};
    a.feats = c->feats.as<float4>(); a.n = n; a.n_dev = c->d_nfeats;
    a.map = c->map_sorted.as<float4>(); a.map_orig = c->map_xyzw.as<float4>(); a.cell_start = c->cell_start.as<int>(); a.g = c->grid;
    a.pose = c->pose_dev.as<double>(); a.pose_out = c->pose_dev.as<double>();
    a.max_sqd = c->prm.knn_max_sqdist; a.plane_thres = c->prm.plane_thres; a.w_gate = c->prm.weight_gate; a.huber_a = c->prm.huber_a;
    a.valid = need_corr ? c->corr_valid.as<unsigned char>() : nullptr;
    a.plane = need_corr ? c->corr_plane.as<float4>() : nullptr;
    a.nn_idx = want_corr ? c->nn_idx.as<int>() : nullptr;
    a.nn_sqd = want_corr ? c->nn_sqd.as<float>() : nullptr;
    a.partials = c->partials.as<double>(); a.neq = c->neq.as<double>();
    a.ticket = c->counter.as<unsigned int>();
    a.cand_total = timed ? reinterpret_cast<unsigned long long*>(c->counter.as<unsigned char>() + 16) : nullptr;
    a.queries_total = timed ? reinterpret_cast<unsigned long long*>(c->counter.as<unsigned char>() + 24) : nullptr;
    a.tau0 = knn_gate_tau(c->prm.knn_max_sqdist);
    LILI_CUDA(c, c->qstate.ensure((size_t)(n > 0 ? n : 1) * sizeof(float4)));
    a.qstate = c->qstate.as<float4>();
    a.use_state = 0;
    a.inv_block = c->shard_inv_block;
    a.rounds = rounds; a.nranks = c->nranks; a.rank = c->rank;
    a.interleave = (lanes >= 8 && !getenv("LILIOM_NO_INTERLEAVE")) ? 1 : 0;
    a.dbg = nullptr;
    if (getenv("LILIOM_DEBUG_TIMING")) {
        LILI_CUDA(c, c->lm_state.ensure(64 * sizeof(long long)));
        LILI_CUDA(c, cudaMemsetAsync(c->lm_state.p, 0, 64 * sizeof(long long), c->stream));
        a.dbg = c->lm_state.as<long long>();
    }

    // one thread per query: per-thread run lists in dynamic shared memory (thread_knn5); static + dynamic exceed 48 KB
    const size_t dyn_smem = lanes == 1 ? (size_t)kRunCap * kBlock * sizeof(int4) : 0;
    if (lanes == 1 && !c->knn1_smem_set) {
        LILI_CUDA(c, cudaFuncSetAttribute((const void*)k_knn_plane<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem));
        LILI_CUDA(c, cudaFuncSetAttribute((const void*)k_gn_persistent<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem));
        c->knn1_smem_set = true;
    }
    // ---- single-GPU GN: all iterations in one cooperative launch
    const bool peer = c->peer_ready && c->peer_ptrs[c->rank] != nullptr;       // fused exchange instead of ncclAllReduce + k_gn_update
    const bool persistent = mode == LILIOM_MODE_GN && (c->nranks == 1 || peer) && iters > 0 && !want_corr &&
                            !getenv("LILIOM_NO_PERSISTENT") && grid <= c->sm_count;
    // zero-copy results (GnIo): the persistent kernel takes the start pose as a parameter and leaves pose, query count and
    // VoxelGrid parameters in the pinned block itself; only per-iteration stats / the 29 sums still travel by copy
    const bool host_io = persistent && c->host_results && c->h_pin_dev != nullptr;
    if (!host_io) {
        double p8[8] = {pose7[0], pose7[1], pose7[2], pose7[3], pose7[4], pose7[5], pose7[6], 0.0};      // [7]: peer-loss flag, cleared
        double* pin = reinterpret_cast<double*>(c->h_pin) + 56;      // pinned staging (slots 56..63 of the small block; results land in 0..55)
        memcpy(pin, p8, sizeof(p8));
        LILI_CUDA(c, cudaMemcpyAsync(c->pose_dev.p, pin, 8 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    }
    if (persistent) {
        a.update_pose = 1;
        a.stats = nullptr;
        int iters_arg = iters;
        unsigned int* bar = reinterpret_cast<unsigned int*>(c->counter.as<unsigned char>() + 32);
        double* stats_base = c->stats_dev.as<double>();
        unsigned int bar_base = c->bar_arrivals;
        int sync_mode = c->gn_sync;    // 3: release-only counter barrier (default) | 0: counter barrier with full fences
        PeerArgs pa{};
        if (peer) {
            for (int p = 0; p < c->nranks; ++p) pa.buf[p] = reinterpret_cast<ulonglong2*>(c->peer_ptrs[p]);
            pa.nranks = c->nranks; pa.rank = c->rank; pa.enabled = 1;
            pa.lsum = c->peer_local.as<double>(); pa.lflag = reinterpret_cast<unsigned int*>(c->peer_local.as<double>() + 64);
            pa.epoch0 = c->peer_epoch + 1u;          // every rank makes the same sequence of calls: same epochs everywhere
            c->peer_epoch += (unsigned int)iters;
        }
        GnIo io{};
        if (host_io) {
            for (int k = 0; k < 7; ++k) io.pose0[k] = pose7[k];
            io.use_pose0 = 1;
            io.host_out = reinterpret_cast<double*>(c->h_pin_dev);
            io.vgp = c->vg_check ? c->vg_params.as<int>() : nullptr;
            io.n_vgp_words = c->vg_check ? (int)(sizeof(VgParams) / sizeof(int)) : 0;
        }
        void* kargs[] = {&a, &iters_arg, &bar, &stats_base, &bar_base, &sync_mode, &pa, &io};
        const bool tma = lanes == 16 && c->knn_tma;
        size_t dyn_p = dyn_smem;
        if (tma) {
            dyn_p = (size_t)kWarps * 2 * sizeof(StageTile);
            if (!c->knn_tma_smem_set) {
                LILI_CUDA(c, cudaFuncSetAttribute((const void*)k_gn_persistent<16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_p));
                c->knn_tma_smem_set = true;
            }
        }
        const void* fn = tma ? (const void*)k_gn_persistent<16, true> : lanes == 16 ? (const void*)k_gn_persistent<16> : lanes == 1 ? (const void*)k_gn_persistent<1>
                       : lanes == 2 ? (const void*)k_gn_persistent<2> : lanes == 4 ? (const void*)k_gn_persistent<4>
                                                                                   : (const void*)k_gn_persistent<8>;
        size_t ev = 0;
        if (timed) {
            if (c->ev_used + 2 > c->ev_pool.size()) {
                for (int k = 0; k < 64; ++k) { cudaEvent_t e; LILI_CUDA(c, cudaEventCreate(&e)); c->ev_pool.push_back(e); }
            }
            ev = c->ev_used; c->ev_used += 2;
            LILI_CUDA(c, cudaEventRecord(c->ev_pool[ev], c->stream));
        }
        LILI_CUDA(c, cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(kBlock), kargs, dyn_p, c->stream));
        LILI_TRY(launch_check(c, "k_gn_persistent"));
        c->bar_arrivals += (unsigned int)iters * (unsigned int)grid;
        if (timed) {
            LILI_CUDA(c, cudaEventRecord(c->ev_pool[ev + 1], c->stream));
            c->ev_pending.push_back({ev, (unsigned long long)n_est * iters});
            c->ev_iters.push_back(iters);
        }
    }
    const int launches = persistent ? 0 : ((iters == 0 && want_corr) ? 1 : iters);
    for (int it = 0; it < launches; ++it) {
        a.stats = c->stats_dev.as<double>() + (size_t)it * kStatsDoubles;
        a.use_state = it > 0 ? 1 : 0;
        const bool multi = c->nranks > 1;
        // fused exchange, one launch per iteration (large scans: grid > one block per SM): the last block trades the sums with the
        // peers and performs the GN step itself — no ncclAllReduce, no update kernel
        const bool peer_iter = peer && multi && mode == LILIOM_MODE_GN && iters > 0;
        PeerArgs pit{};
        if (peer_iter) {
            for (int p = 0; p < c->nranks; ++p) pit.buf[p] = reinterpret_cast<ulonglong2*>(c->peer_ptrs[p]);
            pit.nranks = c->nranks; pit.rank = c->rank; pit.enabled = 1;
            pit.epoch0 = ++c->peer_epoch;
        }
        a.update_pose = (mode == LILIOM_MODE_GN && (!multi || peer_iter) && iters > 0) ? 1 : 0;
        size_t ev = 0;
        if (timed) {
            if (c->ev_used + 2 > c->ev_pool.size()) {
                for (int k = 0; k < 64; ++k) { cudaEvent_t e; LILI_CUDA(c, cudaEventCreate(&e)); c->ev_pool.push_back(e); }
            }
            ev = c->ev_used; c->ev_used += 2;
            LILI_CUDA(c, cudaEventRecord(c->ev_pool[ev], c->stream));
        }
        if (lanes == 16) k_knn_plane<16><<<grid, kBlock, 0, c->stream>>>(a, pit);
        else if (lanes == 1) k_knn_plane<1><<<grid, kBlock, dyn_smem, c->stream>>>(a, pit);
        else if (lanes == 2) k_knn_plane<2><<<grid, kBlock, 0, c->stream>>>(a, pit);
        else if (lanes == 4) k_knn_plane<4><<<grid, kBlock, 0, c->stream>>>(a, pit);
        else k_knn_plane<8><<<grid, kBlock, 0, c->stream>>>(a, pit);
        LILI_TRY(launch_check(c, "k_knn_plane"));
        if (timed) {
            LILI_CUDA(c, cudaEventRecord(c->ev_pool[ev + 1], c->stream));
            c->ev_pending.push_back({ev, (unsigned long long)n_est});
            c->ev_iters.push_back(1);
        }
        if (iters == 0) break;
        if (multi && !peer_iter) LILI_TRY(nccl_allreduce_sum_f64(c, c->neq.as<double>(), kNormEq));
        if (mode == LILIOM_MODE_GN) {
            if (multi && !peer_iter) {
                k_gn_update<<<1, 32, 0, c->stream>>>(c->neq.as<double>(), c->pose_dev.as<double>(), a.stats);
                LILI_TRY(launch_check(c, "k_gn_update"));
            }
        } else {
            LmArgs l{};
            l.feats = a.feats; l.valid = a.valid; l.plane = a.plane; l.n = n; l.n_dev = c->d_nfeats;
            l.pose = c->pose_dev.as<double>(); l.stats = a.stats; l.huber_a = a.huber_a; l.max_num_iter = max_num_iter;
            l.neq0 = c->neq.as<double>();
            if (multi) { c->last_error = "CERES mode is single-GPU (the LM solve is one block); use GN mode with a communicator"; return LILIOM_E_ARG; }
            k_lm_solve<<<1, kLmBlock, 0, c->stream>>>(l);
            LILI_TRY(launch_check(c, "k_lm_solve"));
        }
    }
    // ---- results: pose + stats in one pinned block
    double* hp = reinterpret_cast<double*>(c->h_pin);
    if (!host_io) LILI_CUDA(c, cudaMemcpyAsync(hp, c->pose_dev.p, 8 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));     // pose + peer-loss flag
    if (out29) LILI_CUDA(c, cudaMemcpyAsync(hp + 8, c->neq.p, kNormEq * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (c->d_nfeats && !host_io) LILI_CUDA(c, cudaMemcpyAsync(hp + 40, c->d_nfeats, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    if (c->vg_check && !host_io) LILI_CUDA(c, cudaMemcpyAsync(hp + 48, c->vg_params.p, sizeof(VgParams), cudaMemcpyDeviceToHost, c->stream));
    const bool want_stats = stats && iters > 0;
    if (want_stats) {
        size_t bytes = (size_t)iters * kStatsDoubles * sizeof(double);
        LILI_CUDA(c, cudaMemcpyAsync(hp + 64, c->stats_dev.p, bytes, cudaMemcpyDeviceToHost, c->stream));
    }
    LILI_CUDA(c, cudaStreamSynchronize(c->stream));
    c->n_feats_actual = c->d_nfeats ? min(*reinterpret_cast<int*>(hp + 40), n) : n;
    c->last_n_feats = c->n_feats_actual;
    if (c->vg_check) {
        const VgParams* vp = reinterpret_cast<const VgParams*>(hp + 48);
        c->vg_ncells = vp->overflow ? 0 : (long long)vp->div_b[0] * vp->div_b[1] * vp->div_b[2];
        c->vg_bail = vp->bail != 0;
    }
    if (iters > 0 && peer && hp[7] != 0.0) {
        c->last_error = "fused exchange: a peer rank did not publish its sums within the wait bound (collective call not entered on every rank?)";
        return LILIOM_E_NCCL;
    }
    if (iters > 0) for (int k = 0; k < 7; ++k) pose7[k] = hp[k];
    if (out29) for (int k = 0; k < kNormEq; ++k) out29[k] = hp[8 + k];
    if (want_stats) {
        for (int it = 0; it < iters; ++it) {
            const double* s = hp + 64 + (size_t)it * kStatsDoubles;
            stats[it].n_corr = (int)s[0]; stats[it].lm_iters = (int)s[1]; stats[it].cost = s[2];
            for (int k = 0; k < 27; ++k) stats[it].jtj_jtr[k] = s[3 + k];
            for (int k = 0; k < 7; ++k) stats[it].pose7[k] = s[30 + k];
        }
    }
    if (a.dbg && c->dbg_timing) {       // stage clocks of the two cooperative kernels that ran before this call (resident pipeline)
        long long t[12], gt[6] = {0, 0, 0, 0, 0, 0};
        if (c->hz_ctl.p && cudaMemcpy(t, c->hz_ctl.as<unsigned char>() + 16, sizeof(t), cudaMemcpyDeviceToHost) == cudaSuccess && t[4] > t[0]) {
            gt[0] = t[10]; gt[1] = t[11];
            fprintf(stderr, "[k_hz_coop, cycles, block 0] A keep-flags+barrier %lld, B de-skew/bin+barrier %lld, C patches+barrier %lld (patch 0: window %lld, "
                            "lists %lld, centroid+scatter %lld, eigen %lld, decisions %lld), D emit %lld\n",
                    t[1] - t[0], t[2] - t[1], t[3] - t[2], t[5] - t[2], t[6] - t[5], t[7] - t[6], t[8] - t[7], t[9] - t[8], t[4] - t[3]);
        }
        const long long* vs = vg_coop_stamps(c);
        if (vs && cudaMemcpy(t, vs, 7 * sizeof(long long), cudaMemcpyDeviceToHost) == cudaSuccess && t[4] > t[0]) {
            gt[2] = t[5]; gt[3] = t[6];
            fprintf(stderr, "[k_vg_coop, cycles, block 0] 1 hash insert+barrier %lld, 2 ranks+barrier %lld, 3 group+barrier %lld, 4 centroids %lld\n",
                    t[1] - t[0], t[2] - t[1], t[3] - t[2], t[4] - t[3]);
        }
        if (cudaMemcpy(t, a.dbg + 30, 2 * sizeof(long long), cudaMemcpyDeviceToHost) == cudaSuccess) { gt[4] = t[0]; gt[5] = t[1]; }
        if (gt[0] && gt[2] && gt[4])      // the step's device timeline (block 0 of each kernel, %globaltimer)
            fprintf(stderr, "[step timeline, ns from k_hz_coop start] hz 0..%lld | gap %lld | vg %lld..%lld | gap %lld | gn %lld..%lld (%lld)\n",
                    gt[1] - gt[0], gt[2] - gt[1], gt[2] - gt[0], gt[3] - gt[0], gt[4] - gt[3], gt[4] - gt[0], gt[5] - gt[0], gt[5] - gt[4]);
    }
    if (a.dbg) {
        long long h[24];
        cudaMemcpy(h, a.dbg, sizeof(h), cudaMemcpyDeviceToHost);
        if (h[16]) fprintf(stderr, "[persistent it=2, cycles] phases %lld, partials+fence %lld, barrier %lld, reduce %lld, solve %lld\n",
                           h[17] - h[16], h[18] - h[17], h[19] - h[18], h[20] - h[19], h[21] - h[20]);
        if (!persistent && h[0] && h[1] > h[0])
        fprintf(stderr, "[phase A detail] pose+feat+transform %lld, first row bounds %lld, candidates+rank %lld, merge %lld, slot %lld\n",
                h[8] - h[0], h[9] - h[8], h[10] - h[9], h[11] - h[10], h[1] - h[11]);
        if (!persistent && h[0] && h[4] > h[0])
        fprintf(stderr, "[knn phases, cycles] blk0: start->A %lld, B %lld, C %lld, loop-end %lld, block-reduce+ticket %lld | last block: reduce %lld, solve %lld (abs tail %lld after blk0 start)\n",
                h[1] - h[0], h[2] - h[1], h[3] - h[2], 0LL, h[4] - h[3], h[6] - h[5], h[7] - h[6], h[7] - h[0]);
    }
    // fold finished kernel timings into the counters
    if (timed) {
        for (size_t k = 0; k < c->ev_pending.size(); ++k) {
            auto& pr = c->ev_pending[k];
            float ms = 0;
            if (cudaEventElapsedTime(&ms, c->ev_pool[pr.first], c->ev_pool[pr.first + 1]) == cudaSuccess) {
                // a persistent launch covers `iters` passes of the kernel body: count each pass as one "launch"
                c->cnt.knn_ms += ms; c->cnt.knn_launches += (unsigned long long)c->ev_iters[k];
            }
        }
        c->ev_pending.clear();
        c->ev_iters.clear();
        c->ev_used = 0;
    }
    return LILIOM_OK;
}

}  // namespace lili
