// Livox-Horizon feature extraction on sm_100a — replaces the loops of
// Preprocessing::cloudHandler, L/src/Preprocessing.cpp:225-383:
//   k_hz_flags     removeNaN (:225) + removeClosedPointCloud 0.1 m (:72-97,226) + scan_id>=0 (:253)
//   (scan)         stable compaction index = position in lidar_cloud_cutted
//   k_hz_deskew_bin  undistortion (:104-127) -> cutted cloud (:257); range/reflectivity gates
//                  (:259-261); time column (:262); first-writer-wins cell occupancy (:265-267)
//                  via atomicMin on the point index (the reference loop is sequential)
//   k_hz_patch     one warp per 6x6 patch (664 patches, :270): PCA of the valid cells (:271-298),
//                  per-line depth-Laplacian arg-max (:302-331), edge PCA + gate (:333-365),
//                  planar gate (:367-382).  Patches are independent.
//   k_hz_offsets / k_hz_emit   patch-major ordered compaction into the published clouds.
// Compiled with --fmad=false: every fp32/fp64 expression rounds as in the reference's
// non-FMA x86-64 build, so labels and feature indices are bit-exact against the oracle.
#include "ctx.cuh"
#include "dev_math.cuh"
#include <climits>

namespace lili {

constexpr int HZ_LINES = 6;       // N_SCANS, Preprocessing.cpp:34
constexpr int HZ_COLS = 4000;     // H_SCANS, Preprocessing.cpp:35
constexpr int HZ_PATCHES = 664;   // i = 5 .. 3983 step 6
constexpr int HZ_EMPTY = INT_MAX;

struct Pt48 { float4 a, b, c; };  // {x,y,z,1} {nx,ny,nz,0} {intensity,curvature,0,0}

__global__ void k_hz_flags(const Pt48* __restrict__ pts, int n, int* __restrict__ flags) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    int f = 0;
    if (i < n) {
        float4 a = pts[i].a;
        float inten = pts[i].c.x;
        const float thres = 0.1f;
        bool fin = isfinite(a.x) && isfinite(a.y) && isfinite(a.z);
        bool close = (a.x * a.x + a.y * a.y + a.z * a.z) < thres * thres;
        int scan_id = (int)inten;
        f = fin && !close && scan_id >= 0;
    }
    flags[i] = f;
}

__global__ void k_hz_fill(int* __restrict__ mat, int n, int v) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) mat[i] = v;
}

__global__ void k_hz_deskew_bin(const Pt48* __restrict__ pts, int n, const int* __restrict__ flags, const int* __restrict__ cidx,
                                Q4 q_imu, Pt48* __restrict__ cut, int* __restrict__ mat) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !flags[i]) return;
    const float4 a = pts[i].a;
    const float intensity = pts[i].c.x, curvature = pts[i].c.y;
    const int scan_id = (int)intensity;
    // undistortion, :104-127
    int line = (int)intensity;
    double dt_i = (double)(intensity - (float)line);
    double ratio_i = dt_i / 0.1;
    if (ratio_i >= 1.0) ratio_i = 1.0;
    Q4 q_si = qslerp_x(Q4{1, 0, 0, 0}, ratio_i, q_imu);
    D3 ps = qrot_x(q_si, D3{(double)a.x, (double)a.y, (double)a.z});
    const float ux = (float)ps.x, uy = (float)ps.y, uz = (float)ps.z;
    const int ci = cidx[i];
    Pt48 o;
    o.a = make_float4(ux, uy, uz, 1.0f);
    o.b = make_float4(0.f, 0.f, 0.f, 0.f);
    o.c = make_float4(intensity, curvature, 0.f, 0.f);
    cut[ci] = o;
    double dep = (double)(ux * ux + uy * uy + uz * uz);                                          // :259
    if (dep > 40000.0 || dep < 4.0 || (double)curvature < 0.05 || (double)curvature > 25.45) return;   // :260
    const double t_interval = 0.1 / (HZ_COLS - 1);                                               // :239
    int col = (int)round((double)(intensity - (float)scan_id) / t_interval);                     // :262
    if (col >= HZ_COLS || col < 0) return;
    if (scan_id >= HZ_LINES) return;   // the reference indexes mat[] out of bounds here (UB); guarded
    atomicMin(&mat[scan_id * HZ_COLS + col], ci);                                                // :265-267 first writer wins
}

struct HzCell { float x, y, z, curv, inten, depth; };

constexpr int HZ_WIN = 14;   // columns i-4 .. i+9
constexpr int HZ_WARPS = 4;

__global__ void __launch_bounds__(HZ_WARPS * 32) k_hz_patch(const Pt48* __restrict__ cut, const int* __restrict__ mat,
                                                           double surf_thres, double edge_thres,
                                                           Pt48* __restrict__ stage_surf, Pt48* __restrict__ stage_edge,
                                                           int* __restrict__ counts) {
    __shared__ HzCell win[HZ_WARPS][HZ_LINES * HZ_WIN];
    __shared__ int s_list[HZ_WARPS][36];
    __shared__ int e_list[HZ_WARPS][6];
    __shared__ float nrm[HZ_WARPS][6];   // surf normal (0..2), edge direction (3..5)
    __shared__ int cnts[HZ_WARPS][2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int patch = blockIdx.x * HZ_WARPS + warp;
    if (patch >= HZ_PATCHES) return;
    const int i0 = 5 + 6 * patch;
    HzCell* W = win[warp];
    for (int e = lane; e < HZ_LINES * HZ_WIN; e += 32) {
        int k = e / HZ_WIN, cc = e % HZ_WIN;
        int idx = mat[k * HZ_COLS + (i0 - 4 + cc)];
        HzCell h{0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (idx != HZ_EMPTY) {
            float4 a = cut[idx].a, c = cut[idx].c;
            h.x = a.x; h.y = a.y; h.z = a.z; h.inten = c.x; h.curv = c.y;
            h.depth = sqrtf(a.x * a.x + a.y * a.y + a.z * a.z);                                   // getDepth :99-102
        }
        W[e] = h;
    }
    __syncwarp();
    if (lane == 0) {
        auto C = [&](int k, int j) -> HzCell& { return W[k * HZ_WIN + (j + 4)]; };   // j = column offset from i0 (-4..9)
        int ns = 0, ne = 0;
        double cx = 0, cy = 0, cz = 0;
        int num = 36;
        for (int j = 0; j < 6; ++j)
            for (int k = 0; k < HZ_LINES; ++k) {
                const HzCell& h = C(k, j);
                if (h.curv <= 0) { num--; continue; }                                            // :276-279
                cx += (double)h.x; cy += (double)h.y; cz += (double)h.z;
            }
        if (num >= 25) {                                                                          // :287
            cx /= num; cy /= num; cz /= num;
            double a00 = 0, a01 = 0, a02 = 0, a10 = 0, a11 = 0, a12 = 0, a20 = 0, a21 = 0, a22 = 0;
            for (int j = 0; j < 6; ++j)
                for (int k = 0; k < HZ_LINES; ++k) {
                    const HzCell& h = C(k, j);
                    if (h.curv <= 0) continue;
                    double z0 = (double)h.x - cx, z1 = (double)h.y - cy, z2 = (double)h.z - cz;
                    a00 += z0 * z0; a01 += z0 * z1; a02 += z0 * z2;
                    a10 += z1 * z0; a11 += z1 * z1; a12 += z1 * z2;
                    a20 += z2 * z0; a21 += z2 * z1; a22 += z2 * z2;
                }
            double ev[3], evec[3][3];
            eigen_sym3(a00, a10, a20, a11, a21, a22, ev, evec);                                   // :298
            // per-line depth Laplacian, :302-331
            int idsx[HZ_LINES], idsy[HZ_LINES], nedge = 0;
            for (int k = 0; k < HZ_LINES; ++k) {
                double max_s = 0;
                int idx = 0;
                for (int j = 0; j < 6; ++j) {
                    if (C(k, j).curv <= 0) continue;
                    double g1 = (double)C(k, j - 4).depth + (double)C(k, j - 3).depth + (double)C(k, j - 2).depth + (double)C(k, j - 1).depth -
                                8 * (double)C(k, j).depth + (double)C(k, j + 1).depth + (double)C(k, j + 2).depth +
                                (double)C(k, j + 3).depth + (double)C(k, j + 4).depth;
                    g1 = g1 / (8 * (double)C(k, j).depth + 1e-3);
                    if (g1 > 0.06 && g1 > max_s) { max_s = g1; idx = j; }
                }
                if (max_s != 0) { idsx[nedge] = k; idsy[nedge] = idx; ++nedge; }
            }
            bool flipped[HZ_LINES][6];
            for (int k = 0; k < HZ_LINES; ++k) for (int j = 0; j < 6; ++j) flipped[k][j] = false;
            if (nedge > 0) {                                                                      // :333-365
                double ex = 0, ey = 0, ez = 0;
                for (int m = 0; m < nedge; ++m) { const HzCell& h = C(idsx[m], idsy[m]); ex += (double)h.x; ey += (double)h.y; ez += (double)h.z; }
                ex /= nedge; ey /= nedge; ez /= nedge;
                double b00 = 0, b10 = 0, b20 = 0, b11 = 0, b21 = 0, b22 = 0;
                for (int m = 0; m < nedge; ++m) {
                    const HzCell& h = C(idsx[m], idsy[m]);
                    double z0 = (double)h.x - ex, z1 = (double)h.y - ey, z2 = (double)h.z - ez;
                    b00 += z0 * z0; b10 += z1 * z0; b20 += z2 * z0; b11 += z1 * z1; b21 += z2 * z1; b22 += z2 * z2;
                }
                double eev[3], eevec[3][3];
                eigen_sym3(b00, b10, b20, b11, b21, b22, eev, eevec);                             // :351
                if (eev[2] > edge_thres * eev[1] && nedge > 3) {                                  // :353
                    nrm[warp][3] = (float)eevec[0][2]; nrm[warp][4] = (float)eevec[1][2]; nrm[warp][5] = (float)eevec[2][2];
                    for (int m = 0; m < nedge; ++m) {
                        const HzCell& h = C(idsx[m], idsy[m]);
                        if (h.curv <= 0 && h.inten <= 0) continue;                                // :356
                        e_list[warp][ne++] = idsx[m] * HZ_WIN + (idsy[m] + 4);
                        flipped[idsx[m]][idsy[m]] = true;                                         // :363 curvature *= -1
                    }
                }
            }
            if (ev[0] < surf_thres * ev[1]) {                                                     // :367
                nrm[warp][0] = (float)evec[0][0]; nrm[warp][1] = (float)evec[1][0]; nrm[warp][2] = (float)evec[2][0];
                for (int j = 0; j < 6; ++j)
                    for (int k = 0; k < HZ_LINES; ++k) {
                        if (C(k, j).curv <= 0 || flipped[k][j]) continue;                         // :371
                        s_list[warp][ns++] = k * HZ_WIN + (j + 4);
                    }
            }
        }
        cnts[warp][0] = ns; cnts[warp][1] = ne;
        counts[patch] = ns;
        counts[HZ_PATCHES + 1 + patch] = ne;
    }
    __syncwarp();
    const int ns = cnts[warp][0], ne = cnts[warp][1];
    for (int s = lane; s < ns; s += 32) {
        const HzCell& h = W[s_list[warp][s]];
        Pt48 o;
        o.a = make_float4(h.x, h.y, h.z, 1.0f);
        o.b = make_float4(nrm[warp][0], nrm[warp][1], nrm[warp][2], 0.f);
        o.c = make_float4(h.inten, h.curv, 0.f, 0.f);
        stage_surf[patch * 36 + s] = o;
    }
    if (lane < ne) {
        const HzCell& h = W[e_list[warp][lane]];
        Pt48 o;
        o.a = make_float4(h.x, h.y, h.z, 1.0f);
        o.b = make_float4(nrm[warp][3], nrm[warp][4], nrm[warp][5], 0.f);
        o.c = make_float4(h.inten, h.curv, 0.f, 0.f);
        stage_edge[patch * 6 + lane] = o;
    }
}

// counts layout: [0..663] surf counts, [664] spare, [665..1328] edge counts, [1329] spare.
// offs layout  : same, exclusive prefix; totals -> totals[0] (surf), totals[1] (edge).
__global__ void __launch_bounds__(1024) k_hz_offsets(const int* __restrict__ counts, int* __restrict__ offs, int* __restrict__ totals) {
    __shared__ int wsum[32];
    for (int part = 0; part < 2; ++part) {
        const int base = part * (HZ_PATCHES + 1);
        int t = threadIdx.x;
        int v = (t < HZ_PATCHES) ? counts[base + t] : 0;
        int lane = t & 31, warp = t >> 5;
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
        if (lane == 31) wsum[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            int w = wsum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { int u = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += u; }
            wsum[lane] = w;
        }
        __syncthreads();
        int excl = inc - v + (warp > 0 ? wsum[warp - 1] : 0);
        if (t < HZ_PATCHES) offs[base + t] = excl;
        if (t == HZ_PATCHES - 1) totals[part] = excl + v;
        __syncthreads();
    }
}

__global__ void k_hz_emit(const Pt48* __restrict__ stage_surf, const Pt48* __restrict__ stage_edge, const int* __restrict__ counts,
                          const int* __restrict__ offs, Pt48* __restrict__ surf, Pt48* __restrict__ edge) {
    const int patch = blockIdx.x;
    const int t = threadIdx.x;   // 64 threads: 0..35 surf slots, 36..41 edge slots
    if (t < 36) {
        if (t < counts[patch]) surf[offs[patch] + t] = stage_surf[patch * 36 + t];
    } else if (t < 42) {
        int e = t - 36;
        if (e < counts[HZ_PATCHES + 1 + patch]) edge[offs[HZ_PATCHES + 1 + patch] + e] = stage_edge[patch * 6 + e];
    }
}

// raw points must already be in c->raw (n x 48 B).  Leaves cut/surf/edge on the device and the
// three counts in pinned host memory (returned through the pointers after a stream sync).
int horizon_extract_dev(liliom_ctx* c, int n, const double q_imu[4], int* n_surf, int* n_edge, int* n_cut, bool sync_counts) {
    const size_t npts = (size_t)(n > 0 ? n : 1);
    LILI_CUDA(c, c->cut.ensure(npts * sizeof(Pt48)));
    LILI_CUDA(c, c->surf.ensure((size_t)HZ_PATCHES * 36 * sizeof(Pt48)));
    LILI_CUDA(c, c->edge.ensure((size_t)HZ_PATCHES * 6 * sizeof(Pt48)));
    LILI_CUDA(c, c->flags.ensure((npts + 2) * sizeof(int)));
    LILI_CUDA(c, c->idx_a.ensure((npts + 2) * sizeof(int)));
    LILI_CUDA(c, c->hz_mat.ensure((size_t)HZ_LINES * HZ_COLS * sizeof(int)));
    LILI_CUDA(c, c->hz_stage_surf.ensure((size_t)HZ_PATCHES * 36 * sizeof(Pt48)));
    LILI_CUDA(c, c->hz_stage_edge.ensure((size_t)HZ_PATCHES * 6 * sizeof(Pt48)));
    LILI_CUDA(c, c->hz_counts.ensure((size_t)(4 * (HZ_PATCHES + 1) + 8) * sizeof(int)));
    Q4 q{q_imu[0], q_imu[1], q_imu[2], q_imu[3]};
    if (std::isnan(q.w) || std::isnan(q.x) || std::isnan(q.y) || std::isnan(q.z)) q = Q4{1, 0, 0, 0};   // :232-234
    const Pt48* raw = c->raw.as<Pt48>();
    int* flags = c->flags.as<int>();
    int* cidx = c->idx_a.as<int>();
    int* mat = c->hz_mat.as<int>();
    int* counts = c->hz_counts.as<int>();
    int* offs = counts + 2 * (HZ_PATCHES + 1);
    int* totals = counts + 4 * (HZ_PATCHES + 1);
    k_hz_flags<<<cdiv(n + 1, 256), 256, 0, c->stream>>>(raw, n, flags);
    LILI_TRY(launch_check(c, "k_hz_flags"));
    LILI_TRY(exclusive_scan_i32(c, flags, cidx, n));
    k_hz_fill<<<cdiv(HZ_LINES * HZ_COLS, 256), 256, 0, c->stream>>>(mat, HZ_LINES * HZ_COLS, HZ_EMPTY);
    LILI_TRY(launch_check(c, "k_hz_fill"));
    if (n > 0) {
        k_hz_deskew_bin<<<cdiv(n, 128), 128, 0, c->stream>>>(raw, n, flags, cidx, q, c->cut.as<Pt48>(), mat);
        LILI_TRY(launch_check(c, "k_hz_deskew_bin"));
    }
    k_hz_patch<<<cdiv(HZ_PATCHES, HZ_WARPS), HZ_WARPS * 32, 0, c->stream>>>(c->cut.as<Pt48>(), mat, c->prm.surf_thres, c->prm.edge_thres,
                                                                            c->hz_stage_surf.as<Pt48>(), c->hz_stage_edge.as<Pt48>(), counts);
    LILI_TRY(launch_check(c, "k_hz_patch"));
    k_hz_offsets<<<1, 1024, 0, c->stream>>>(counts, offs, totals);
    LILI_TRY(launch_check(c, "k_hz_offsets"));
    k_hz_emit<<<HZ_PATCHES, 64, 0, c->stream>>>(c->hz_stage_surf.as<Pt48>(), c->hz_stage_edge.as<Pt48>(), counts, offs,
                                                c->surf.as<Pt48>(), c->edge.as<Pt48>());
    LILI_TRY(launch_check(c, "k_hz_emit"));
    c->d_nsurf = totals;
    c->n_surf_max = min(n, HZ_PATCHES * 36);
    if (!sync_counts) {          // resident pipeline: the counts stay on the device, no host round trip
        c->n_surf_dev = -1;
        *n_surf = *n_edge = *n_cut = -1;
        return LILIOM_OK;
    }
    int* hp = reinterpret_cast<int*>(c->h_pin);
    LILI_CUDA(c, cudaMemcpyAsync(hp, totals, 2 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    LILI_CUDA(c, cudaMemcpyAsync(hp + 2, cidx + n, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    LILI_CUDA(c, cudaStreamSynchronize(c->stream));
    *n_surf = hp[0]; *n_edge = hp[1]; *n_cut = hp[2];
    c->n_surf_dev = hp[0];
    return LILIOM_OK;
}

}  // namespace lili
