// BackendFusion correspondence searches on the same kNN core:
//   point-to-line : L/src/BackendFusion.cpp:1531-1599 (variant 0), R/src/BackendFusion.cpp:1394-1462 (variant 1)
//   point-to-plane: R/src/BackendFusion.cpp:1464-1520 (configurable radius / plane gate / weight gate / score)
// The sliding-window optimiser that consumes them is out of scope; these kernels only produce
// the per-feature residual inputs (line end points a/b, weighted plane, score).
#include "ctx.cuh"
#include "dev_math.cuh"
#include "knn_core.cuh"

namespace lili {

constexpr int kLanes = 8;

struct BkArgs {
    const float4* feats; int n;
    const float4* map; const float4* map_orig; const int* cell_start; GridDesc g;
    Q4 q; D3 t;
    int variant;
    double max_sqd, plane_thres, w_gate, lidar_const;
    float tau0;        // largest fp32 distance inside the gate (search pruning, knn_core.cuh)
    unsigned char* valid; float* pa; float* pb; float4* plane; double* score;
    const float* map_refl; const float* feat_refl; double reflect_thres;   // Horizon backend variant (L:1617-1638), nullptr = ROT variant
};

__global__ void __launch_bounds__(kBlock) k_backend_edge(BkArgs a) {
    const int lane = threadIdx.x & 31;
    const int sub = lane & (kLanes - 1), oct = lane / kLanes;
    const unsigned omask = ((1u << kLanes) - 1u) << (oct * kLanes);
    const int qi = (blockIdx.x * blockDim.x + threadIdx.x) / kLanes;
    if (qi >= a.n) return;   // uniform per octet
    float4 f = a.feats[qi];
    D3 pw = qrot_x(a.q, D3{(double)f.x, (double)f.y, (double)f.z});
    const float sx = (float)addx(pw.x, a.t.x), sy = (float)addx(pw.y, a.t.y), sz = (float)addx(pw.z, a.t.z);
    Top5 top;
    top5_init(top);
    unsigned long long cand = 0;
    group_knn5<kLanes>(sx, sy, sz, a.map, a.cell_start, a.g, sub, omask, a.tau0, top, cand);
    if (sub != 0) return;
    bool ok = false;
    float A3[3] = {0, 0, 0}, B3[3] = {0, 0, 0};
    if (top.k4 != ~0ull && (double)top5_dist(top.k4) < 1.0) {                                        // L:1543
        const int pos[5] = {top5_index(top.k0), top5_index(top.k1), top5_index(top.k2), top5_index(top.k3), top5_index(top.k4)};
        double px[5], py[5], pz[5], cx = 0, cy = 0, cz = 0;
        for (int j = 0; j < 5; ++j) {
            float4 m = a.map_orig[pos[j]];
            px[j] = m.x; py[j] = m.y; pz[j] = m.z;
            cx = addx(cx, px[j]); cy = addx(cy, py[j]); cz = addx(cz, pz[j]);
        }
        cx = cx / 5.0; cy = cy / 5.0; cz = cz / 5.0;                                  // L:1555
        double a00 = 0, a10 = 0, a20 = 0, a11 = 0, a21 = 0, a22 = 0;
        for (int j = 0; j < 5; ++j) {                                                  // L:1560-1564
            double z0 = subx(px[j], cx), z1 = subx(py[j], cy), z2 = subx(pz[j], cz);
            a00 = addx(a00, mulx(z0, z0)); a10 = addx(a10, mulx(z1, z0)); a20 = addx(a20, mulx(z2, z0));
            a11 = addx(a11, mulx(z1, z1)); a21 = addx(a21, mulx(z2, z1)); a22 = addx(a22, mulx(z2, z2));
        }
        double ev[3], evec[3][3];
        eigen_sym3(a00, a10, a20, a11, a21, a22, ev, evec);                            // L:1568
        if (ev[2] > 3 * ev[1]) {                                                       // L:1575
            const double ux = evec[0][2], uy = evec[1][2], uz = evec[2][2];
            const double ax = cx + 0.1 * ux, ay = cy + 0.1 * uy, az = cz + 0.1 * uz;   // L:1579
            const double bx = cx - 0.1 * ux, by = cy - 0.1 * uy, bz = cz - 0.1 * uz;   // L:1580
            ok = true;
            if (a.variant == 1) {                                                      // R:1435-1439
                const double ux_ = sx - ax, uy_ = sy - ay, uz_ = sz - az;
                const double vx_ = sx - bx, vy_ = sy - by, vz_ = sz - bz;
                const double nx = uy_ * vz_ - uz_ * vy_, ny = uz_ * vx_ - ux_ * vz_, nz = ux_ * vy_ - uy_ * vx_;
                const double dx = ax - bx, dy = ay - by, dz = az - bz;
                const double dist = sqrt(nx * nx + ny * ny + nz * nz) / sqrt(dx * dx + dy * dy + dz * dz);
                if (!(dist < 0.1)) ok = false;
            }
            if (ok) {
                A3[0] = (float)ax; A3[1] = (float)ay; A3[2] = (float)az;
                B3[0] = (float)bx; B3[1] = (float)by; B3[2] = (float)bz;
            }
        }
    }
    a.valid[qi] = ok ? 1 : 0;
    for (int k = 0; k < 3; ++k) { a.pa[3 * (size_t)qi + k] = A3[k]; a.pb[3 * (size_t)qi + k] = B3[k]; }
}

__global__ void __launch_bounds__(kBlock) k_backend_surf(BkArgs a) {
    const int lane = threadIdx.x & 31;
    const int sub = lane & (kLanes - 1), oct = lane / kLanes;
    const unsigned omask = ((1u << kLanes) - 1u) << (oct * kLanes);
    const int qi = (blockIdx.x * blockDim.x + threadIdx.x) / kLanes;
    if (qi >= a.n) return;
    float4 f = a.feats[qi];
    D3 pw = qrot_x(a.q, D3{(double)f.x, (double)f.y, (double)f.z});
    const float sx = (float)addx(pw.x, a.t.x), sy = (float)addx(pw.y, a.t.y), sz = (float)addx(pw.z, a.t.z);
    Top5 top;
    top5_init(top);
    unsigned long long cand = 0;
    group_knn5<kLanes>(sx, sy, sz, a.map, a.cell_start, a.g, sub, omask, a.tau0, top, cand);
    if (sub != 0) return;
    bool ok = false;
    float4 pl = make_float4(0, 0, 0, 0);
    double sc = 0;
    if (top.k4 != ~0ull && (double)top5_dist(top.k4) < a.max_sqd) {                                   // R:1476
        const int pos[5] = {top5_index(top.k0), top5_index(top.k1), top5_index(top.k2), top5_index(top.k3), top5_index(top.k4)};
        double A[5][3], B[5];
        float4 m[5];
        for (int j = 0; j < 5; ++j) { m[j] = a.map_orig[pos[j]]; A[j][0] = m[j].x; A[j][1] = m[j].y; A[j][2] = m[j].z; B[j] = -1.0; }
        double sum_w = 0.0;
        bool skip = false;
        if (a.map_refl) {                                                              // L:1617-1638: rows weighted by 1/|d reflectivity| / sum
            double vw[5];
            const float fr = a.feat_refl[qi];
            for (int j = 0; j < 5; ++j) {
                const double tmp_w = (double)fabsf(fsubx(fr, a.map_refl[pos[j]]));
                sum_w = addx(sum_w, tmp_w);
                vw[j] = 1.0 / tmp_w;          // +inf when the reflectivities coincide, as in the reference
            }
            if (sum_w > a.reflect_thres) skip = true;                                  // L:1628
            for (int j = 0; j < 5; ++j) {
                const double w = vw[j] / sum_w;
                A[j][0] = mulx(w, (double)m[j].x); A[j][1] = mulx(w, (double)m[j].y); A[j][2] = mulx(w, (double)m[j].z);
                B[j] = mulx(B[j], w);
            }
        }
        double nv[3] = {0, 0, 0};
        if (!skip) colpiv_qr_solve_5x3(A, B, nv);                                      // R:1484 / L:1641
        double n2 = nv[0] * nv[0] + nv[1] * nv[1] + nv[2] * nv[2];
        double nn = sqrt(n2);
        double normInverse = 1.0 / nn;
        if (n2 > 0) { nv[0] /= nn; nv[1] /= nn; nv[2] /= nn; }
        bool planeValid = true;
        for (int j = 0; j < 5; ++j)                                                    // R:1489-1497
            if (fabs(nv[0] * m[j].x + nv[1] * m[j].y + nv[2] * m[j].z + normInverse) > a.plane_thres) planeValid = false;
        if (skip) planeValid = false;
        if (planeValid) {
            float pd = (float)addx(addx(addx(mulx(nv[0], (double)sx), mulx(nv[1], (double)sy)), mulx(nv[2], (double)sz)), normInverse);
            float rng = __fsqrt_rn(__fsqrt_rn(faddx(faddx(fmulx(sx, sx), fmulx(sy, sy)), fmulx(sz, sz))));
            float weight = (float)subx(1.0, mulx(0.9, (double)fabsf(pd)) / (double)rng);   // R:1502
            if ((double)weight > a.w_gate) {                                           // R:1504
                pl = make_float4((float)mulx((double)weight, nv[0]), (float)mulx((double)weight, nv[1]),
                                 (float)mulx((double)weight, nv[2]), (float)mulx((double)weight, normInverse));
                sc = a.map_refl ? a.lidar_const * ((double)weight + exp(-sum_w))          // L:1676
                                : a.lidar_const * (double)weight;                             // R:1515
                ok = true;
            }
        }
    }
    a.valid[qi] = ok ? 1 : 0;
    a.plane[qi] = pl;
    a.score[qi] = sc;
}

// curvature (= 0.1 * reflectivity, L/src/FormatConvert.cpp:21) of 48-byte points
__global__ void k_gather_refl(const unsigned char* __restrict__ pts, int n, float* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = reinterpret_cast<const float*>(pts + (size_t)i * 48)[9];
}

static int backend_prepare(liliom_ctx* c, const void* feats, int n, int stride) {
    if (!c->map_ready) return LILIOM_E_NOMAP;
    if (n < 0 || (n > 0 && !feats)) return LILIOM_E_ARG;
    if (stride != 16 && stride != 32 && stride != 48) return LILIOM_E_ARG;
    c->n_feats = n;
    c->d_nfeats = nullptr;
    c->bk_kind = 0; c->bk_n = 0;
    if (n == 0) return LILIOM_OK;
    LILI_CUDA(c, c->feats.ensure((size_t)n * sizeof(float4)));
    if (stride == 16) {
        LILI_CUDA(c, cudaMemcpyAsync(c->feats.p, feats, (size_t)n * 16, cudaMemcpyHostToDevice, c->stream));
    } else {
        LILI_CUDA(c, c->raw.ensure((size_t)n * stride));
        LILI_CUDA(c, cudaMemcpyAsync(c->raw.p, feats, (size_t)n * stride, cudaMemcpyHostToDevice, c->stream));
        LILI_TRY(repack_to_f4(c, c->raw.p, n, stride, c->feats.as<float4>()));
    }
    return LILIOM_OK;
}


// ---------------------------------------------------------------------------------------
// the LiDAR residual blocks of one window keyframe (L/src/BackendFusion.cpp:919-979) reduced on the
// device to that keyframe's 6x6 normal-equation block.  Rows are the closed forms of SURVEY.md Appendix A:
//   LidarEdgeFactor      (LidarKeyframeFactor.h:12-62):  r = s |u x v| / |a-b|, u = p_w-a, v = p_w-b, p_w = q p + t
//                                                        g = dr/dp_w = s ((a-b) x c^) / |a-b|
//   LidarPlaneNormFactor (LidarKeyframeFactor.h:65-108): r = score (n~ . p_w + d~), p_w = q (q_lb^-1 (p - t_lb)) + t, g = score n~
//   row = sqrt(rho') [ g^T , 2 (R p' x g)^T ]   (parameter blocks t, q -> tangent order [t, rot]),  CauchyLoss(b) (:845)
// Works on the correspondences the last liliom_correspond_* call left resident in this context.
// ---------------------------------------------------------------------------------------
struct BlkArgs {
    const float4* feats; int n;
    const unsigned char* valid;
    const float* pa; const float* pb;          // edge
    const float4* plane; const double* score;  // surf
    Q4 q; D3 t; Q4 qlb_inv; D3 tlb;
    double s_weight, cauchy_b;
    double* partials;                          // [gridDim.x][29]
};

constexpr int kBlkThreads = 256;

template <bool EDGE>
__global__ void __launch_bounds__(kBlkThreads) k_backend_block(BlkArgs a) {
    __shared__ double red[kBlkThreads / 32][kNormEq];
    double acc[kNormEq];
#pragma unroll
    for (int k = 0; k < kNormEq; ++k) acc[k] = 0.0;
    const double b2 = a.cauchy_b * a.cauchy_b, c2 = 1.0 / b2;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += gridDim.x * blockDim.x) {
        if (!a.valid[i]) continue;
        const float4 f = a.feats[i];
        D3 pr, g;        // the point the keyframe rotation acts on, and dr/dp_w
        double r;
        if (EDGE) {
            pr = D3{(double)f.x, (double)f.y, (double)f.z};
            const D3 rp = qrot_x(a.q, pr);
            const D3 lp{rp.x + a.t.x, rp.y + a.t.y, rp.z + a.t.z};
            const D3 A{(double)a.pa[3 * (size_t)i], (double)a.pa[3 * (size_t)i + 1], (double)a.pa[3 * (size_t)i + 2]};
            const D3 B{(double)a.pb[3 * (size_t)i], (double)a.pb[3 * (size_t)i + 1], (double)a.pb[3 * (size_t)i + 2]};
            const D3 u{lp.x - A.x, lp.y - A.y, lp.z - A.z}, v{lp.x - B.x, lp.y - B.y, lp.z - B.z};
            const D3 nu = cross_x(u, v);
            const D3 de{A.x - B.x, A.y - B.y, A.z - B.z};
            const double nn = sqrt(nu.x * nu.x + nu.y * nu.y + nu.z * nu.z);
            const double dn = sqrt(de.x * de.x + de.y * de.y + de.z * de.z);
            r = nn / dn * a.s_weight;
            g = D3{0, 0, 0};
            if (nn > 0) {    // a zero cross product has no derivative (the reference's autodiff yields NaN there)
                const D3 cx = cross_x(de, nu);
                const double k = a.s_weight / (dn * nn);
                g = D3{k * cx.x, k * cx.y, k * cx.z};
            }
        } else {
            const D3 d{(double)f.x - a.tlb.x, (double)f.y - a.tlb.y, (double)f.z - a.tlb.z};
            pr = qrot_x(a.qlb_inv, d);
            const D3 rp = qrot_x(a.q, pr);
            const float4 pl = a.plane[i];
            const double sc = a.score[i];
            r = sc * ((double)pl.x * (rp.x + a.t.x) + (double)pl.y * (rp.y + a.t.y) + (double)pl.z * (rp.z + a.t.z) + (double)pl.w);
            g = D3{sc * (double)pl.x, sc * (double)pl.y, sc * (double)pl.z};
        }
        const D3 rp = qrot_x(a.q, pr);
        double J[6];
        J[0] = g.x; J[1] = g.y; J[2] = g.z;
        J[3] = 2.0 * (rp.y * g.z - rp.z * g.y);
        J[4] = 2.0 * (rp.z * g.x - rp.x * g.z);
        J[5] = 2.0 * (rp.x * g.y - rp.y * g.x);
        // ceres::CauchyLoss(b) + Corrector (rho'' < 0 branch)
        const double sum = 1.0 + r * r * c2;
        const double rho0 = b2 * log(sum);
        const double sr = sqrt(fmax(DBL_MIN, 1.0 / sum));
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
        double v = acc[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[warp][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < kNormEq) {
        double v = 0;
#pragma unroll
        for (int w = 0; w < kBlkThreads / 32; ++w) v += red[w][threadIdx.x];
        a.partials[(size_t)blockIdx.x * kNormEq + threadIdx.x] = v;
    }
}

// fixed-order sum over the block partials (run-to-run deterministic)
__global__ void k_backend_block_sum(const double* __restrict__ partials, int nblocks, double* __restrict__ out29) {
    if (threadIdx.x >= kNormEq) return;
    double v = 0;
    for (int b = 0; b < nblocks; ++b) v += partials[(size_t)b * kNormEq + threadIdx.x];
    out29[threadIdx.x] = v;
}

static int backend_block_run(liliom_ctx* c, bool edge, BlkArgs& a, double out29[29]) {
    const int n = c->bk_n;
    const int nblocks = n > 0 ? min(cdiv(n, kBlkThreads), 64) : 1;
    LILI_CUDA(c, c->partials.ensure((size_t)64 * kNormEq * sizeof(double)));
    LILI_CUDA(c, c->neq.ensure(32 * sizeof(double)));
    a.feats = c->feats.as<float4>(); a.n = n; a.valid = c->corr_valid.as<unsigned char>();
    a.partials = c->partials.as<double>();
    if (edge) k_backend_block<true><<<nblocks, kBlkThreads, 0, c->stream>>>(a);
    else k_backend_block<false><<<nblocks, kBlkThreads, 0, c->stream>>>(a);
    LILI_TRY(launch_check(c, "k_backend_block"));
    k_backend_block_sum<<<1, 32, 0, c->stream>>>(a.partials, nblocks, c->neq.as<double>());
    LILI_TRY(launch_check(c, "k_backend_block_sum"));
    double* hp = reinterpret_cast<double*>(c->h_pin);
    LILI_CUDA(c, cudaMemcpyAsync(hp, c->neq.p, kNormEq * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    LILI_CUDA(c, cudaStreamSynchronize(c->stream));
    for (int k = 0; k < kNormEq; ++k) out29[k] = hp[k];
    return LILIOM_OK;
}

}  // namespace lili

using namespace lili;

extern "C" int liliom_correspond_edge(liliom_ctx* c, const void* feats, int n, int stride, const double pose7[7], int variant,
                                      unsigned char* valid, float* pa, float* pb) {
    if (!c || !pose7 || !valid || !pa || !pb || (variant != 0 && variant != 1)) return LILIOM_E_ARG;
    LILI_CUDA(c, cudaSetDevice(c->device));
    LILI_TRY(backend_prepare(c, feats, n, stride));
    if (n == 0) { c->bk_kind = 1; return LILIOM_OK; }
    LILI_CUDA(c, c->corr_valid.ensure((size_t)n + 16));
    LILI_CUDA(c, c->corr_plane.ensure((size_t)n * 24 + 16));
    BkArgs a{};
    a.feats = c->feats.as<float4>(); a.n = n; a.map = c->map_sorted.as<float4>(); a.map_orig = c->map_xyzw.as<float4>(); a.cell_start = c->cell_start.as<int>(); a.g = c->grid;
    a.q = Q4{pose7[0], pose7[1], pose7[2], pose7[3]}; a.t = D3{pose7[4], pose7[5], pose7[6]};
    a.variant = variant;
    a.tau0 = knn_gate_tau(1.0);                                                   // L:1543 sqdist[4] < 1.0
    a.valid = c->corr_valid.as<unsigned char>();
    a.pa = c->corr_plane.as<float>(); a.pb = a.pa + 3 * (size_t)n;
    k_backend_edge<<<cdiv((long long)n * kLanes, kBlock), kBlock, 0, c->stream>>>(a);
    LILI_TRY(launch_check(c, "k_backend_edge"));
    c->bk_kind = 1; c->bk_n = n;
    LILI_CUDA(c, cudaMemcpyAsync(valid, a.valid, (size_t)n, cudaMemcpyDeviceToHost, c->stream));
    LILI_CUDA(c, cudaMemcpyAsync(pa, a.pa, (size_t)n * 12, cudaMemcpyDeviceToHost, c->stream));
    LILI_CUDA(c, cudaMemcpyAsync(pb, a.pb, (size_t)n * 12, cudaMemcpyDeviceToHost, c->stream));
    LILI_CUDA(c, cudaStreamSynchronize(c->stream));
    return LILIOM_OK;
}

static int correspond_surf_impl(liliom_ctx* c, const void* feats, int n, int stride, const double pose7[7], double kd_max_radius,
                                double surf_dist_thres, double w_gate, double lidar_const, bool refl, double reflect_thres,
                                unsigned char* valid, float* plane, double* score) {
    if (!c || !pose7 || !valid || !plane || !score) return LILIOM_E_ARG;
    LILI_CUDA(c, cudaSetDevice(c->device));
    {   // the grid's cell size bounds the exact search radius
        float cell = 1.0f / c->grid.inv_cell;
        if (c->map_ready && c->map_n > 0 && kd_max_radius > (double)cell * (double)cell) {
            c->last_error = "kd_max_radius exceeds the cell size the map grid was built with (raise params.knn_max_sqdist)";
            return LILIOM_E_ARG;
        }
    }
    LILI_TRY(backend_prepare(c, feats, n, stride));
    if (n == 0) { c->bk_kind = 2; return LILIOM_OK; }
    LILI_CUDA(c, c->corr_valid.ensure((size_t)n + 16));
    LILI_CUDA(c, c->corr_plane.ensure((size_t)n * 16 + 16));
    LILI_CUDA(c, c->nn_sqd.ensure((size_t)n * 8 + 16));
    BkArgs a{};
    a.feats = c->feats.as<float4>(); a.n = n; a.map = c->map_sorted.as<float4>(); a.map_orig = c->map_xyzw.as<float4>(); a.cell_start = c->cell_start.as<int>(); a.g = c->grid;
    a.q = Q4{pose7[0], pose7[1], pose7[2], pose7[3]}; a.t = D3{pose7[4], pose7[5], pose7[6]};
    a.tau0 = knn_gate_tau(kd_max_radius);
    a.max_sqd = kd_max_radius; a.plane_thres = surf_dist_thres; a.w_gate = w_gate; a.lidar_const = lidar_const;
    if (refl) {
        if (stride != 48 || !c->map_refl.p) { c->last_error = "reflectivity variant needs 48-byte features and a map installed with liliom_map_set_cloud"; return LILIOM_E_ARG; }
        LILI_CUDA(c, c->nn_idx.ensure((size_t)n * 4 + 16));
        k_gather_refl<<<cdiv(n, 256), 256, 0, c->stream>>>((const unsigned char*)c->raw.p, n, c->nn_idx.as<float>());
        LILI_TRY(launch_check(c, "k_gather_refl"));
        a.map_refl = c->map_refl.as<float>(); a.feat_refl = c->nn_idx.as<float>(); a.reflect_thres = reflect_thres;
    }
    a.valid = c->corr_valid.as<unsigned char>(); a.plane = c->corr_plane.as<float4>(); a.score = c->nn_sqd.as<double>();
    k_backend_surf<<<cdiv((long long)n * kLanes, kBlock), kBlock, 0, c->stream>>>(a);
    LILI_TRY(launch_check(c, "k_backend_surf"));
    c->bk_kind = 2; c->bk_n = n;
    LILI_CUDA(c, cudaMemcpyAsync(valid, a.valid, (size_t)n, cudaMemcpyDeviceToHost, c->stream));
    LILI_CUDA(c, cudaMemcpyAsync(plane, a.plane, (size_t)n * 16, cudaMemcpyDeviceToHost, c->stream));
    LILI_CUDA(c, cudaMemcpyAsync(score, a.score, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
    LILI_CUDA(c, cudaStreamSynchronize(c->stream));
    return LILIOM_OK;
}

extern "C" int liliom_correspond_surf(liliom_ctx* c, const void* feats, int n, int stride, const double pose7[7], double kd_max_radius,
                                      double surf_dist_thres, double w_gate, double lidar_const, unsigned char* valid, float* plane,
                                      double* score) {
    return correspond_surf_impl(c, feats, n, stride, pose7, kd_max_radius, surf_dist_thres, w_gate, lidar_const, false, 0.0, valid, plane, score);
}

extern "C" int liliom_correspond_surf_refl(liliom_ctx* c, const void* feats48, int n, const double pose7[7], double kd_max_radius,
                                           double surf_dist_thres, double w_gate, double lidar_const, double reflect_thres,
                                           unsigned char* valid, float* plane, double* score) {
    return correspond_surf_impl(c, feats48, n, 48, pose7, kd_max_radius, surf_dist_thres, w_gate, lidar_const, true, reflect_thres, valid, plane, score);
}

extern "C" int liliom_backend_edge_block(liliom_ctx* c, const double pose7_body[7], double s_weight, double cauchy_b, double out29[29]) {
    if (!c || !pose7_body || !out29 || !(cauchy_b > 0)) return LILIOM_E_ARG;
    LILI_CUDA(c, cudaSetDevice(c->device));
    if (c->bk_kind != 1) { c->last_error = "liliom_backend_edge_block needs the correspondences of a preceding liliom_correspond_edge call"; return LILIOM_E_ARG; }
    BlkArgs a{};
    a.q = Q4{pose7_body[0], pose7_body[1], pose7_body[2], pose7_body[3]}; a.t = D3{pose7_body[4], pose7_body[5], pose7_body[6]};
    a.pa = c->corr_plane.as<float>(); a.pb = a.pa + 3 * (size_t)c->bk_n;
    a.s_weight = s_weight; a.cauchy_b = cauchy_b;
    return backend_block_run(c, true, a, out29);
}

extern "C" int liliom_backend_surf_block(liliom_ctx* c, const double pose7_body[7], const double q_lb_wxyz[4], const double t_lb[3],
                                         double cauchy_b, double out29[29]) {
    if (!c || !pose7_body || !q_lb_wxyz || !t_lb || !out29 || !(cauchy_b > 0)) return LILIOM_E_ARG;
    LILI_CUDA(c, cudaSetDevice(c->device));
    if (c->bk_kind != 2) { c->last_error = "liliom_backend_surf_block needs the correspondences of a preceding liliom_correspond_surf* call"; return LILIOM_E_ARG; }
    BlkArgs a{};
    a.q = Q4{pose7_body[0], pose7_body[1], pose7_body[2], pose7_body[3]}; a.t = D3{pose7_body[4], pose7_body[5], pose7_body[6]};
    // Eigen::Quaternion::inverse(): conjugate / squared norm (zero quaternion for a zero input)
    const double n2 = q_lb_wxyz[0] * q_lb_wxyz[0] + q_lb_wxyz[1] * q_lb_wxyz[1] + q_lb_wxyz[2] * q_lb_wxyz[2] + q_lb_wxyz[3] * q_lb_wxyz[3];
    a.qlb_inv = n2 > 0 ? Q4{q_lb_wxyz[0] / n2, -q_lb_wxyz[1] / n2, -q_lb_wxyz[2] / n2, -q_lb_wxyz[3] / n2} : Q4{0, 0, 0, 0};
    a.tlb = D3{t_lb[0], t_lb[1], t_lb[2]};
    a.plane = c->corr_plane.as<float4>(); a.score = c->nn_sqd.as<double>();
    a.cauchy_b = cauchy_b;
    return backend_block_run(c, false, a, out29);
}
