// pcl::VoxelGrid<PointT>::filter on the device (L/src/LidarOdometry.cpp:315-323 leaf 0.4;
// R/src/Preprocessing.cpp:502-508 leaf 0.6): bounding box -> integer voxel index
// (idx = i + j*dx + k*dx*dy, fp32 floor(p*inv_leaf) - min_b) -> stable sort by index ->
// per-voxel centroid of all fields in ascending index order.  PCL's std::sort leaves the
// within-voxel order unspecified; here it is the original point order, so the fp32 sums are
// reproducible and equal to the oracle's.  Compiled with --fmad=false.
// The box / divisions are computed on the device (VgParams): no host round trip inside the call.
#include "ctx.cuh"
#include <climits>

namespace lili {

__device__ __forceinline__ int vg_f2ord(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float vg_ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void k_vg_init(int* mm) {
    if (threadIdx.x < 3) mm[threadIdx.x] = INT_MAX;
    else if (threadIdx.x < 6) mm[threadIdx.x] = INT_MIN;
    else if (threadIdx.x == 6) mm[6] = 0;
}

// mm[0..2] min, mm[3..5] max (ordered ints), mm[6] = finite count
__global__ void k_vg_minmax(const unsigned char* __restrict__ pts, int n, int stride, int* __restrict__ mm) {
    int lo[3] = {INT_MAX, INT_MAX, INT_MAX}, hi[3] = {INT_MIN, INT_MIN, INT_MIN};
    int cnt = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float4 v = *reinterpret_cast<const float4*>(pts + (size_t)i * stride);
        if (!(isfinite(v.x) && isfinite(v.y) && isfinite(v.z))) continue;
        ++cnt;
        int a = vg_f2ord(v.x), b = vg_f2ord(v.y), c = vg_f2ord(v.z);
        lo[0] = min(lo[0], a); hi[0] = max(hi[0], a);
        lo[1] = min(lo[1], b); hi[1] = max(hi[1], b);
        lo[2] = min(lo[2], c); hi[2] = max(hi[2], c);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            lo[k] = min(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
            hi[k] = max(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
        }
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    if ((threadIdx.x & 31) == 0 && cnt > 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { atomicMin(&mm[k], lo[k]); atomicMax(&mm[3 + k], hi[k]); }
        atomicAdd(&mm[6], cnt);
    }
}

__global__ void k_vg_params(const int* __restrict__ mm, float leaf, VgParams* __restrict__ out) {
    if (threadIdx.x != 0) return;
    VgParams p;
    p.inv_leaf = 1.0f / leaf;                       // Eigen::Array4f::Ones() / leaf_size_
    p.n_finite = mm[6];
    p.overflow = 0;
    if (p.n_finite == 0) {
        for (int k = 0; k < 3; ++k) { p.min_b[k] = 0; p.div_b[k] = 1; }
    } else {
        long long d[3];
        for (int k = 0; k < 3; ++k) {
            float lo = vg_ord2f(mm[k]), hi = vg_ord2f(mm[3 + k]);
            d[k] = (long long)((hi - lo) * p.inv_leaf) + 1;
            p.min_b[k] = (int)floorf(lo * p.inv_leaf);
            int max_b = (int)floorf(hi * p.inv_leaf);
            p.div_b[k] = max_b - p.min_b[k] + 1;
        }
        if (d[0] * d[1] * d[2] > (long long)INT_MAX) p.overflow = 1;
    }
    p.mul[0] = 1; p.mul[1] = p.div_b[0]; p.mul[2] = p.div_b[0] * p.div_b[1];
    *out = p;
}

__global__ void k_vg_keys(const unsigned char* __restrict__ pts, int n, int stride, const VgParams* __restrict__ pp,
                          uint32_t* __restrict__ keys, int* __restrict__ vals) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const VgParams p = *pp;
    float4 v = *reinterpret_cast<const float4*>(pts + (size_t)i * stride);
    uint32_t key = 0xffffffffu;
    if (isfinite(v.x) && isfinite(v.y) && isfinite(v.z)) {
        int i0 = (int)(floorf(v.x * p.inv_leaf) - (float)p.min_b[0]);
        int i1 = (int)(floorf(v.y * p.inv_leaf) - (float)p.min_b[1]);
        int i2 = (int)(floorf(v.z * p.inv_leaf) - (float)p.min_b[2]);
        key = (uint32_t)(i0 * p.mul[0] + i1 * p.mul[1] + i2 * p.mul[2]);
    }
    keys[i] = key;
    vals[i] = i;
}

// flags[i] = 1 at the first sorted entry of every occupied voxel; flags[n] = 0 (scan sentinel)
__global__ void k_vg_heads(const uint32_t* __restrict__ keys, int n, const VgParams* __restrict__ pp, int* __restrict__ flags) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    int f = 0;
    if (i < n) {
        int nf = pp->n_finite;   // non-finite points carry key 0xffffffff and sort last; exclude by count
        f = (i < nf) && (i == 0 || keys[i] != keys[i - 1]);
    }
    flags[i] = f;
}

template <int STRIDE>
__global__ void k_vg_centroid(const unsigned char* __restrict__ pts, const uint32_t* __restrict__ keys, const int* __restrict__ vals,
                              const int* __restrict__ flags, const int* __restrict__ rank, int n, const VgParams* __restrict__ pp,
                              unsigned char* __restrict__ out, int cap, int* __restrict__ count_out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    const VgParams p = *pp;
    if (i == 0) *count_out = p.overflow ? n : rank[n];
    if (p.overflow) {   // PCL: output = *input_
        if (i < n && i < cap) {
            const float4* src = reinterpret_cast<const float4*>(pts + (size_t)i * STRIDE);
            float4* dst = reinterpret_cast<float4*>(out + (size_t)i * STRIDE);
#pragma unroll
            for (int k = 0; k < STRIDE / 16; ++k) dst[k] = src[k];
        }
        return;
    }
    if (i >= n || !flags[i]) return;
    const int o = rank[i];
    if (o >= cap) return;
    const uint32_t key = keys[i];
    float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f, sc = 0.f, snx = 0.f, sny = 0.f, snz = 0.f;
    int cnt = 0;
    for (int j = i; j < p.n_finite && keys[j] == key; ++j) {
        const unsigned char* src = pts + (size_t)vals[j] * STRIDE;
        float4 a = *reinterpret_cast<const float4*>(src);
        sx += a.x; sy += a.y; sz += a.z;
        if (STRIDE == 48) {
            float4 b = *reinterpret_cast<const float4*>(src + 16);
            float4 cc = *reinterpret_cast<const float4*>(src + 32);
            snx += b.x; sny += b.y; snz += b.z;
            si += cc.x; sc += cc.y;
        } else {
            float4 b = *reinterpret_cast<const float4*>(src + 16);
            si += b.x;
        }
        ++cnt;
    }
    const float fc = (float)cnt;
    unsigned char* dst = out + (size_t)o * STRIDE;
    *reinterpret_cast<float4*>(dst) = make_float4(sx / fc, sy / fc, sz / fc, 1.0f);
    if (STRIDE == 48) {
        // pcl::CentroidPoint: normal accumulated then normalised (not divided), curvature/intensity averaged
        float n2 = snx * snx + sny * sny + snz * snz;
        if (n2 > 0.0f) { float nn = sqrtf(n2); snx = snx / nn; sny = sny / nn; snz = snz / nn; }
        *reinterpret_cast<float4*>(dst + 16) = make_float4(snx, sny, snz, 0.0f);
        *reinterpret_cast<float4*>(dst + 32) = make_float4(si / fc, sc / fc, 0.0f, 0.0f);
    } else {
        *reinterpret_cast<float4*>(dst + 16) = make_float4(si / fc, 0.0f, 0.0f, 0.0f);
    }
}

int voxelgrid_dev(liliom_ctx* c, const void* d_in, int n, int stride, float leaf, void* d_out, int* d_count) {
    if (stride != 48 && stride != 32) return LILIOM_E_ARG;
    if (n <= 0) {
        LILI_CUDA(c, cudaMemsetAsync(d_count, 0, sizeof(int), c->stream));
        return LILIOM_OK;
    }
    LILI_CUDA(c, c->vg_minmax.ensure(8 * sizeof(int)));
    LILI_CUDA(c, c->vg_params.ensure(sizeof(VgParams)));
    LILI_CUDA(c, c->vg_keys.ensure((size_t)n * 4));
    LILI_CUDA(c, c->vg_keys2.ensure((size_t)n * 4));
    LILI_CUDA(c, c->vg_vals.ensure((size_t)n * 4));
    LILI_CUDA(c, c->vg_vals2.ensure((size_t)n * 4));
    LILI_CUDA(c, c->vg_flags.ensure(((size_t)n + 2) * 4));
    LILI_CUDA(c, c->vg_rank.ensure(((size_t)n + 2) * 4));
    const unsigned char* in = (const unsigned char*)d_in;
    int* mm = c->vg_minmax.as<int>();
    VgParams* pp = c->vg_params.as<VgParams>();
    k_vg_init<<<1, 32, 0, c->stream>>>(mm);
    LILI_TRY(launch_check(c, "k_vg_init"));
    k_vg_minmax<<<min(cdiv(n, 256), c->sm_count * 8), 256, 0, c->stream>>>(in, n, stride, mm);
    LILI_TRY(launch_check(c, "k_vg_minmax"));
    k_vg_params<<<1, 32, 0, c->stream>>>(mm, leaf, pp);
    LILI_TRY(launch_check(c, "k_vg_params"));
    k_vg_keys<<<cdiv(n, 256), 256, 0, c->stream>>>(in, n, stride, pp, c->vg_keys.as<uint32_t>(), c->vg_vals.as<int>());
    LILI_TRY(launch_check(c, "k_vg_keys"));
    LILI_TRY(sort_pairs_u32(c, c->vg_keys.as<uint32_t>(), c->vg_keys2.as<uint32_t>(), c->vg_vals.as<int>(), c->vg_vals2.as<int>(), n, 32));
    k_vg_heads<<<cdiv(n + 1, 256), 256, 0, c->stream>>>(c->vg_keys2.as<uint32_t>(), n, pp, c->vg_flags.as<int>());
    LILI_TRY(launch_check(c, "k_vg_heads"));
    LILI_TRY(exclusive_scan_i32(c, c->vg_flags.as<int>(), c->vg_rank.as<int>(), n));
    if (stride == 48)
        k_vg_centroid<48><<<cdiv(n, 128), 128, 0, c->stream>>>(in, c->vg_keys2.as<uint32_t>(), c->vg_vals2.as<int>(), c->vg_flags.as<int>(),
                                                               c->vg_rank.as<int>(), n, pp, (unsigned char*)d_out, INT_MAX, d_count);
    else
        k_vg_centroid<32><<<cdiv(n, 128), 128, 0, c->stream>>>(in, c->vg_keys2.as<uint32_t>(), c->vg_vals2.as<int>(), c->vg_flags.as<int>(),
                                                               c->vg_rank.as<int>(), n, pp, (unsigned char*)d_out, INT_MAX, d_count);
    LILI_TRY(launch_check(c, "k_vg_centroid"));
    return LILIOM_OK;
}

}  // namespace lili
