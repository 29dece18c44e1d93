// pcl::VoxelGrid<PointT>::filter on the device (L/src/LidarOdometry.cpp:315-323 leaf 0.4;
// R/src/Preprocessing.cpp:502-508 leaf 0.6): bounding box -> integer voxel index
// (idx = i + j*dx + k*dx*dy, fp32 floor(p*inv_leaf) - min_b) -> stable sort by index ->
// per-voxel centroid of all fields in ascending index order.  PCL's std::sort leaves the
// within-voxel order unspecified; here it is the original point order, so the fp32 sums are
// reproducible and equal to the oracle's.  Compiled with --fmad=false.
// The box / divisions are computed on the device (VgParams): no host round trip inside the call.
#include "ctx.cuh"
#include <climits>
#include <cstdlib>
#include <cub/block/block_radix_sort.cuh>
#include <cub/block/block_scan.cuh>

namespace lili {

__device__ __forceinline__ int vg_f2ord(float f) { int i = __float_as_int(f); return i >= 0 ? i : i ^ 0x7fffffff; }
__device__ __forceinline__ float vg_ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void k_vg_init(int* mm) {
    if (threadIdx.x < 3) mm[threadIdx.x] = INT_MAX;
    else if (threadIdx.x < 6) mm[threadIdx.x] = INT_MIN;
    else if (threadIdx.x == 6) mm[6] = 0;
}

// mm[0..2] min, mm[3..5] max (ordered ints), mm[6] = finite count
__global__ void k_vg_minmax(const unsigned char* __restrict__ pts, int n_max, const int* __restrict__ d_n, int stride, int* __restrict__ mm) {
    const int n = d_n ? min(*d_n, n_max) : n_max;
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

__global__ void k_vg_keys(const unsigned char* __restrict__ pts, int n_max, const int* __restrict__ d_n, int stride, const VgParams* __restrict__ pp,
                          uint32_t* __restrict__ keys, int* __restrict__ vals) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_max) return;
    const int n = d_n ? min(*d_n, n_max) : n_max;
    const VgParams p = *pp;
    float4 v = make_float4(NAN, 0.f, 0.f, 0.f);
    if (i < n) v = *reinterpret_cast<const float4*>(pts + (size_t)i * stride);
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
                              const int* __restrict__ flags, const int* __restrict__ rank, int n_max, const int* __restrict__ d_n,
                              const VgParams* __restrict__ pp, unsigned char* __restrict__ out, int cap, int* __restrict__ count_out,
                              float4* __restrict__ feats_out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    const VgParams p = *pp;
    const int n = d_n ? min(*d_n, n_max) : n_max;
    if (i == 0) *count_out = p.overflow ? n : rank[n_max];
    if (p.overflow) {   // PCL: output = *input_
        if (i < n && i < cap) {
            const float4* src = reinterpret_cast<const float4*>(pts + (size_t)i * STRIDE);
            float4* dst = reinterpret_cast<float4*>(out + (size_t)i * STRIDE);
#pragma unroll
            for (int k = 0; k < STRIDE / 16; ++k) dst[k] = src[k];
            if (feats_out) feats_out[i] = make_float4(src[0].x, src[0].y, src[0].z, __int_as_float(i));
        }
        return;
    }
    if (i >= n || !flags[i]) return;
    const int o = rank[i];
    if (o >= cap) return;
    const uint32_t key = keys[i];
    float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f, sc = 0.f, snx = 0.f, sny = 0.f, snz = 0.f;
    int cnt = 0;
    // strictly sequential fp32 sums (index order), memory latency overlapped in batches of 8 members
    bool more = true;
#pragma unroll 1
    for (int k0 = i; more; k0 += 8) {
        int idx[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) idx[u] = (k0 + u < p.n_finite && keys[k0 + u] == key) ? vals[k0 + u] : -1;
        float4 A[8], B[8], Cc[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (idx[u] >= 0) {
                const unsigned char* src = pts + (size_t)idx[u] * STRIDE;
                A[u] = *reinterpret_cast<const float4*>(src);
                B[u] = *reinterpret_cast<const float4*>(src + 16);
                if (STRIDE == 48) Cc[u] = *reinterpret_cast<const float4*>(src + 32);
            }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (idx[u] >= 0) {
                sx += A[u].x; sy += A[u].y; sz += A[u].z;
                if (STRIDE == 48) { snx += B[u].x; sny += B[u].y; snz += B[u].z; si += Cc[u].x; sc += Cc[u].y; }
                else si += B[u].x;
                ++cnt;
            }
        }
        more = idx[7] >= 0;
    }
    const float fc = (float)cnt;
    unsigned char* dst = out + (size_t)o * STRIDE;
    *reinterpret_cast<float4*>(dst) = make_float4(sx / fc, sy / fc, sz / fc, 1.0f);
    if (feats_out) feats_out[o] = make_float4(sx / fc, sy / fc, sz / fc, __int_as_float(o));
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

// ---------------------------------------------------------------------------------------
// Scan-sized VoxelGrid in ONE CTA (n <= 24576): bounding box, voxel keys, stable block radix sort
// (cub::BlockRadixSort on registers/shared memory, only as many key bits as the box needs), voxel
// heads, block scan, centroids — what used to be ~12 dependent launches (4 onesweep passes of ~11 us
// each for 16k keys) with a host round trip for the count.  Semantics identical to the chain above.
// ---------------------------------------------------------------------------------------
constexpr int VGS_THREADS = 1024;
constexpr int VGS_ITEMS = 24;
constexpr int VGS_CAP = VGS_THREADS * VGS_ITEMS;

typedef cub::BlockRadixSort<uint32_t, VGS_THREADS, VGS_ITEMS, int> VgsSort;
typedef cub::BlockScan<int, VGS_THREADS> VgsScan;

struct VgSmallSmem {
    union {
        typename VgsSort::TempStorage sort;
        typename VgsScan::TempStorage scan;
    } u;
    int red[6][32];
    int cnt[32];
    uint32_t last_key[VGS_THREADS];
    VgParams prm;
    int n;
};

template <int STRIDE>
__global__ void __launch_bounds__(VGS_THREADS) k_vg_small(const unsigned char* __restrict__ pts, int n_max, const int* __restrict__ d_n, float leaf,
                                                          uint32_t* __restrict__ skeys, int* __restrict__ svals, int* __restrict__ head_pos,
                                                          unsigned char* __restrict__ out, int* __restrict__ count_out, float4* __restrict__ feats_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    VgSmallSmem& S = *reinterpret_cast<VgSmallSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = d_n ? min(*d_n, n_max) : n_max;
    // ---- bounding box of the finite points
    int lo[3] = {INT_MAX, INT_MAX, INT_MAX}, hi[3] = {INT_MIN, INT_MIN, INT_MIN}, cnt = 0;
    for (int i = tid; i < n; i += VGS_THREADS) {
        const float4 v = *reinterpret_cast<const float4*>(pts + (size_t)i * STRIDE);
        if (!(isfinite(v.x) && isfinite(v.y) && isfinite(v.z))) continue;
        ++cnt;
        const int a = vg_f2ord(v.x), b = vg_f2ord(v.y), c = vg_f2ord(v.z);
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
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) { S.red[k][warp] = lo[k]; S.red[3 + k][warp] = hi[k]; }
        S.cnt[warp] = cnt;
    }
    __syncthreads();
    if (tid == 0) {
        int mm[7] = {INT_MAX, INT_MAX, INT_MAX, INT_MIN, INT_MIN, INT_MIN, 0};
        for (int w = 0; w < 32; ++w) {
            for (int k = 0; k < 3; ++k) { mm[k] = min(mm[k], S.red[k][w]); mm[3 + k] = max(mm[3 + k], S.red[3 + k][w]); }
            mm[6] += S.cnt[w];
        }
        VgParams p;
        p.inv_leaf = 1.0f / leaf;
        p.n_finite = mm[6];
        p.overflow = 0;
        if (p.n_finite == 0) {
            for (int k = 0; k < 3; ++k) { p.min_b[k] = 0; p.div_b[k] = 1; }
        } else {
            long long d[3];
            for (int k = 0; k < 3; ++k) {
                const float flo = vg_ord2f(mm[k]), fhi = vg_ord2f(mm[3 + k]);
                d[k] = (long long)((fhi - flo) * p.inv_leaf) + 1;
                p.min_b[k] = (int)floorf(flo * p.inv_leaf);
                p.div_b[k] = (int)floorf(fhi * p.inv_leaf) - p.min_b[k] + 1;
            }
            if (d[0] * d[1] * d[2] > (long long)INT_MAX) p.overflow = 1;
        }
        p.mul[0] = 1; p.mul[1] = p.div_b[0]; p.mul[2] = p.div_b[0] * p.div_b[1];
        S.prm = p;
        S.n = n;
    }
    __syncthreads();
    const VgParams p = S.prm;
    if (p.overflow) {   // PCL: output = *input_
        for (int i = tid; i < n; i += VGS_THREADS) {
            const float4* src = reinterpret_cast<const float4*>(pts + (size_t)i * STRIDE);
            float4* dst = reinterpret_cast<float4*>(out + (size_t)i * STRIDE);
#pragma unroll
            for (int k = 0; k < STRIDE / 16; ++k) dst[k] = src[k];
            if (feats_out) feats_out[i] = make_float4(src[0].x, src[0].y, src[0].z, __int_as_float(i));
        }
        if (tid == 0) *count_out = n;
        return;
    }
    // ---- keys (blocked arrangement: thread t owns items t*ITEMS .. t*ITEMS+ITEMS-1, i.e. original order)
    uint32_t keys[VGS_ITEMS];
    int vals[VGS_ITEMS];
#pragma unroll
    for (int j = 0; j < VGS_ITEMS; ++j) {
        const int i = tid * VGS_ITEMS + j;
        uint32_t key = 0xffffffffu;
        if (i < n) {
            const float4 v = *reinterpret_cast<const float4*>(pts + (size_t)i * STRIDE);
            if (isfinite(v.x) && isfinite(v.y) && isfinite(v.z)) {
                const int i0 = (int)(floorf(v.x * p.inv_leaf) - (float)p.min_b[0]);
                const int i1 = (int)(floorf(v.y * p.inv_leaf) - (float)p.min_b[1]);
                const int i2 = (int)(floorf(v.z * p.inv_leaf) - (float)p.min_b[2]);
                key = (uint32_t)(i0 * p.mul[0] + i1 * p.mul[1] + i2 * p.mul[2]);
            }
        }
        keys[j] = key;
        vals[j] = i;
    }
    // sentinel keys are 0xffffffff: sort all 32 bits only if the box needs them, else box bits + 1 for the sentinel
    long long ncell = (long long)p.div_b[0] * p.div_b[1] * p.div_b[2];
    int bits = 1;
    while ((1LL << bits) < ncell && bits < 31) ++bits;
    const int end_bit = min(32, bits + 1);
    if (end_bit < 32) {
#pragma unroll
        for (int j = 0; j < VGS_ITEMS; ++j) if (keys[j] == 0xffffffffu) keys[j] = (1u << (end_bit - 1)) | ((1u << (end_bit - 1)) - 1u);
    }
    VgsSort(S.u.sort).Sort(keys, vals, 0, end_bit);
    __syncthreads();
    // ---- sorted pairs to global scratch (the centroid walk crosses thread boundaries)
#pragma unroll
    for (int j = 0; j < VGS_ITEMS; ++j) { skeys[tid * VGS_ITEMS + j] = keys[j]; svals[tid * VGS_ITEMS + j] = vals[j]; }
    S.last_key[tid] = keys[VGS_ITEMS - 1];
    __syncthreads();
    // ---- voxel heads + ranks
    const int nf = p.n_finite;
    int heads = 0;
    unsigned head_mask = 0;
#pragma unroll
    for (int j = 0; j < VGS_ITEMS; ++j) {
        const int pos = tid * VGS_ITEMS + j;
        const uint32_t prev = j > 0 ? keys[j - 1] : (tid > 0 ? S.last_key[tid - 1] : 0u);
        const bool h = pos < nf && (pos == 0 || keys[j] != prev);
        if (h) { ++heads; head_mask |= 1u << j; }
    }
    int base, total;
    VgsScan(S.u.scan).ExclusiveSum(heads, base, total);
    __syncthreads();   // global scratch written above is visible block-wide from here on
    if (tid == 0) *count_out = total;
    // head positions by rank, so that the voxels can be dealt round-robin to the threads below
    {
        int o = base;
#pragma unroll
        for (int j = 0; j < VGS_ITEMS; ++j) if (head_mask & (1u << j)) head_pos[o++] = tid * VGS_ITEMS + j;
    }
    __syncthreads();
    // ---- centroids: sequential fp32 sums in sorted (= original index) order, like the oracle.
    // One voxel per thread at a time; member indices and points are fetched in batches of 8 independent
    // loads (the sums stay strictly sequential, only the memory latency is overlapped).
#pragma unroll 1
    for (int o = tid; o < total; o += VGS_THREADS) {
        const int pos = head_pos[o];
        const int end = (o + 1 < total) ? head_pos[o + 1] : nf;
        float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f, sc = 0.f, snx = 0.f, sny = 0.f, snz = 0.f;
#pragma unroll 1
        for (int k0 = pos; k0 < end; k0 += 8) {
            int idx[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) idx[u] = (k0 + u < end) ? svals[k0 + u] : -1;
            float4 A[8], B[8], Cc[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (idx[u] >= 0) {
                    const unsigned char* src = pts + (size_t)idx[u] * STRIDE;
                    A[u] = *reinterpret_cast<const float4*>(src);
                    B[u] = *reinterpret_cast<const float4*>(src + 16);
                    if (STRIDE == 48) Cc[u] = *reinterpret_cast<const float4*>(src + 32);
                }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (idx[u] >= 0) {
                    sx += A[u].x; sy += A[u].y; sz += A[u].z;
                    if (STRIDE == 48) { snx += B[u].x; sny += B[u].y; snz += B[u].z; si += Cc[u].x; sc += Cc[u].y; }
                    else si += B[u].x;
                }
            }
        }
        const float fc = (float)(end - pos);
        unsigned char* dst = out + (size_t)o * STRIDE;
        *reinterpret_cast<float4*>(dst) = make_float4(sx / fc, sy / fc, sz / fc, 1.0f);
        if (feats_out) feats_out[o] = make_float4(sx / fc, sy / fc, sz / fc, __int_as_float(o));
        if (STRIDE == 48) {
            const float n2 = snx * snx + sny * sny + snz * snz;
            if (n2 > 0.0f) { const float nn = sqrtf(n2); snx = snx / nn; sny = sny / nn; snz = snz / nn; }
            *reinterpret_cast<float4*>(dst + 16) = make_float4(snx, sny, snz, 0.0f);
            *reinterpret_cast<float4*>(dst + 32) = make_float4(si / fc, sc / fc, 0.0f, 0.0f);
        } else {
            *reinterpret_cast<float4*>(dst + 16) = make_float4(si / fc, 0.0f, 0.0f, 0.0f);
        }
    }
}

int voxelgrid_dev2(liliom_ctx* c, const void* d_in, int n_max, const int* d_n, int stride, float leaf, void* d_out, int* d_count,
                   float4* d_feats, int key_bits) {
    if (stride != 48 && stride != 32) return LILIOM_E_ARG;
    if (n_max <= 0) {
        LILI_CUDA(c, cudaMemsetAsync(d_count, 0, sizeof(int), c->stream));
        return LILIOM_OK;
    }
    const size_t scratch = (size_t)(n_max > VGS_CAP ? n_max : VGS_CAP) * 4;
    LILI_CUDA(c, c->vg_keys2.ensure(scratch));
    LILI_CUDA(c, c->vg_vals2.ensure(scratch));
    LILI_CUDA(c, c->vg_rank.ensure(scratch + 8));
    const unsigned char* in = (const unsigned char*)d_in;
    if (n_max <= VGS_CAP && getenv("LILIOM_VG_SMALL")) {   // opt-in: a single SM sorts 16k keys in ~300 us, the multi-CTA chain below in ~50
        // scan-sized input: the whole filter in ONE persistent CTA (no host-visible intermediate, 1 launch instead of ~12)
        LILI_CUDA(c, cudaFuncSetAttribute(k_vg_small<48>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(VgSmallSmem)));
        LILI_CUDA(c, cudaFuncSetAttribute(k_vg_small<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(VgSmallSmem)));
        if (stride == 48)
            k_vg_small<48><<<1, VGS_THREADS, sizeof(VgSmallSmem), c->stream>>>(in, n_max, d_n, leaf, c->vg_keys2.as<uint32_t>(), c->vg_vals2.as<int>(),
                                                                                c->vg_rank.as<int>(), (unsigned char*)d_out, d_count, d_feats);
        else
            k_vg_small<32><<<1, VGS_THREADS, sizeof(VgSmallSmem), c->stream>>>(in, n_max, d_n, leaf, c->vg_keys2.as<uint32_t>(), c->vg_vals2.as<int>(),
                                                                                c->vg_rank.as<int>(), (unsigned char*)d_out, d_count, d_feats);
        return launch_check(c, "k_vg_small");
    }
    const int n = n_max;
    LILI_CUDA(c, c->vg_minmax.ensure(8 * sizeof(int)));
    LILI_CUDA(c, c->vg_params.ensure(sizeof(VgParams)));
    LILI_CUDA(c, c->vg_keys.ensure((size_t)n * 4));
    LILI_CUDA(c, c->vg_vals.ensure((size_t)n * 4));
    LILI_CUDA(c, c->vg_flags.ensure(((size_t)n + 2) * 4));
    LILI_CUDA(c, c->vg_rank.ensure(((size_t)n + 2) * 4));
    int* mm = c->vg_minmax.as<int>();
    VgParams* pp = c->vg_params.as<VgParams>();
    k_vg_init<<<1, 32, 0, c->stream>>>(mm);
    LILI_TRY(launch_check(c, "k_vg_init"));
    k_vg_minmax<<<min(cdiv(n, 256), c->sm_count * 8), 256, 0, c->stream>>>(in, n, d_n, stride, mm);
    LILI_TRY(launch_check(c, "k_vg_minmax"));
    k_vg_params<<<1, 32, 0, c->stream>>>(mm, leaf, pp);
    LILI_TRY(launch_check(c, "k_vg_params"));
    k_vg_keys<<<cdiv(n, 256), 256, 0, c->stream>>>(in, n, d_n, stride, pp, c->vg_keys.as<uint32_t>(), c->vg_vals.as<int>());
    LILI_TRY(launch_check(c, "k_vg_keys"));
    // key_bits < 32: the caller speculates that the voxel index fits (one onesweep pass less per 8 bits);
    // invalid (sentinel) keys are folded to the all-ones key of that width and VgParams::n_finite excludes them.
    // The caller must check `ncells <= 2^key_bits - 1` afterwards (vg_params stays on the device).
    LILI_TRY(sort_pairs_u32(c, c->vg_keys.as<uint32_t>(), c->vg_keys2.as<uint32_t>(), c->vg_vals.as<int>(), c->vg_vals2.as<int>(), n, key_bits));
    k_vg_heads<<<cdiv(n + 1, 256), 256, 0, c->stream>>>(c->vg_keys2.as<uint32_t>(), n, pp, c->vg_flags.as<int>());
    LILI_TRY(launch_check(c, "k_vg_heads"));
    LILI_TRY(exclusive_scan_i32(c, c->vg_flags.as<int>(), c->vg_rank.as<int>(), n));
    if (stride == 48)
        k_vg_centroid<48><<<cdiv(n, 128), 128, 0, c->stream>>>(in, c->vg_keys2.as<uint32_t>(), c->vg_vals2.as<int>(), c->vg_flags.as<int>(),
                                                               c->vg_rank.as<int>(), n, d_n, pp, (unsigned char*)d_out, INT_MAX, d_count, d_feats);
    else
        k_vg_centroid<32><<<cdiv(n, 128), 128, 0, c->stream>>>(in, c->vg_keys2.as<uint32_t>(), c->vg_vals2.as<int>(), c->vg_flags.as<int>(),
                                                               c->vg_rank.as<int>(), n, d_n, pp, (unsigned char*)d_out, INT_MAX, d_count, d_feats);
    LILI_TRY(launch_check(c, "k_vg_centroid"));
    return LILIOM_OK;
}

int voxelgrid_dev(liliom_ctx* c, const void* d_in, int n, int stride, float leaf, void* d_out, int* d_count) {
    return voxelgrid_dev2(c, d_in, n, nullptr, stride, leaf, d_out, d_count, nullptr, 32);
}

}  // namespace lili
