// Internal context of libliliom_b200.so (not part of the ABI).
#pragma once
#include <climits>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/liliom.h"

namespace lili {

// One growable device allocation.  Buffers only grow; 180 GB of HBM3e per GPU makes
// reallocation-on-demand a cold path (first scan), never a steady-state cost.
struct DevBuf {
    void*  p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

// Parameters of one VoxelGrid pass, produced on the device (no host round trip).
struct VgParams {
    float inv_leaf;
    int   min_b[3];
    int   div_b[3];
    int   mul[3];
    int   overflow;    // PCL: "Leaf size is too small" -> output = input
    int   n_finite;
    int   bail;        // cooperative single-launch filter declined this input (see k_vg_coop): redo with the sort chain
};

// Dense cell grid over the down-sampled map (the kd-tree's stand-in).
struct GridDesc {
    float inv_cell;        // 1 / cell size (power of two => exact)
    int   org[3];          // cell coordinate of grid cell (0,0,0)
    int   dim[3];
    int   ncells;
};

// Fused multi-GPU exchange (single node): pointers to every rank's exchange buffer (own entry = own buffer), see
// peer_exchange() in grid_knn.cu.  Passed to the persistent GN kernel by value.
constexpr int kMaxPeers = 16;
struct PeerArgs {
    ulonglong2* buf[kMaxPeers];
    int nranks, rank;
    unsigned int epoch0;       // epoch of iteration 0 of this launch (monotonic across launches, identical on all ranks)
    int enabled;
    double* lsum;              // [2 parities][32] the sums over all ranks, republished by block 0 for the other blocks of this GPU
    unsigned int* lflag;       // [2] epoch of the republished sums
};

// Scan-to-map results written straight into the context's pinned host block by the persistent GN kernel (zero-copy: the
// stream then ends with the kernel and the host reads the block after its stream synchronise), and the start pose as a
// kernel parameter: per scan this removes one host-to-device and three device-to-host copy operations from the stream,
// each of which costs a few microseconds of dependent latency on a ~200 us step.  Layout of host_out = s2m_run's `hp`.
struct GnIo {
    double pose0[8];           // start pose (wxyz, t) + cleared peer-loss flag; used when use_pose0
    int    use_pose0;
    int    n_vgp_words;        // 32-bit words of VgParams to forward (0: none)
    double* host_out;          // device-visible address of the pinned block, nullptr: results stay on the device
    const int* vgp;            // VgParams of the scan's VoxelGrid (speculation check), forwarded to host_out + 48
};

#ifdef __CUDACC__
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#endif

constexpr int kStatsDoubles = 40;   // per outer iteration on the device: n_corr, lm_iters, cost, 27, pose7, pad
constexpr int kNormEq = 29;         // 21 + 6 + cost + count

struct Frame {        // one entry of recent_surf_frames (world frame, point_stride bytes per point)
    DevBuf buf;
    int n = 0;
    // incremental map (map_inc.cu): slot id carried by the frame's entries, finite points, box (ordered ints), unrepresentable key seen
    int slot = 0, nfin = 0, box[6] = {0, 0, 0, 0, 0, 0};
    bool bad = false;
    // box and finite count of the stored points (k_vg_minmax's encoding), taken when the frame was pushed: the rebuild's
    // VoxelGrid and cell grid get their bounding boxes from the union over the frames instead of two passes over the map
    int mm[7] = {INT_MAX, INT_MAX, INT_MAX, INT_MIN, INT_MIN, INT_MIN, 0};
};

}  // namespace lili

struct liliom_ctx {
    liliom_params prm;
    int device = 0;
    cudaStream_t stream = nullptr;       // stream all work is issued on (own_stream unless liliom_set_stream)
    cudaStream_t own_stream = nullptr;
    cudaStream_t copy_stream = nullptr;  // D2H of a finished output while later kernels of the same call still run
    cudaEvent_t  ev_ready = nullptr, ev_copied = nullptr;
    void*  early_cut_dst = nullptr;      // host destination for the cutted cloud (set per call by the API layer)
    int    early_cut_cap = 0;
    bool   early_cut_issued = false;
    std::string last_error;
    int sm_count = 148;

    // ---- staging (pinned host) ----
    void*  h_pin = nullptr;      // small pinned block: counts, pose, stats
    void*  h_pin_dev = nullptr;  // the same block as the device sees it (mapped); nullptr: not mappable, results are copied
    bool   host_results = true;  // LILIOM_HOST_RESULTS=0: always copy (A/B switch)
    size_t h_pin_bytes = 0;

    // ---- extraction ----
    lili::DevBuf raw, cut, surf, edge, flags, scan_tmp, idx_a, idx_b;
    lili::DevBuf hz_mat, hz_stage_surf, hz_stage_edge, hz_counts;
    lili::DevBuf rot_keys, rot_keys2, rot_vals, rot_vals2, rot_cloud, rot_curv, rot_label, rot_picked, rot_sort, rot_ring, rot_meta, rot_lessflat, rot_seg_edge;
    int n_surf_dev = 0;          // surf features resident after the last extract call (-1: only known on the device)
    const int* d_nsurf = nullptr; // device-side count of the resident surf features
    int n_surf_max = 0;          // host upper bound for it
    const int* d_nfeats = nullptr; // device-side query count for scan-to-map (nullptr: n_feats is exact)
    int last_n_feats = 0;        // query count of the previous scan (kernel-shape predictor)
    int n_feats_actual = 0;      // query count read back with the pose
    bool vg_check = false;       // s2m_run also reads vg_params back (speculative key width of the scan VoxelGrid)
    bool vg_used24 = false;      // ... the sort chain ran with 24-bit keys
    bool vg_bail = false;        // ... the cooperative filter declined (valid after the sync)
    unsigned int vg_coop_calls = 0;   // launches of k_vg_coop so far (selects the rotating control slot)
    lili::DevBuf vg_coop;        // hash table + scratch of the cooperative filter
    lili::DevBuf hz_ctl;         // barrier words + per-block counts of the cooperative Horizon extractor
    unsigned int hz_coop_calls = 0;
    long long vg_ncells = 0;     // voxel-box cell count of that VoxelGrid, valid after the sync
    lili::DevBuf raw_scan;       // resident raw sweep (liliom_upload_scan / liliom_convert_livox)
    lili::DevBuf livox_in;       // staged livox CustomPoint records (19/20 bytes each)
    int n_raw_scan = 0;
    const void* raw_src = nullptr; // when set, the extractors read the sweep from here instead of c->raw (resident pipeline: no copy)
    int n_rot_cloud = 0;

    // ---- voxel grid scratch ----
    lili::DevBuf vg_keys, vg_keys2, vg_vals, vg_vals2, vg_flags, vg_rank, vg_params, vg_out, vg_minmax, vg_count;
    lili::DevBuf cub_tmp;

    // ---- map ----
    std::vector<lili::Frame> frames;     // FIFO, oldest first
    // incremental map (liliom_map_update, map_inc.cu): entries {voxel key, slot<<24|index} sorted by (key, frame age, index), double-buffered
    lili::DevBuf inc_key[2], inc_ref[2], inc_newkey[2], inc_newref[2], inc_removed, inc_rpos, inc_flags, inc_rank, inc_mm;
    int inc_cur = 0, inc_E = 0;
    bool inc_valid = false;
    lili::DevBuf map_raw;                // concatenated frames (stride bytes)
    lili::DevBuf map_ds;                 // VoxelGrid output (stride bytes) or installed float4
    lili::DevBuf map_xyzw;               // float4 in map_download order (w = index)
    lili::DevBuf map_refl;               // optional per-point reflectivity channel (liliom_map_set_cloud)
    lili::DevBuf map_sorted;             // float4 sorted by cell, w = original index bits
    lili::DevBuf cell_start;             // ncells + 1
    lili::DevBuf grid_keys, grid_keys2, grid_vals, grid_vals2;
    lili::GridDesc grid{};
    int map_n = 0;                       // points resident on this rank (sorted grid)
    int map_n_global = 0;
    bool map_ready = false;

    // ---- scan-to-map ----
    lili::DevBuf feats;                  // float4 body-frame queries
    int n_feats = 0;
    lili::DevBuf corr_valid, corr_plane, nn_idx, nn_sqd;
    lili::DevBuf qstate;                 // float4 per query: transformed position + fifth distance of the previous GN pass
    lili::DevBuf pose_dev;               // 7 doubles (current) + 7 (candidate)
    lili::DevBuf partials;               // grid x 29 doubles
    lili::DevBuf neq;                    // 29 doubles (reduced)
    lili::DevBuf stats_dev;              // iterations x kStatsDoubles
    lili::DevBuf counter;                // last-block ticket + scratch ints
    lili::DevBuf lm_state;
    int bk_kind = 0, bk_n = 0;           // backend correspondences resident from the last liliom_correspond_* call (1 edge, 2 surf)
    unsigned int bar_arrivals = 0;       // total grid-barrier arrivals issued so far (persistent GN kernel)

    // ---- multi-GPU ----
    void* nccl_comm = nullptr;
    int nranks = 1, rank = 0;
    float shard_inv_block = 0.0625f;     // 1 / shard block edge (LILIOM_SHARD_BLOCK, metres, power of two; must agree on all ranks)
    lili::DevBuf peer_local;             // block 0 -> other blocks of this GPU: [2][32] doubles + 2 epoch words (fused exchange)
    lili::DevBuf peer_buf;               // this rank's exchange buffer: [2 parities][kMaxPeers][32] {epoch|lo32, epoch|hi32}
    void* peer_ptrs[lili::kMaxPeers] = {};   // every rank's buffer as mapped into this process (cudaIpcOpenMemHandle)
    bool peer_ready = false;
    unsigned int peer_epoch = 0;         // last epoch issued

    // ---- instrumentation ----
    liliom_counters cnt{};
    bool time_kernels = false;
    int  time_every = 1;          // liliom_set_kernel_timing(c, N): time every N-th scan-to-map call
    unsigned time_calls = 0;
    int force_lanes = 0, force_rounds = 0;   // tuning override (LILIOM_KNN_LANES / LILIOM_KNN_ROUNDS)
    int gn_sync = 3;                     // persistent GN kernel grid barrier (LILIOM_GN_SYNC): 3 = release-only arrival, no acquire fence (default),
                                         // 0 = full fences on both sides
    bool dbg_timing = false;             // LILIOM_DEBUG_TIMING at create: stage clocks of the cooperative kernels, printed by s2m_run
    bool knn_tma = false, knn_tma_smem_set = false;   // LILIOM_KNN_TMA=1: bulk-copy (cp.async.bulk + mbarrier) staging of the runs in the 16-lane search
    bool knn1_smem_set = false;          // dynamic shared memory limit raised for the one-thread-per-query kernels on this device
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used = 0;
    std::vector<std::pair<size_t, unsigned long long>> ev_pending;  // (event pair index, queries)
    std::vector<int> ev_iters;           // kernel-body passes covered by each pending pair
};

namespace lili {

inline int fail_cuda(liliom_ctx* c, cudaError_t e, const char* where) {
    char buf[256];
    snprintf(buf, sizeof(buf), "%s: %s", where, cudaGetErrorString(e));
    if (c) c->last_error = buf;
    return LILIOM_E_CUDA;
}

#define LILI_CUDA(c, expr)                                                      \
    do {                                                                        \
        cudaError_t _e = (expr);                                                \
        if (_e != cudaSuccess) return lili::fail_cuda((c), _e, #expr);          \
    } while (0)

#define LILI_TRY(expr)                    \
    do {                                  \
        int _r = (expr);                  \
        if (_r != LILIOM_OK) return _r;   \
    } while (0)

inline int launch_check(liliom_ctx* c, const char* name) {
    cudaError_t e = cudaGetLastError();
    c->cnt.launches++;
    if (e != cudaSuccess) return fail_cuda(c, e, name);
    return LILIOM_OK;
}

inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- modules (implemented in the .cu files) ----
int sort_pairs_u32(liliom_ctx* c, const uint32_t* kin, uint32_t* kout, const int* vin, int* vout, int n, int end_bit);
int sort_pairs_u64(liliom_ctx* c, const unsigned long long* kin, unsigned long long* kout, const int* vin, int* vout, int n, int end_bit);
int exclusive_scan_i32(liliom_ctx* c, const int* in, int* out, int n);   // out has n+1 entries (total at out[n])
int inclusive_max_scan_i32(liliom_ctx* c, int* data, int n);            // in-place running maximum

// VoxelGrid on device buffers; d_count receives the output count (int, device).
int voxelgrid_dev(liliom_ctx* c, const void* d_in, int n, int stride, float leaf, void* d_out, int* d_count, float4* d_feats = nullptr,
                  const int* host_mm = nullptr);
int vg_minmax_dev(liliom_ctx* c, const void* d_in, int n_max, const int* d_n, int stride);
// n_max = host upper bound, d_n = optional device-side count (<= n_max); d_feats (optional) also receives float4{x,y,z,index}
int voxelgrid_coop(liliom_ctx* c, const void* d_in, int n_max, const int* d_n, int stride, float leaf, void* d_out, int* d_count, float4* d_feats,
                   bool* used);
int voxelgrid_dev2(liliom_ctx* c, const void* d_in, int n_max, const int* d_n, int stride, float leaf, void* d_out, int* d_count, float4* d_feats,
                   int key_bits = 32, const int* host_mm = nullptr);

// grid build from float4 points already on the device (map_xyzw[0..m)).  host_box (optional): 6 ordered ints, a box known to
// contain the points up to rounding of a mean (the union of the frames' boxes): no min/max pass and no host round trip.
int grid_build(liliom_ctx* c, int m, const int* host_box = nullptr);

const long long* vg_coop_stamps(liliom_ctx* c);   // voxelgrid.cu

int s2m_run(liliom_ctx* c, double pose7[7], int match_cnt, int max_num_iter, int mode, liliom_iter_stats* stats,
            bool want_corr, double out29[29]);

int horizon_extract_dev(liliom_ctx* c, int n, const double q_imu[4], int* n_surf, int* n_edge, int* n_cut, bool sync_counts = true);
int rot_extract_dev(liliom_ctx* c, int n, const double q_imu[4], const double q_lb[4], int* n_surf, int* n_edge, int* n_cut);

int icp_align(liliom_ctx* c, const float4* d_src, int n, double max_corr_dist, int max_iter, double trans_eps, double fit_eps,
              double T16[16], double* fitness, int* converged, int* iters);           // icp.cu
int map_inc_update(liliom_ctx* c, int popped_slot, int popped_nfin, int* m_out);     // map_inc.cu
int map_finish_from_ds(liliom_ctx* c, int m);                                        // api.cu
void frames_box(const liliom_ctx* c, int mm[7]);                                     // api.cu: union of the frames' boxes

int block27_stats(liliom_ctx* c, const double pose7[7], unsigned long long out[2]);   // grid_knn.cu

int repack_to_f4(liliom_ctx* c, const void* d_in, int n, int stride, float4* d_out, const int* d_n = nullptr);

}  // namespace lili
