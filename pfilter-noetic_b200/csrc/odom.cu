// Odometry + persistence filter + local map: device-resident replacement of Odom_ES_EstimationClass
// (/root/reference/include/odomEstimationClass.h:140-167, src/odomEstimationClass.cpp:182-282, :589-647).
//
// One pf_odom_update enqueues, on the handle's stream and without any host round trip in between:
//   k_predict        constant-velocity prediction odom * (last^-1 * odom)                         (:235-240)
//   voxelize(PCL)    VoxelGrid down-sampling of both feature clouds, leaf 0.4 / 0.8                (:244-245)
//   build_grids      1 m search grids over both local maps (stands in for the two kd-tree builds)  (:249-250)
//   optimization_count x { associate_pass (kNN + line/plane fit + persistence counters), 5 x k_lm_eval }   (:252-272)
//   k_append         transform all down-sampled points with the final pose and append them to the maps      (:592-604)
//   map_merge        CropBox + rgbds + extractstablepoint + r += 2 as a streaming merge (merge.cuh) (:606-647)
// and finally copies the 7-double pose to the host.  Maps, counters and the pose stay in HBM between frames.
//
// The same machinery runs Odom_BPF_EstimationClass (include/odomEstimationClass.h:169-205, src/odomEstimationClass.cpp:649-1306),
// the reference's second odometry class: identical arithmetic over three feature kinds -- beam and pillar (point-to-line, leaf
// = map_resolution) and facade (point-to-plane, leaf = 2 x map_resolution).  The kernels process feature kinds in PAIRS (one
// line-type + one plane-type cloud per launch, blockIdx.y); the ES path is one pair (edge, surf), the BPF path two pairs
// (beam, facade) and (pillar, <empty>).
#include <vector>

#include "match.cuh"
#include "math.cuh"
#include "merge.cuh"
#include "solve.cuh"
#include "voxel.cuh"

namespace pf {

struct IsoDev { double R[9]; double t[3]; };

struct OdomShared {   // small device-resident block
    IsoDev odom, last_odom;
    double pose[7];        // pose of the last finished update (what the caller reads)
    int n_app[4];          // map sizes after appending the new points (per kind slot)
    int err;               // bit 0: map capacity exceeded; bits 1..4: map merge (voxel coordinate range, exception capacity, internal);
                           // kErrGrid / kErrVoxel / kErrRing (primitives.cuh): search grid, VoxelGrid index space, extractor ring capacity
    int guard;             // the maps hold enough points to associate (:247 / BPF :720)
    int n_map[4];          // map sizes after the last update (read back with the pose: tight launch bounds for the next frame)
    long long frame;       // frame index this block describes
    int n_ds[4];           // down-sampled feature counts of the last update (read back with the pose: launch geometry of the next frames)
};

constexpr int kKinds = 4;     // kind slots of a handle: ES uses 0 (edge) 1 (surf); BPF 0 (beam) 1 (pillar) 2 (facade); slot 3 stays empty

__global__ void k_odom_reset(OdomShared* sh, LmState* S) {
    if (threadIdx.x != 0) return;
    for (int i = 0; i < 9; ++i) { sh->odom.R[i] = (i % 4 == 0) ? 1.0 : 0.0; sh->last_odom.R[i] = sh->odom.R[i]; }
    for (int i = 0; i < 3; ++i) { sh->odom.t[i] = 0; sh->last_odom.t[i] = 0; }
    const double id[7] = {0, 0, 0, 1, 0, 0, 0};
    for (int i = 0; i < 7; ++i) { sh->pose[i] = id[i]; S->x[i] = id[i]; }
    for (int k = 0; k < 4; ++k) { sh->n_app[k] = 0; sh->n_map[k] = 0; sh->n_ds[k] = 0; }
    sh->err = 0;
    sh->guard = 0;
}

// odom_prediction = odom * (last_odom.inverse() * odom); last_odom = odom; odom = prediction;
// q_w_curr = Quaterniond(odom.rotation()); t_w_curr = odom.translation()                     (:235-240)
struct GuardSpec { const int* n_map[3]; int min_pts[3]; int count; };

__global__ void k_predict(OdomShared* sh, LmState* S, GuardSpec G) {
    PF_PDL_ENTRY();
    if (threadIdx.x != 0) return;
    {   // laserCloudCornerMap->points.size() > 10 && laserCloudSurfMap->points.size() > 50 (:247); BPF: beam, pillar > 10, facade > 50 (:720)
        int ok = 1;
        for (int k = 0; k < G.count; ++k) ok &= (*G.n_map[k] > G.min_pts[k]) ? 1 : 0;
        sh->guard = ok;
    }
    const IsoDev a = sh->odom, l = sh->last_odom;
    IsoDev li, m, p;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) li.R[3 * i + j] = l.R[3 * j + i];
    for (int i = 0; i < 3; ++i) li.t[i] = -(li.R[3 * i] * l.t[0] + li.R[3 * i + 1] * l.t[1] + li.R[3 * i + 2] * l.t[2]);
    auto mul = [](const IsoDev& x, const IsoDev& y, IsoDev& o) {
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 3; ++j) o.R[3 * i + j] = x.R[3 * i] * y.R[j] + x.R[3 * i + 1] * y.R[3 + j] + x.R[3 * i + 2] * y.R[6 + j];
            o.t[i] = x.R[3 * i] * y.t[0] + x.R[3 * i + 1] * y.t[1] + x.R[3 * i + 2] * y.t[2] + x.t[i];
        }
    };
    mul(li, a, m);
    mul(a, m, p);
    sh->last_odom = a;
    sh->odom = p;
    double q[4];
    mat_to_quat(p.R, q);
    for (int i = 0; i < 4; ++i) S->x[i] = q[i];
    for (int i = 0; i < 3; ++i) S->x[4 + i] = p.t[i];
}

constexpr int kPoseHist = 4096;
constexpr int kMergeExcCap = 4096;   // centroids per update that may leave their voxel by rounding (a handful in practice)

struct AppendParams {
    const Pt* ds[2]; const int* n_ds[2];
    Pt* map[2]; const int* n_map[2];
    int slot[2];      // kind slots of the pair (index into OdomShared::n_app)
    int write_pose;   // first pair of the frame: publish the pose
    OdomShared* sh; const LmState* S;
    int map_cap;      // capacity of the map buffers (points)
    double* pose_hist;   // [kPoseHist][7]; slot = (frame of this update) % kPoseHist, the frame counter lives in OdomShared
};

// odom <- (q_w_curr, t_w_curr) (:278-280) and addPointsToMap's append loop (:592-604)
__global__ void __launch_bounds__(256) k_append(AppendParams A) {
    PF_PDL_ENTRY();
    const int kind = blockIdx.y;
    __shared__ double s_pose[7];
    if (threadIdx.x < 7) s_pose[threadIdx.x] = A.S->x[threadIdx.x];
    __syncthreads();
    const int n = *A.n_ds[kind], m = *A.n_map[kind];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        int tot = m + n;
        if (tot > A.map_cap) { atomicOr(&A.sh->err, 1); tot = A.map_cap; }
        A.sh->n_app[A.slot[kind]] = tot;
        A.sh->n_ds[A.slot[kind]] = n;
        if (kind == 0 && A.write_pose) {
            bool finite = true;
            for (int i = 0; i < 7; ++i) finite = finite && isfinite(s_pose[i]);
            if (!finite) atomicOr(&A.sh->err, (int)kErrPose);
            quat_to_mat(s_pose, A.sh->odom.R);
            for (int i = 0; i < 3; ++i) A.sh->odom.t[i] = s_pose[4 + i];
            const int hist_slot = (int)((A.sh->frame + 1) % kPoseHist);      // this update is frame (last finished + 1)
            for (int i = 0; i < 7; ++i) { A.sh->pose[i] = s_pose[i]; A.pose_hist[7 * hist_slot + i] = s_pose[i]; }
        }
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (m + i >= A.map_cap) break;
        const Pt q = A.ds[kind][i];
        const D3 w = pose_apply(s_pose, d3((double)q.x, (double)q.y, (double)q.z));
        Pt o;
        o.x = (float)w.x; o.y = (float)w.y; o.z = (float)w.z;
        o.rgba = (q.rgba & 0x00ffffffu) | 0xff000000u;   // r, g, b copied (:169-171), a = 255 of a fresh PointXYZRGB
        A.map[kind][m + i] = o;
    }
}

struct InitParams { const float4* feat[2]; const int* n_feat[2]; Pt* map[2]; int* n_map[2]; int* n_sorted[2]; int slot[2]; int map_cap; OdomShared* sh;
                    const unsigned* extract_err; };

// initMapWithPoints (:217-222): the raw first-frame clouds become the maps
__global__ void __launch_bounds__(256) k_init_map(InitParams I) {
    const int kind = blockIdx.y;
    int n = *I.n_feat[kind];
    if (n > I.map_cap) { n = I.map_cap; if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(&I.sh->err, 1); }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        *I.n_map[kind] = n; *I.n_sorted[kind] = 0; I.sh->n_map[I.slot[kind]] = n; I.sh->frame = 0;
        if (I.extract_err && *I.extract_err) atomicOr(&I.sh->err, (int)kErrRing);
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 f = I.feat[kind][i];
        Pt o;
        o.x = f.x; o.y = f.y; o.z = f.z; o.rgba = pack_rgba(0, 0, 0, 255);
        I.map[kind][i] = o;
    }
}

// Error words of everything that worked on this frame: the control blocks of the handle's three workspaces (main stream, forked grid
// build, overlapped down-sampling; sticky + live, see ws_error_bits) and the extractor's error word when the features came from one.
struct ErrSources { const unsigned* ctrl[3]; const unsigned* extract_err; };

__global__ void k_check_map_cap(const int* n_map0, const int* n_map1, int slot0, int slot1, int cap, OdomShared* sh, int last_pair,
                                ErrSources E) {
    PF_PDL_ENTRY();
    if (threadIdx.x != 0) return;
    if (*n_map0 > cap || *n_map1 > cap) atomicOr(&sh->err, 1);
    unsigned e = 0u;
    for (int i = 0; i < 3; ++i) if (E.ctrl[i]) e |= ws_error_bits(E.ctrl[i]);
    if (E.extract_err && *E.extract_err) e |= kErrRing;
    if (e) atomicOr(&sh->err, (int)e);
    sh->n_map[slot0] = *n_map0;
    sh->n_map[slot1] = *n_map1;
    if (last_pair) sh->frame += 1;     // the frame this block now describes
}

}  // namespace pf

using namespace pf;

// accessors implemented in extract.cu
void pf_extract_device_outputs(pf_extract* h, const float4** edge, const int** n_edge, const float4** surf, const int** n_surf,
                               cudaStream_t* stream, int* edge_cap, int* surf_cap, int* slot, const unsigned** err_word);
int pf_extract_enqueue_single(pf_extract* h, const float* xyzi, int n, int device_input, int want_label);
int pf_extract_enqueue_pre(pf_extract* h, const float* xyzi, int n, int device_input, const float4** src_out);
int pf_extract_enqueue_kernels(pf_extract* h, const float4* src, int want_label);
void pf_extract_count_launches(pf_extract* h, int n);
uint64_t pf_extract_uid(const pf_extract* h);

struct pf_odom {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev = nullptr, ev_done[2] = {nullptr, nullptr};   // ev_done[slot]: the odometry has consumed the extractor's output slot
    bool ev_done_set[2] = {false, false};
    pf_odom_params prm{};
    int fcap = 0, mcap = 0, bufcap = 0;
    // feature kinds and the pairs they are processed in
    int nk = 2;                          // live kinds: 2 (ES: edge, surf) or 3 (BPF: beam, pillar, facade)
    int type[kKinds] = {0, 1, 0, 0};     // 0 point-to-line, 1 point-to-plane
    int leaf_mul[kKinds] = {1, 2, 1, 1}; // leaf = map_resolution x this (:189-190, :658-660)
    int min_map[kKinds] = {10, 50, 0, 0};// guard thresholds (:247, :720)
    int npairs = 1;
    int pair[2][2] = {{0, 1}, {0, 0}};   // [pair] = {line-type kind, plane-type kind}; the null kind (kKinds - 1) stands in for "none"
    Workspace ws;
    // the search grids over the current maps do not depend on this frame's features: they are built on a second stream (own
    // workspace), forked at the start of the update and joined before the first association pass (PF_ODOM_FORK=0: in line)
    Workspace ws_grid;
    cudaStream_t stream_grid = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool fork_grid = true;
    float4* d_feat[kKinds] = {};
    int* d_nfeat = nullptr;              // [kKinds]
    Pt* d_ds[2][kKinds] = {};            // down-sampled feature clouds, double buffered like the maps (slot = map buffer of the update)
    int* d_nds[2] = {nullptr, nullptr};  // [slot] -> [kKinds]
    // The VoxelGrid down-sampling of a frame needs its features only, not the map: it runs on its own stream / workspace, so the
    // down-sampling of frame k+1 overlaps the pose solve and map update of frame k (PF_ODOM_OVERLAP=0 or timing taps: in line)
    Workspace ws_ds;
    cudaStream_t stream_ds = nullptr;
    cudaEvent_t ev_ds_done[2] = {nullptr, nullptr}, ev_upd_done[2] = {nullptr, nullptr}, ev_feat = nullptr;
    bool ev_upd_set[2] = {false, false};
    bool overlap_ds = true;
    Pt* d_map[2][kKinds] = {};           // [buffer][kind]
    int* d_nmap[2] = {nullptr, nullptr}; // [buffer] -> 2 * kKinds ints: n_map[kind], then n_sorted[kind] (merge.cuh: sorted prefix)
    MapMergeScratch msc{};
    bool sorted_known = false;           // false until the first update after init: the raw first-frame maps are unsorted
    int cur = 0;
    float4* d_gpts[kKinds] = {};
    int *d_cs[kKinds] = {}, *d_ce[kKinds] = {}, *d_geom = nullptr;
    int *d_head[kKinds] = {}, *d_hits[kKinds] = {}, *d_next[kKinds] = {}, *d_nn[kKinds] = {};
    uint8_t* d_flag[kKinds] = {};
    double* d_g8[kKinds] = {};
    float* d_wobs[kKinds] = {};          // residual weights (weightType 1 / 2 / 12)
    double* d_wspa[kKinds] = {};
    unsigned long long* d_wminmax = nullptr;   // [kKinds][4]
    LmState* d_state = nullptr;
    double* d_iter_poses = nullptr;
    OdomShared* d_sh = nullptr;
    double* d_pose_hist = nullptr;       // [kPoseHist][7] pose of every update (ring buffer)
    long long frame = 0;                 // frames processed (frame 0 = init)
    // pinned host mirrors
    OdomShared* h_sh = nullptr;
    LmState* h_state = nullptr;
    int* h_counts = nullptr;             // [16]
    double* h_iter = nullptr;            // [16*7]
    bool inited = false;
    const unsigned* extract_err = nullptr;   // error word of the extractor whose outputs feed this handle (null: host feature clouds)
    long long waited_upto = -1;          // newest frame whose result the caller has collected (pf_frame_wait / any blocking call)
    int optimization_count = 2;          // :198
    int last_passes = 0;
    // host-side upper bounds of the device-resident counts (launch geometry only; the kernels read the exact counts)
    int map_ub[kKinds] = {};
    static constexpr int kRing = 32;
    OdomShared* h_ring = nullptr;        // pinned [kRing]: asynchronous read-back of the shared block after every frame
    cudaEvent_t ring_ev[kRing] = {};
    long long ring_frame[kRing] = {};
    int ring_add[kRing][kKinds] = {};
    long long ring_head = 0;             // frames recorded in the ring
    long long known_frame = -1;          // newest frame whose exact map sizes have been folded into map_ub
    // steady-state replay: the launch sequence of one update (optimization_count == 2) captured as a CUDA graph per map buffer
    bool use_graph = true;               // PF_ODOM_GRAPH=0 disables
    cudaGraphExec_t graph_exec[2] = {nullptr, nullptr};
    const float4* graph_feat[2][kKinds] = {};
    const int* graph_nfeat[2][kKinds] = {};
    int graph_ub[2][kKinds] = {};        // feature upper bounds the graph was sized for
    int graph_mub[2][kKinds] = {};       // map upper bounds the graph was sized for
    uint64_t graph_launches[2] = {0, 0};
    bool graph_overlap[2] = {false, false};   // whether the graph was captured without the down-sampling (it ran on the second stream)
    int graph_captures = 0;
    // Front graph (pf_frame_submit): the extraction kernels and the down-sampling of one frame, captured once per (extractor output
    // slot, map buffer) and replayed on the extractor's stream.  With the update graph a queued frame costs the host a dozen driver
    // calls instead of ~25; a host thread that feeds several sequences spends its time in exactly those calls.  PF_FRAME_GRAPH=0: off.
    struct FrontGraph {
        cudaGraphExec_t exec = nullptr;
        uint64_t ex_uid = 0;             // the extractor it was captured with
        const float4* src = nullptr;
        int ub[2] = {0, 0};              // feature bounds the down-sampling was sized for
        int ex_launches = 0, ds_launches = 0;
    };
    FrontGraph front[2][2];
    cudaEvent_t ev_front_fork = nullptr, ev_front_join = nullptr;
    bool use_front = true;
    int front_captures = 0;
    int map_exact[kKinds] = {};          // last exactly known map sizes (read-backs); map_ub may run ahead of them when frames are queued
    int nds_known[kKinds] = {};          // down-sampled feature counts of the newest frame read back (0: none yet): launch geometry hint
    int graph_q[2][kKinds] = {};         // query-count hints the graph was sized for
    // optional phase timing (PF_ODOM_TIMING=1): CUDA events at the phase boundaries of the last update
    bool timing = false;
    cudaEvent_t tev[8] = {};
    float phase_ms[8] = {};
    int cap_of(int k) const { return k == kKinds - 1 ? 16 : 0; }   // the null kind gets token buffers
};

namespace {

constexpr int kNull = kKinds - 1;

int odom_alloc(pf_odom* h) {
    const int fcap = h->fcap, bufcap = h->bufcap;
    PF_CHECK(workspace_create(h->ws, 2 * bufcap, h->stream));
    PF_CHECK(workspace_create(h->ws_grid, 2 * bufcap, h->stream_grid));
    PF_CUDA(cudaMalloc(&h->d_nfeat, sizeof(int) * kKinds));
    PF_CHECK(workspace_create(h->ws_ds, 2 * fcap, h->stream_ds));
    for (int b = 0; b < 2; ++b) {
        PF_CUDA(cudaMalloc(&h->d_nds[b], sizeof(int) * kKinds));
        PF_CUDA(cudaMemset(h->d_nds[b], 0, sizeof(int) * kKinds));
    }
    PF_CUDA(cudaMemset(h->d_nfeat, 0, sizeof(int) * kKinds));
    PF_CUDA(cudaMalloc(&h->d_geom, sizeof(int) * 6 * kKinds));
    for (int b = 0; b < 2; ++b) {
        PF_CUDA(cudaMalloc(&h->d_nmap[b], sizeof(int) * 2 * kKinds));
        PF_CUDA(cudaMemset(h->d_nmap[b], 0, sizeof(int) * 2 * kKinds));
    }
    for (int k = 0; k < kKinds; ++k) {
        if (k >= h->nk && k != kNull) continue;
        const int fc = k == kNull ? 16 : fcap, bc = k == kNull ? 16 : bufcap;
        PF_CUDA(cudaMalloc(&h->d_feat[k], sizeof(float4) * fc));
        for (int b = 0; b < 2; ++b) PF_CUDA(cudaMalloc(&h->d_ds[b][k], sizeof(Pt) * fc));
        for (int b = 0; b < 2; ++b) PF_CUDA(cudaMalloc(&h->d_map[b][k], sizeof(Pt) * bc));
        PF_CUDA(cudaMalloc(&h->d_gpts[k], sizeof(float4) * bc));
        const size_t cells = k == kNull ? 16 : (size_t)kGridCellCap;
        PF_CUDA(cudaMalloc(&h->d_cs[k], sizeof(int) * cells));
        PF_CUDA(cudaMalloc(&h->d_ce[k], sizeof(int) * cells));
        PF_CUDA(cudaMalloc(&h->d_head[k], sizeof(int) * bc));
        PF_CUDA(cudaMalloc(&h->d_hits[k], sizeof(int) * bc));
        PF_CUDA(cudaMemset(h->d_head[k], 0xff, sizeof(int) * bc));
        PF_CUDA(cudaMemset(h->d_hits[k], 0, sizeof(int) * bc));
        PF_CUDA(cudaMalloc(&h->d_next[k], sizeof(int) * 5 * fc));
        PF_CUDA(cudaMalloc(&h->d_nn[k], sizeof(int) * 5 * fc));
        PF_CUDA(cudaMalloc(&h->d_flag[k], fc));
        PF_CUDA(cudaMemset(h->d_flag[k], 0, fc));
        PF_CUDA(cudaMalloc(&h->d_g8[k], sizeof(double) * 8 * fc));
        PF_CUDA(cudaMalloc(&h->d_wobs[k], sizeof(float) * fc));
        PF_CUDA(cudaMalloc(&h->d_wspa[k], sizeof(double) * fc));
    }
    PF_CUDA(cudaMalloc(&h->d_wminmax, sizeof(unsigned long long) * 4 * kKinds));
    PF_CUDA(cudaMemset(h->d_wminmax, 0, sizeof(unsigned long long) * 4 * kKinds));
    PF_CHECK(map_merge_scratch_create(h->msc, 2 * fcap + kMergeExcCap, kMergeExcCap, bufcap));
    PF_CUDA(cudaMalloc(&h->d_state, sizeof(LmState)));
    PF_CUDA(cudaMemset(h->d_state, 0, sizeof(LmState)));
    PF_CUDA(cudaMalloc(&h->d_iter_poses, sizeof(double) * 16 * 7));
    PF_CUDA(cudaMemset(h->d_iter_poses, 0, sizeof(double) * 16 * 7));
    PF_CUDA(cudaMalloc(&h->d_sh, sizeof(OdomShared)));
    PF_CUDA(cudaMalloc(&h->d_pose_hist, sizeof(double) * 7 * kPoseHist));
    PF_CUDA(cudaMemset(h->d_pose_hist, 0, sizeof(double) * 7 * kPoseHist));
    PF_CUDA(cudaMallocHost(&h->h_sh, sizeof(OdomShared)));
    PF_CUDA(cudaMallocHost(&h->h_state, sizeof(LmState)));
    PF_CUDA(cudaMallocHost(&h->h_counts, sizeof(int) * 16));
    PF_CUDA(cudaMallocHost(&h->h_iter, sizeof(double) * 16 * 7));
    PF_CUDA(cudaMallocHost(&h->h_ring, sizeof(OdomShared) * pf_odom::kRing));
    h->timing = getenv("PF_ODOM_TIMING") != nullptr;
    { const char* g = getenv("PF_ODOM_GRAPH"); h->use_graph = !(g && g[0] == '0') && !h->timing; }
    { const char* g = getenv("PF_ODOM_OVERLAP"); h->overlap_ds = !(g && g[0] == '0') && !h->timing; }
    if (h->timing) for (int i = 0; i < 8; ++i) PF_CUDA(cudaEventCreate(&h->tev[i]));
    for (int i = 0; i < pf_odom::kRing; ++i) PF_CUDA(cudaEventCreateWithFlags(&h->ring_ev[i], cudaEventDisableTiming));
    k_odom_reset<<<1, 32, 0, h->stream>>>(h->d_sh, h->d_state);
    PF_CUDA(cudaStreamSynchronize(h->stream));
    return PF_OK;
}

int upload_features(pf_odom* h, const float* const feat[], const int n[]) {
    for (int k = 0; k < h->nk; ++k) {
        PF_REQUIRE(n[k] >= 0 && (feat[k] || n[k] == 0), "bad feature arrays");
        PF_REQUIRE(n[k] <= h->fcap, "feature cloud of %d points exceeds max_features %d", n[k], h->fcap);
        h->h_counts[k] = n[k];
    }
    for (int k = h->nk; k < kKinds; ++k) h->h_counts[k] = 0;
    PF_CUDA(cudaMemcpyAsync(h->d_nfeat, h->h_counts, sizeof(int) * kKinds, cudaMemcpyHostToDevice, h->stream));
    for (int k = 0; k < h->nk; ++k)
        if (n[k]) PF_CUDA(cudaMemcpyAsync(h->d_feat[k], feat[k], sizeof(float4) * n[k], cudaMemcpyHostToDevice, h->stream));
    return PF_OK;
}

// asynchronous read-back of the shared block: lets later frames size their launches from exact map counts
int ring_record(pf_odom* h, const int add[kKinds]) {
    const int slot = (int)(h->ring_head % pf_odom::kRing);
    PF_CUDA(cudaMemcpyAsync(h->h_ring + slot, h->d_sh, sizeof(OdomShared), cudaMemcpyDeviceToHost, h->stream));
    PF_CUDA(cudaEventRecord(h->ring_ev[slot], h->stream));
    h->ring_frame[slot] = h->frame;
    for (int k = 0; k < kKinds; ++k) h->ring_add[slot][k] = add[k];
    h->ring_head += 1;
    return PF_OK;
}

void ring_refresh(pf_odom* h) {
    // newest completed read-back wins; frames enqueued after it contribute their (upper-bound) additions
    const long long lo = h->ring_head > pf_odom::kRing ? h->ring_head - pf_odom::kRing : 0;
    for (long long i = h->ring_head - 1; i >= lo; --i) {
        const int slot = (int)(i % pf_odom::kRing);
        if (h->ring_frame[slot] <= h->known_frame) break;
        if (cudaEventQuery(h->ring_ev[slot]) != cudaSuccess) continue;
        int ub[kKinds];
        for (int k = 0; k < kKinds; ++k) ub[k] = h->h_ring[slot].n_map[k];
        for (long long j = i + 1; j < h->ring_head; ++j) {
            const int sj = (int)(j % pf_odom::kRing);
            for (int k = 0; k < kKinds; ++k) ub[k] += h->ring_add[sj][k];
        }
        for (int k = 0; k < kKinds; ++k) { h->map_ub[k] = ub[k] < h->bufcap ? ub[k] : h->bufcap; h->map_exact[k] = h->h_ring[slot].n_map[k]; }
        if (h->ring_frame[slot] > 0) for (int k = 0; k < kKinds; ++k) h->nds_known[k] = h->h_ring[slot].n_ds[k];
        h->known_frame = h->ring_frame[slot];
        break;
    }
    cudaGetLastError();   // cudaErrorNotReady from the queries is not an error
}

// feat / n_feat / ub are indexed by kind slot; unused slots may be null
int enqueue_init(pf_odom* h, const float4* const feat[kKinds], const int* const n_feat[kKinds], const int ub[kKinds]) {
    for (int p = 0; p < h->npairs; ++p) {
        InitParams I{};
        for (int j = 0; j < 2; ++j) {
            const int k = h->pair[p][j];
            I.feat[j] = k == kNull ? h->d_feat[kNull] : feat[k];
            I.n_feat[j] = k == kNull ? h->d_nfeat + kNull : n_feat[k];
            I.map[j] = h->d_map[h->cur][k];
            I.n_map[j] = h->d_nmap[h->cur] + k; I.n_sorted[j] = h->d_nmap[h->cur] + kKinds + k;
            I.slot[j] = k;
        }
        I.map_cap = h->mcap;
        I.sh = h->d_sh;
        I.extract_err = h->extract_err;
        k_init_map<<<dim3(2 * kSMs, 2), 256, 0, h->stream>>>(I);
        h->ws.launches += 1;
    }
    PF_CUDA(cudaGetLastError());
    h->optimization_count = 12;   // :221
    h->inited = true;
    h->sorted_known = false;
    for (int k = 0; k < kKinds; ++k) h->map_ub[k] = k < h->nk ? (ub[k] < h->mcap ? ub[k] : h->mcap) : 0;
    h->frame = 0;
    const int none[kKinds] = {0, 0, 0, 0};
    PF_CHECK(ring_record(h, none));
    h->frame = 1;
    return PF_OK;
}

// VoxelGrid down-sampling of the frame's feature clouds into slot `slot`, leaf sizes as set by init (:189-190): setLeafSize takes
// floats.  Enqueued on the workspace's stream.
int record_downsample(pf_odom* h, Workspace& w, const float4* const feat[kKinds], const int* const n_feat[kKinds], const int ub[kKinds], int slot) {
    for (int p = 0; p < h->npairs; ++p) {
        VoxParams V{};
        V.mode = VOX_PCL;
        for (int j = 0; j < 2; ++j) {
            const int k = h->pair[p][j];
            V.c[j] = VoxCloud{reinterpret_cast<const Pt*>(feat[k]), n_feat[k], h->d_ds[slot][k], h->d_nds[slot] + k,
                              (float)(h->prm.map_resolution * h->leaf_mul[k]), 1};
        }
        PF_CHECK(voxelize(w, V, p, ub[h->pair[p][0]], ub[h->pair[p][1]]));
    }
    return PF_OK;
}

// The launch sequence of one update, enqueued on h->stream (directly, or into a stream capture).  ub / mub: upper bounds of the
// feature and map counts (launch geometry only: every kernel is grid-stride or persistent and reads the exact device counts).
// qh: how many down-sampled query points to expect per kind (from the read-backs of the last frames; <= ub): sizes the grids of the
// kernels that walk the query clouds.  ub is the only bound the host has when it enqueues (the raw feature count, 10 - 30 x the
// down-sampled one); grids sized by it are mostly CTAs that find nothing to do -- harmless for one sequence, but with several
// sequences on one GPU they take the slots other sequences' kernels could run in.
int record_update(pf_odom* h, const float4* const feat[kKinds], const int* const n_feat[kKinds], const int ub[kKinds], const int mub[kKinds],
                  const int qh[kKinds], int passes, bool sorted_known, int app[kKinds], bool overlap) {
    Workspace& ws = h->ws;
    const int cur = h->cur, nxt = cur ^ 1;
    auto mark = [&](int i) { if (h->timing) cudaEventRecord(h->tev[i], h->stream); };
    mark(0);
    PF_CHECK(workspace_begin_step(ws));
    {
        GuardSpec G{};
        G.count = h->nk;
        for (int k = 0; k < h->nk; ++k) { G.n_map[k] = h->d_nmap[cur] + k; G.min_pts[k] = h->min_map[k]; }
        PF_CUDA(launch_pdl(k_predict, dim3(1), dim3(32), 0, h->stream, h->d_sh, h->d_state, G));
        ws.launches += 1;
    }
    // search grids over the current maps: forked onto the second stream (captured as a parallel branch of the graph)
    auto grids = [&](Workspace& gw) -> int {
        for (int p = 0; p < h->npairs; ++p) {
            if (p > 0 || &gw != &ws) PF_CHECK(workspace_begin_step(gw));      // the second pair reuses the tickets and state slots
            GridBuild G{};
            for (int j = 0; j < 2; ++j) {
                const int k = h->pair[p][j];
                G.map[j] = h->d_map[cur][k]; G.n_map[j] = h->d_nmap[cur] + k; G.pts[j] = h->d_gpts[k];
                G.cell_start[j] = h->d_cs[k]; G.cell_end[j] = h->d_ce[k]; G.geom[j] = h->d_geom + 6 * k;
            }
            PF_CHECK(build_grids(gw, G, 2, mub[h->pair[p][0]], mub[h->pair[p][1]]));
        }
        return PF_OK;
    };
    if (h->fork_grid) {
        PF_CUDA(cudaEventRecord(h->ev_fork, h->stream));
        PF_CUDA(cudaStreamWaitEvent(h->stream_grid, h->ev_fork, 0));
        const uint64_t l0 = h->ws_grid.launches;
        PF_CHECK(grids(h->ws_grid));
        ws.launches += h->ws_grid.launches - l0;
        PF_CUDA(cudaEventRecord(h->ev_join, h->stream_grid));
    }
    if (!overlap) PF_CHECK(record_downsample(h, ws, feat, n_feat, ub, cur));
    mark(1);
    if (h->fork_grid) PF_CUDA(cudaStreamWaitEvent(h->stream, h->ev_join, 0));
    else PF_CHECK(grids(ws));
    mark(2);
    // optimisation passes
    AssocParams A[2]{};
    LmParams L{};
    for (int p = 0; p < h->npairs; ++p) {
        for (int j = 0; j < 2; ++j) {
            const int k = h->pair[p][j];
            A[p].c[j] = AssocCloud{h->d_ds[cur][k], h->d_nds[cur] + k, h->d_map[cur][k], h->d_nmap[cur] + k,
                                   KnnGrid{h->d_gpts[k], h->d_cs[k], h->d_ce[k], h->d_geom + 6 * k},
                                   h->d_head[k], h->d_hits[k], h->d_next[k], h->d_nn[k], h->d_flag[k], h->d_g8[k], h->d_wobs[k], h->d_wspa[k]};
        }
        A[p].pose = h->d_state->x;
        A[p].k_new = h->prm.k_new; A[p].theta_p = h->prm.theta_p; A[p].theta_max = h->prm.theta_max;
        A[p].guard = &h->d_sh->guard;
        A[p].weight_type = (int)h->prm.weight_type;
        A[p].w_minmax = h->d_wminmax + 8 * p;           // [2][4] of this pair
    }
    // residual blocks in the order the reference adds them: kinds 0, 1, (2)
    int ub_src[kLmMaxSrc] = {0, 0, 0};
    L.nsrc = h->nk;
    for (int k = 0; k < h->nk; ++k) {
        int pp = 0, jj = 0;
        for (int p = 0; p < h->npairs; ++p)
            for (int j = 0; j < 2; ++j)
                if (h->pair[p][j] == k) { pp = p; jj = j; }
        L.src[k] = ResidualSrc{h->type[k], h->d_ds[cur][k], nullptr, h->d_flag[k], h->d_g8[k], h->d_nds[cur] + k, h->d_wobs[k], h->d_wspa[k],
                               h->d_wminmax + 8 * pp + 4 * jj};
        ub_src[k] = ub[k];
    }
    L.state = h->d_state; L.iter_poses = h->d_iter_poses; L.eval_only = 0;
    L.weight_type = (int)h->prm.weight_type;
    for (int it = 0; it < passes; ++it) {
        for (int p = 0; p < h->npairs; ++p) PF_CHECK(associate_pass(h->stream, A[p], qh[h->pair[p][0]], qh[h->pair[p][1]], &ws.launches));
        if (it == passes - 1) mark(3);
        PF_CHECK(lm_solve(h->stream, L, nullptr, it == 0, &ws.launches, ub_src));
    }
    mark(4);
    // append + map maintenance
    const float mres = (float)h->prm.map_resolution;   // rgbds(tmp, map_resolution [* 2]) with the float member map_resolution (:625-626)
    for (int k = 0; k < kKinds; ++k) app[k] = 0;
    for (int p = 0; p < h->npairs; ++p) {
        AppendParams P{};
        for (int j = 0; j < 2; ++j) {
            const int k = h->pair[p][j];
            P.ds[j] = h->d_ds[cur][k]; P.n_ds[j] = h->d_nds[cur] + k; P.map[j] = h->d_map[cur][k]; P.n_map[j] = h->d_nmap[cur] + k; P.slot[j] = k;
        }
        P.write_pose = p == 0;
        P.sh = h->d_sh; P.S = h->d_state; P.map_cap = h->bufcap;
        P.pose_hist = h->d_pose_hist;
        const int qmax = qh[h->pair[p][0]] > qh[h->pair[p][1]] ? qh[h->pair[p][0]] : qh[h->pair[p][1]];
        const int ablk = div_up(qmax > 0 ? qmax : 1, 256);
        PF_CUDA(launch_pdl(k_append, dim3(ablk < kSMs ? ablk : kSMs, 2), dim3(256), 0, h->stream, P));
        ws.launches += 1;
    }
    for (int p = 0; p < h->npairs; ++p) {
        if (p > 0) PF_CHECK(workspace_begin_step(ws));
        const int ka = h->pair[p][0], kb = h->pair[p][1];
        MapMergeParams M{};
        for (int j = 0; j < 2; ++j) {
            const int k = h->pair[p][j];
            M.c[j] = MapMergeCloud{h->d_map[cur][k], h->d_nmap[cur] + kKinds + k, h->d_sh->n_app + k, h->d_map[nxt][k], h->d_nmap[nxt] + k,
                                   h->d_nmap[nxt] + kKinds + k, mres * (float)h->leaf_mul[k]};
        }
        M.center = h->d_sh->odom.t;
        M.k_new = h->prm.k_new; M.theta_p = h->prm.theta_p; M.theta_max = h->prm.theta_max;
        M.s = h->msc;
        int capb[2], capa[2];
        for (int j = 0; j < 2; ++j) {
            const int k = h->pair[p][j];
            app[k] = mub[k] + ub[k] < h->bufcap ? mub[k] + ub[k] : h->bufcap;
            // unsorted part: everything on the first update (raw first-frame maps), later last update's exceptions + this frame's points
            capb[j] = sorted_known ? (kMergeExcCap + qh[k] < app[k] ? kMergeExcCap + qh[k] : app[k]) : app[k];   // launch geometry only
            if (k == kNull) capb[j] = 0;
            capa[j] = mub[k];
        }
        PF_CHECK(map_merge(ws, M, capb[0], capb[1], capa[0], capa[1]));
        ErrSources E{};
        E.ctrl[0] = ws.ctrl;
        E.ctrl[1] = h->fork_grid ? h->ws_grid.ctrl : nullptr;
        E.ctrl[2] = overlap ? h->ws_ds.ctrl : nullptr;
        E.extract_err = h->extract_err;
        PF_CUDA(launch_pdl(k_check_map_cap, dim3(1), dim3(32), 0, h->stream, h->d_nmap[nxt] + ka, h->d_nmap[nxt] + kb, ka, kb, h->mcap, h->d_sh, p == h->npairs - 1 ? 1 : 0, E));
        ws.launches += 1;
    }
    PF_CUDA(cudaGetLastError());
    mark(5);
    return PF_OK;
}

// feat_ready: event after which the feature clouds may be read (null: they are ordered on h->stream already)
// overlap: down-sample on the second stream (frames queued back to back); a caller that blocks on every frame gains nothing from it
// and saves the stream hops by keeping it in line
// ds_done: the down-sampling of this frame has already been enqueued (front graph), ev_ds_done[cur] marks its end
int enqueue_update(pf_odom* h, const float4* const feat_in[kKinds], const int* const n_feat_in[kKinds], const int ub_in[kKinds],
                   cudaEvent_t feat_ready, bool overlap, bool ds_done = false) {
    overlap = overlap && h->overlap_ds;
    if (h->optimization_count > 2) h->optimization_count--;   // :232-233
    ring_refresh(h);
    const float4* feat[kKinds];
    const int* n_feat[kKinds];
    int ub[kKinds], mub[kKinds], app[kKinds];
    for (int k = 0; k < kKinds; ++k) {
        const bool live = k < h->nk;
        feat[k] = live ? feat_in[k] : h->d_feat[kNull];
        n_feat[k] = live ? n_feat_in[k] : h->d_nfeat + kNull;
        ub[k] = live ? ub_in[k] : 0;
        if (ub[k] > h->fcap) ub[k] = h->fcap;
        if (live && ub[k] < 1) ub[k] = 1;
        mub[k] = live ? (h->map_ub[k] > 1 ? h->map_ub[k] : 1) : 0;
    }
    int qh[kKinds];
    for (int k = 0; k < kKinds; ++k) {
        const int seen = h->nds_known[k];
        qh[k] = seen > 0 && seen + seen / 4 + 512 < ub[k] ? seen + seen / 4 + 512 : ub[k];
    }
    const int passes = h->optimization_count;
    h->last_passes = passes;
    const int cur = h->cur;
    if (overlap && ds_done) {
        PF_CUDA(cudaStreamWaitEvent(h->stream, h->ev_ds_done[cur], 0));
    } else if (overlap) {
        // down-sampling on its own stream: after the features are there and after the update that last read this slot
        if (!feat_ready) { PF_CUDA(cudaEventRecord(h->ev_feat, h->stream)); feat_ready = h->ev_feat; }
        PF_CUDA(cudaStreamWaitEvent(h->stream_ds, feat_ready, 0));
        if (h->ev_upd_set[cur]) PF_CUDA(cudaStreamWaitEvent(h->stream_ds, h->ev_upd_done[cur], 0));
        const uint64_t l0 = h->ws_ds.launches;
        PF_CHECK(workspace_begin_step(h->ws_ds));
        PF_CHECK(record_downsample(h, h->ws_ds, feat, n_feat, ub, cur));
        h->ws.launches += h->ws_ds.launches - l0;
        PF_CUDA(cudaEventRecord(h->ev_ds_done[cur], h->stream_ds));
        PF_CUDA(cudaStreamWaitEvent(h->stream, h->ev_ds_done[cur], 0));
    }
    bool replayed = false;
    if (h->use_graph && passes == 2 && h->sorted_known) {
        // steady state: replay the captured launch sequence of this map buffer; (re)capture when the inputs moved or outgrew its
        // sizing.  The sizing only sets grid dimensions (the kernels are grid-stride and read exact device counts), so the test uses
        // the last exactly known map sizes, not the host's running upper bound, which runs ahead while frames are queued.
        bool fits = h->graph_exec[cur] != nullptr && h->graph_overlap[cur] == overlap;
        for (int k = 0; k < kKinds && fits; ++k)
            fits = h->graph_feat[cur][k] == feat[k] && h->graph_nfeat[cur][k] == n_feat[k] && ub[k] <= h->graph_ub[cur][k] &&
                   h->map_exact[k] <= h->graph_mub[cur][k] && qh[k] <= h->graph_q[cur][k] &&
                   h->graph_q[cur][k] <= 4 * qh[k] + 8192;      // captured before the first read-back: sized by the raw bound
        if (!fits) {
            int gub[kKinds], gmub[kKinds], gq[kKinds];
            for (int k = 0; k < kKinds; ++k) {
                const bool live = k < h->nk;
                gub[k] = live ? (ub[k] + ub[k] / 4 + 1024 < h->fcap ? ub[k] + ub[k] / 4 + 1024 : h->fcap) : 0;   // headroom: scans differ in size
                gq[k] = qh[k] + qh[k] / 4 + 1024 < gub[k] ? qh[k] + qh[k] / 4 + 1024 : gub[k];
                const long long want = 2ll * h->map_exact[k] + 65536;
                gmub[k] = live ? (int)(want < h->bufcap ? want : h->bufcap) : 0;
            }
            if (h->graph_exec[cur]) { cudaGraphExecDestroy(h->graph_exec[cur]); h->graph_exec[cur] = nullptr; }
            cudaGraph_t graph = nullptr;
            const uint64_t l0 = h->ws.launches;
            bool ok = cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
            if (ok) {
                const int rc = record_update(h, feat, n_feat, gub, gmub, gq, passes, true, app, overlap);
                ok = cudaStreamEndCapture(h->stream, &graph) == cudaSuccess && rc == PF_OK && graph != nullptr;
            }
            if (ok) ok = cudaGraphInstantiate(&h->graph_exec[cur], graph, 0) == cudaSuccess;
            if (graph) cudaGraphDestroy(graph);
            h->graph_launches[cur] = h->ws.launches - l0;
            h->ws.launches = l0;
            h->graph_captures += 1;
            if (!ok) {          // capture not possible on this driver: run the plain launch sequence from now on
                cudaGetLastError();
                h->graph_exec[cur] = nullptr;
                h->use_graph = false;
            } else {
                h->graph_overlap[cur] = overlap;
                for (int k = 0; k < kKinds; ++k) {
                    h->graph_feat[cur][k] = feat[k]; h->graph_nfeat[cur][k] = n_feat[k];
                    h->graph_ub[cur][k] = gub[k]; h->graph_mub[cur][k] = gmub[k]; h->graph_q[cur][k] = gq[k];
                }
            }
        }
        if (h->graph_exec[cur]) {
            PF_CUDA(cudaGraphLaunch(h->graph_exec[cur], h->stream));
            h->ws.launches += h->graph_launches[cur];
            for (int k = 0; k < kKinds; ++k) app[k] = mub[k] + ub[k] < h->bufcap ? mub[k] + ub[k] : h->bufcap;
            replayed = true;
        }
    }
    if (!replayed) PF_CHECK(record_update(h, feat, n_feat, ub, mub, qh, passes, h->sorted_known, app, overlap));
    PF_CUDA(cudaEventRecord(h->ev_upd_done[cur], h->stream));
    h->ev_upd_set[cur] = true;
    h->sorted_known = true;
    h->cur = cur ^ 1;
    for (int k = 0; k < kKinds; ++k) h->map_ub[k] = app[k];     // the update never grows a map beyond old + appended
    PF_CHECK(ring_record(h, ub));
    h->frame += 1;
    return PF_OK;
}


// device error bits of a frame -> status + text
int decode_frame_error(const pf_odom* h, int err) {
    if (err == 0) return PF_OK;
    if (err & (int)kErrPose) {
        // the reference has no such check: once the maps run empty its constant-velocity prediction (:235-240) doubles every frame
        set_error("the pose is no longer finite: tracking was lost (the reference would go on publishing NaN here)");
        return PF_ERR_STATE;
    }
    if (err & 1) set_error("local map exceeded max_map_points = %d", h->mcap);
    else if (err & (int)kErrMergeMask)
        set_error("map update failed (bits %d: 2 = voxel coordinates outside the key range, 4 = more than %d centroids left their voxel, 8 / 16 = internal)",
                  err & (int)kErrMergeMask, kMergeExcCap);
    else if (err & (int)kErrGrid)
        set_error("local map extent exceeds the search grid (%d cells of 1 m): the matches of this frame are void", kGridCellCap);
    else if (err & (int)kErrVoxel) set_error("VoxelGrid index space overflow while down-sampling the frame's features (extent / leaf too large)");
    else if (err & (int)kErrRing) set_error("a scan ring held more points than the extractor's max_ring_points: the ring's features are missing");
    else set_error("device error bits %d", err);
    return PF_ERR_CAPACITY;
}

int finish_frame(pf_odom* h, double pose_out[7]) {
    PF_CUDA(cudaMemcpyAsync(h->h_sh, h->d_sh, sizeof(OdomShared), cudaMemcpyDeviceToHost, h->stream));
    PF_CUDA(cudaStreamSynchronize(h->stream));
    h->waited_upto = h->frame - 1;
    PF_CHECK(decode_frame_error(h, h->h_sh->err));
    if (pose_out) memcpy(pose_out, h->h_sh->pose, sizeof(double) * 7);
    if (h->inited) {
        for (int k = 0; k < kKinds; ++k) { h->map_ub[k] = h->h_sh->n_map[k]; h->map_exact[k] = h->h_sh->n_map[k]; }
        if (h->frame - 1 > 0) for (int k = 0; k < kKinds; ++k) h->nds_known[k] = h->h_sh->n_ds[k];
        h->known_frame = h->frame - 1;
    }
    return PF_OK;
}

int odom_create(const pf_odom_params* p, int device, bool bpf, pf_odom** out) {
    PF_REQUIRE(p && out, "null argument");
    PF_REQUIRE(p->map_resolution >= 0.2, "map_resolution %g: the streaming map update keeps 10-bit voxel coordinates inside the 200 m crop box (needs >= 0.2 m)",
               p->map_resolution);
    PF_REQUIRE(p->weight_type == 0.0 || p->weight_type == 1.0 || p->weight_type == 2.0 || p->weight_type == 12.0,
               "weight_type %g: the reference knows 0, 1, 2 and 12 (src/odomEstimationClass.cpp:405-421)", p->weight_type);
    int ndev = 0;
    PF_CUDA(cudaGetDeviceCount(&ndev));
    PF_REQUIRE(device >= 0 && device < ndev, "device %d not available (%d devices)", device, ndev);
    PF_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    PF_CUDA(cudaGetDeviceProperties(&prop, device));
    PF_REQUIRE(prop.major == 10, "pfilter_b200 needs an sm_100a device, found sm_%d%d", prop.major, prop.minor);
    pf_odom* h = new pf_odom();
    odom_handles_changed(+1);
    h->device = device;
    h->prm = *p;
    if (bpf) {      // beam, pillar: line residuals at map_resolution; facade: plane residuals at 2 x map_resolution (:658-660)
        h->nk = 3;
        h->type[0] = 0; h->type[1] = 0; h->type[2] = 1;
        h->leaf_mul[0] = 1; h->leaf_mul[1] = 1; h->leaf_mul[2] = 2;
        h->min_map[0] = 10; h->min_map[1] = 10; h->min_map[2] = 50;     // :720
        h->npairs = 2;
        h->pair[0][0] = 0; h->pair[0][1] = 2;
        h->pair[1][0] = 1; h->pair[1][1] = kNull;
    } else {
        h->pair[0][0] = 0; h->pair[0][1] = 1;
    }
    h->fcap = p->max_features > 0 ? p->max_features : 131072;
    h->mcap = p->max_map_points > 0 ? p->max_map_points : (2 << 20);
    if (h->mcap < h->fcap) h->mcap = h->fcap;   // the raw first-frame clouds become the maps (:217-222)
    h->bufcap = h->mcap + h->fcap;
    PF_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    PF_CUDA(cudaStreamCreateWithFlags(&h->stream_grid, cudaStreamNonBlocking));
    PF_CUDA(cudaStreamCreateWithFlags(&h->stream_ds, cudaStreamNonBlocking));
    PF_CUDA(cudaEventCreateWithFlags(&h->ev_feat, cudaEventDisableTiming));
    PF_CUDA(cudaEventCreateWithFlags(&h->ev_front_fork, cudaEventDisableTiming));
    PF_CUDA(cudaEventCreateWithFlags(&h->ev_front_join, cudaEventDisableTiming));
    { const char* g = getenv("PF_FRAME_GRAPH"); h->use_front = !(g && g[0] == '0'); }
    for (int b = 0; b < 2; ++b) {
        PF_CUDA(cudaEventCreateWithFlags(&h->ev_ds_done[b], cudaEventDisableTiming));
        PF_CUDA(cudaEventCreateWithFlags(&h->ev_upd_done[b], cudaEventDisableTiming));
    }
    PF_CUDA(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    PF_CUDA(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
    { const char* e = getenv("PF_ODOM_FORK"); if (e && atoi(e) == 0) h->fork_grid = false; }
    PF_CUDA(cudaEventCreateWithFlags(&h->ev, cudaEventDisableTiming));
    for (int b = 0; b < 2; ++b) PF_CUDA(cudaEventCreateWithFlags(&h->ev_done[b], cudaEventDisableTiming));
    int rc = odom_alloc(h);
    if (rc != PF_OK) { pf_odom_destroy(h); return rc; }
    *out = h;
    return PF_OK;
}

}  // namespace

extern "C" int pf_odom_create(const pf_odom_params* p, int device, pf_odom** out) { return odom_create(p, device, false, out); }
extern "C" int pf_odom_bpf_create(const pf_odom_params* p, int device, pf_odom** out) { return odom_create(p, device, true, out); }

extern "C" int pf_odom_destroy(pf_odom* h) {
    if (!h) return PF_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    for (int b = 0; b < 2; ++b) if (h->graph_exec[b]) cudaGraphExecDestroy(h->graph_exec[b]);
    workspace_destroy(h->ws);
    workspace_destroy(h->ws_grid);
    workspace_destroy(h->ws_ds);
    if (h->stream_ds) cudaStreamDestroy(h->stream_ds);
    if (h->ev_feat) cudaEventDestroy(h->ev_feat);
    if (h->ev_front_fork) cudaEventDestroy(h->ev_front_fork);
    if (h->ev_front_join) cudaEventDestroy(h->ev_front_join);
    for (int a = 0; a < 2; ++a) for (int b = 0; b < 2; ++b) if (h->front[a][b].exec) cudaGraphExecDestroy(h->front[a][b].exec);
    for (int b = 0; b < 2; ++b) { if (h->ev_ds_done[b]) cudaEventDestroy(h->ev_ds_done[b]); if (h->ev_upd_done[b]) cudaEventDestroy(h->ev_upd_done[b]); }
    if (h->stream_grid) cudaStreamDestroy(h->stream_grid);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    map_merge_scratch_destroy(h->msc);
    cudaFree(h->d_nfeat); cudaFree(h->d_nds[0]); cudaFree(h->d_nds[1]); cudaFree(h->d_geom); cudaFree(h->d_nmap[0]); cudaFree(h->d_nmap[1]);
    for (int k = 0; k < kKinds; ++k) {
        cudaFree(h->d_feat[k]); cudaFree(h->d_ds[0][k]); cudaFree(h->d_ds[1][k]); cudaFree(h->d_map[0][k]); cudaFree(h->d_map[1][k]);
        cudaFree(h->d_gpts[k]); cudaFree(h->d_cs[k]); cudaFree(h->d_ce[k]); cudaFree(h->d_head[k]); cudaFree(h->d_hits[k]);
        cudaFree(h->d_next[k]); cudaFree(h->d_nn[k]); cudaFree(h->d_flag[k]); cudaFree(h->d_g8[k]);
        cudaFree(h->d_wobs[k]); cudaFree(h->d_wspa[k]);
    }
    cudaFree(h->d_wminmax);
    cudaFree(h->d_state); cudaFree(h->d_iter_poses); cudaFree(h->d_sh); cudaFree(h->d_pose_hist);
    cudaFreeHost(h->h_sh); cudaFreeHost(h->h_state); cudaFreeHost(h->h_counts); cudaFreeHost(h->h_iter); cudaFreeHost(h->h_ring);
    for (int i = 0; i < pf_odom::kRing; ++i) if (h->ring_ev[i]) cudaEventDestroy(h->ring_ev[i]);
    if (h->ev) cudaEventDestroy(h->ev);
    for (int b = 0; b < 2; ++b) if (h->ev_done[b]) cudaEventDestroy(h->ev_done[b]);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    odom_handles_changed(-1);
    return PF_OK;
}

namespace {
void set_extract_err(pf_odom* h, const unsigned* w) {
    if (h->extract_err == w) return;
    for (int b = 0; b < 2; ++b)      // the error word is baked into the captured graphs
        if (h->graph_exec[b]) { cudaGraphExecDestroy(h->graph_exec[b]); h->graph_exec[b] = nullptr; }
    h->extract_err = w;
}
// host feature clouds -> device, then init / update
int host_frame(pf_odom* h, int nk_expected, const float* const feat[], const int n[], bool init, double* pose_out) {
    PF_REQUIRE(h, "null handle");
    PF_REQUIRE(h->nk == nk_expected, "this handle was created for %d feature kinds, the call passes %d", h->nk, nk_expected);
    if (!init && !h->inited) { set_error("update before init_map"); return PF_ERR_STATE; }
    PF_CUDA(cudaSetDevice(h->device));
    set_extract_err(h, nullptr);
    PF_CHECK(upload_features(h, feat, n));
    const float4* df[kKinds];
    const int* dn[kKinds];
    int ub[kKinds];
    for (int k = 0; k < kKinds; ++k) { df[k] = h->d_feat[k]; dn[k] = h->d_nfeat + k; ub[k] = k < h->nk ? n[k] : 0; }
    if (init) { PF_CHECK(enqueue_init(h, df, dn, ub)); }
    else { PF_CHECK(enqueue_update(h, df, dn, ub, nullptr, false)); }
    return finish_frame(h, pose_out);
}
}  // namespace

extern "C" int pf_odom_init_map(pf_odom* h, const float* edge, int n_edge, const float* surf, int n_surf) {
    const float* f[2] = {edge, surf};
    const int n[2] = {n_edge, n_surf};
    return host_frame(h, 2, f, n, true, nullptr);
}

extern "C" int pf_odom_update(pf_odom* h, const float* edge, int n_edge, const float* surf, int n_surf, double pose_out[7]) {
    const float* f[2] = {edge, surf};
    const int n[2] = {n_edge, n_surf};
    return host_frame(h, 2, f, n, false, pose_out);
}

// Odom_BPF_EstimationClass::initMapWithPoints (:692-698) / updatePointsToMap (:706-760)
extern "C" int pf_odom_bpf_init_map(pf_odom* h, const float* beam, int n_beam, const float* pillar, int n_pillar, const float* facade, int n_facade) {
    const float* f[3] = {beam, pillar, facade};
    const int n[3] = {n_beam, n_pillar, n_facade};
    return host_frame(h, 3, f, n, true, nullptr);
}

extern "C" int pf_odom_bpf_update(pf_odom* h, const float* beam, int n_beam, const float* pillar, int n_pillar, const float* facade, int n_facade,
                                  double pose_out[7]) {
    const float* f[3] = {beam, pillar, facade};
    const int n[3] = {n_beam, n_pillar, n_facade};
    return host_frame(h, 3, f, n, false, pose_out);
}

static int process_extracted(pf_odom* h, pf_extract* ex, double pose_out[7], bool sync) {
    PF_REQUIRE(h && ex, "null handle");
    PF_REQUIRE(h->nk == 2, "the fused extraction hand-off feeds the edge / surf odometry");
    PF_CUDA(cudaSetDevice(h->device));
    const float4* feat[kKinds] = {};
    const int* nf[kKinds] = {};
    int ub[kKinds] = {0, 0, 0, 0};
    cudaStream_t exs;
    int slot = 0;
    const unsigned* exerr = nullptr;
    pf_extract_device_outputs(ex, &feat[0], &nf[0], &feat[1], &nf[1], &exs, &ub[0], &ub[1], &slot, &exerr);
    set_extract_err(h, exerr);
    PF_REQUIRE(ub[1] <= h->fcap, "scan of %d points exceeds max_features %d", ub[1], h->fcap);
    PF_CUDA(cudaEventRecord(h->ev, exs));
    cudaStream_t reader = h->stream;       // the stream that reads the extractor's outputs
    if (!h->inited) {
        PF_CUDA(cudaStreamWaitEvent(h->stream, h->ev, 0));
        PF_CHECK(enqueue_init(h, feat, nf, ub));
    } else {
        const bool overlap = h->overlap_ds && !sync;
        if (overlap) reader = h->stream_ds;            // only the down-sampling reads them
        else PF_CUDA(cudaStreamWaitEvent(h->stream, h->ev, 0));
        PF_CHECK(enqueue_update(h, feat, nf, ub, h->ev, overlap));
    }
    // The extractor's outputs are double buffered: the next extraction writes the OTHER slot, so it may run while this frame
    // still reads this one; it only has to wait for the reader of that other slot (the previous frame).
    PF_CUDA(cudaEventRecord(h->ev_done[slot], reader));
    h->ev_done_set[slot] = true;
    if (h->ev_done_set[slot ^ 1]) PF_CUDA(cudaStreamWaitEvent(exs, h->ev_done[slot ^ 1], 0));
    return sync ? finish_frame(h, pose_out) : PF_OK;
}

// pf_frame_submit in steady state: `pre` (count + H2D copy) has been enqueued on the extractor's stream; extraction kernels and
// down-sampling follow as one graph launch on that stream, the update as the other on the odometry's.
static int process_front_graph(pf_odom* h, pf_extract* ex, const float4* src) {
    PF_CUDA(cudaSetDevice(h->device));
    const float4* feat[kKinds] = {};
    const int* nf[kKinds] = {};
    int ub[kKinds] = {0, 0, 0, 0};
    cudaStream_t exs;
    int slot = 0;
    const unsigned* exerr = nullptr;
    pf_extract_device_outputs(ex, &feat[0], &nf[0], &feat[1], &nf[1], &exs, &ub[0], &ub[1], &slot, &exerr);
    set_extract_err(h, exerr);
    PF_REQUIRE(ub[1] <= h->fcap, "scan of %d points exceeds max_features %d", ub[1], h->fcap);
    const int cur = h->cur;
    // this frame's extraction writes output slot `slot` (last read by the frame before last, on this stream unless that frame took
    // the plain path) and its down-sampling writes d_ds[cur] (last read by the update of the frame before last)
    if (h->ev_done_set[slot]) { PF_CUDA(cudaStreamWaitEvent(exs, h->ev_done[slot], 0)); h->ev_done_set[slot] = false; }
    if (h->ev_upd_set[cur]) PF_CUDA(cudaStreamWaitEvent(exs, h->ev_upd_done[cur], 0));
    pf_odom::FrontGraph& G = h->front[slot][cur];
    if (!(G.exec && G.ex_uid == pf_extract_uid(ex) && G.src == src && ub[0] <= G.ub[0] && ub[1] <= G.ub[1])) {
        if (G.exec) { cudaGraphExecDestroy(G.exec); G.exec = nullptr; }
        int gub[kKinds] = {0, 0, 0, 0};
        for (int k = 0; k < 2; ++k) gub[k] = ub[k] + ub[k] / 4 + 1024 < h->fcap ? ub[k] + ub[k] / 4 + 1024 : h->fcap;
        cudaGraph_t graph = nullptr;
        const uint64_t l0 = h->ws_ds.launches;
        bool ok = cudaStreamBeginCapture(exs, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
        if (ok) {
            int rc = pf_extract_enqueue_kernels(ex, src, 0);
            pf_extract_count_launches(ex, -3);                    // counted per replay below
            if (rc == PF_OK) rc = cudaEventRecord(h->ev_front_fork, exs) == cudaSuccess ? PF_OK : PF_ERR_CUDA;
            if (rc == PF_OK) rc = cudaStreamWaitEvent(h->stream_ds, h->ev_front_fork, 0) == cudaSuccess ? PF_OK : PF_ERR_CUDA;
            if (rc == PF_OK) rc = workspace_begin_step(h->ws_ds);
            if (rc == PF_OK) rc = record_downsample(h, h->ws_ds, feat, nf, gub, cur);
            if (rc == PF_OK) rc = cudaEventRecord(h->ev_front_join, h->stream_ds) == cudaSuccess ? PF_OK : PF_ERR_CUDA;
            if (rc == PF_OK) rc = cudaStreamWaitEvent(exs, h->ev_front_join, 0) == cudaSuccess ? PF_OK : PF_ERR_CUDA;
            ok = cudaStreamEndCapture(exs, &graph) == cudaSuccess && rc == PF_OK && graph != nullptr;
        }
        if (ok) ok = cudaGraphInstantiate(&G.exec, graph, 0) == cudaSuccess;
        if (graph) cudaGraphDestroy(graph);
        G.ds_launches = (int)(h->ws_ds.launches - l0);
        h->ws_ds.launches = l0;
        h->front_captures += 1;
        if (!ok) {          // not capturable here: the plain launch sequence from now on
            cudaGetLastError();
            G.exec = nullptr;
            h->use_front = false;
            PF_CHECK(pf_extract_enqueue_kernels(ex, src, 0));
            return process_extracted(h, ex, nullptr, false);
        }
        G.ex_uid = pf_extract_uid(ex); G.src = src; G.ub[0] = gub[0]; G.ub[1] = gub[1]; G.ex_launches = 3;
    }
    PF_CUDA(cudaGraphLaunch(G.exec, exs));
    pf_extract_count_launches(ex, G.ex_launches);
    h->ws_ds.launches += (uint64_t)G.ds_launches;
    h->ws.launches += (uint64_t)G.ds_launches;
    PF_CUDA(cudaEventRecord(h->ev_ds_done[cur], exs));
    return enqueue_update(h, feat, nf, ub, nullptr, true, true);
}

extern "C" int pf_odom_process_extracted(pf_odom* h, pf_extract* ex, double pose_out[7]) {
    return process_extracted(h, ex, pose_out, true);
}

extern "C" int pf_frame_process(pf_extract* ex, pf_odom* od, const float* xyzi, int n, double pose_out[7]) {
    PF_REQUIRE(ex && od, "null handle");
    PF_CHECK(pf_extract_enqueue_single(ex, xyzi, n, 0, 0));
    return process_extracted(od, ex, pose_out, true);
}

extern "C" int pf_frame_process_device(pf_extract* ex, pf_odom* od, const void* d_xyzi, int n, double* pose_out) {
    PF_REQUIRE(ex && od, "null handle");
    PF_CHECK(pf_extract_enqueue_single(ex, (const float*)d_xyzi, n, 1, 0));
    return process_extracted(od, ex, pose_out, pose_out != nullptr);
}

// Pipelined form of pf_frame_process: submit returns as soon as the frame is enqueued (H2D copy, extraction, odometry), wait
// blocks until that frame is done and hands out its pose.  A caller that submits frame k+1 before waiting for frame k overlaps the
// upload and the extraction of the next scan with the odometry of the current one -- what the reference gets from running its
// nodes as separate processes.  The scan buffer must stay valid (and should be pinned) until the frame has been waited for.
extern "C" int pf_frame_submit(pf_extract* ex, pf_odom* od, const float* xyzi, int n, long long* frame_id) {
    PF_REQUIRE(ex && od, "null handle");
    if (od->frame - 1 - od->waited_upto >= pf_odom::kRing) {     // the result slot of the oldest uncollected frame would be overwritten
        set_error("%d frames are outstanding: pf_frame_wait for frame %lld first", pf_odom::kRing, od->waited_upto + 1);
        return PF_ERR_STATE;
    }
    if (od->inited && od->use_front && od->use_graph && od->overlap_ds && od->nk == 2) {
        const float4* src = nullptr;
        PF_CHECK(pf_extract_enqueue_pre(ex, xyzi, n, 0, &src));
        PF_CHECK(process_front_graph(od, ex, src));
    } else {
        PF_CHECK(pf_extract_enqueue_single(ex, xyzi, n, 0, 0));
        PF_CHECK(process_extracted(od, ex, nullptr, false));
    }
    if (frame_id) *frame_id = od->frame - 1;
    return PF_OK;
}

extern "C" int pf_frame_wait(pf_odom* h, long long frame_id, double pose_out[7]) {
    PF_REQUIRE(h && pose_out, "null argument");
    PF_CUDA(cudaSetDevice(h->device));
    const long long lo = h->ring_head > pf_odom::kRing ? h->ring_head - pf_odom::kRing : 0;
    for (long long i = h->ring_head - 1; i >= lo; --i) {
        const int slot = (int)(i % pf_odom::kRing);
        if (h->ring_frame[slot] != frame_id) continue;
        PF_CUDA(cudaEventSynchronize(h->ring_ev[slot]));
        const OdomShared& sh = h->h_ring[slot];
        if (frame_id > h->waited_upto) h->waited_upto = frame_id;
        PF_CHECK(decode_frame_error(h, sh.err));
        memcpy(pose_out, sh.pose, sizeof(double) * 7);
        return PF_OK;
    }
    set_error("frame %lld is not among the last %d submitted frames", frame_id, pf_odom::kRing);
    return PF_ERR_STATE;
}

// bytes the library reads back from the device per frame (the pose block: pose, map sizes, error bits)
extern "C" int pf_odom_result_bytes(void) { return (int)sizeof(OdomShared); }

extern "C" int pf_odom_sync(pf_odom* h) {
    PF_REQUIRE(h, "null handle");
    PF_CUDA(cudaSetDevice(h->device));
    return finish_frame(h, nullptr);
}

extern "C" int pf_odom_get_pose_history(pf_odom* h, long long first_frame, int count, double* poses) {
    PF_REQUIRE(h && poses && count >= 0 && first_frame >= 1, "bad argument");
    PF_REQUIRE(first_frame + count <= h->frame && h->frame - first_frame <= kPoseHist, "frames [%lld, %lld) are not in the history (have < %lld, depth %d)",
               first_frame, first_frame + count, h->frame, kPoseHist);
    PF_CUDA(cudaSetDevice(h->device));
    // the requested slots are contiguous modulo the ring: one copy, or two when the range wraps
    const int s0 = (int)(first_frame % kPoseHist), c0 = count < kPoseHist - s0 ? count : kPoseHist - s0;
    if (c0 > 0) PF_CUDA(cudaMemcpyAsync(poses, h->d_pose_hist + 7 * s0, sizeof(double) * 7 * c0, cudaMemcpyDeviceToHost, h->stream));
    if (count > c0) PF_CUDA(cudaMemcpyAsync(poses + 7 * c0, h->d_pose_hist, sizeof(double) * 7 * (count - c0), cudaMemcpyDeviceToHost, h->stream));
    PF_CUDA(cudaStreamSynchronize(h->stream));
    return PF_OK;
}

extern "C" int pf_odom_get_pose(pf_odom* h, double pose[7]) {
    PF_REQUIRE(h && pose, "null argument");
    PF_CUDA(cudaSetDevice(h->device));
    return finish_frame(h, pose);
}

extern "C" int pf_odom_map_size(pf_odom* h, int which, int* n) {
    PF_REQUIRE(h && n && which >= 0 && which < h->nk, "bad argument");
    PF_CUDA(cudaSetDevice(h->device));
    PF_CUDA(cudaMemcpyAsync(h->h_counts, h->d_nmap[h->cur], sizeof(int) * kKinds, cudaMemcpyDeviceToHost, h->stream));
    PF_CUDA(cudaStreamSynchronize(h->stream));
    *n = h->h_counts[which];
    return PF_OK;
}

extern "C" int pf_odom_get_map_part(pf_odom* h, int which, pf_point* out, int cap, int* n) {
    PF_REQUIRE(h && out && n && which >= 0 && which < h->nk, "bad argument");
    int m = 0;
    PF_CHECK(pf_odom_map_size(h, which, &m));
    PF_REQUIRE(m <= cap, "map has %d points, buffer holds %d", m, cap);
    if (m) PF_CUDA(cudaMemcpy(out, h->d_map[h->cur][which], sizeof(Pt) * m, cudaMemcpyDeviceToHost));
    *n = m;
    return PF_OK;
}

// getMap: ES (:210-215) surf map, then corner map; BPF (:683-689) beam, pillar, facade
extern "C" int pf_odom_get_map(pf_odom* h, pf_point* out, int cap, int* n) {
    PF_REQUIRE(h && out && n, "bad argument");
    const int order_es[2] = {1, 0}, order_bpf[3] = {0, 1, 2};
    const int* order = h->nk == 2 ? order_es : order_bpf;
    int total = 0, sz[3] = {0, 0, 0};
    for (int i = 0; i < h->nk; ++i) { PF_CHECK(pf_odom_map_size(h, order[i], &sz[i])); total += sz[i]; }
    PF_REQUIRE(total <= cap, "map has %d points, buffer holds %d", total, cap);
    int off = 0;
    for (int i = 0; i < h->nk; ++i) {
        if (sz[i]) PF_CUDA(cudaMemcpy(out + off, h->d_map[h->cur][order[i]], sizeof(Pt) * sz[i], cudaMemcpyDeviceToHost));
        off += sz[i];
    }
    *n = total;
    return PF_OK;
}

extern "C" int pf_odom_get_iter_poses(pf_odom* h, double* poses, int cap, int* n) {
    PF_REQUIRE(h && poses && n, "bad argument");
    PF_CUDA(cudaSetDevice(h->device));
    PF_CUDA(cudaMemcpyAsync(h->h_iter, h->d_iter_poses, sizeof(double) * 16 * 7, cudaMemcpyDeviceToHost, h->stream));
    PF_CUDA(cudaStreamSynchronize(h->stream));
    int k = h->last_passes < 16 ? h->last_passes : 16;
    if (k > cap) k = cap;
    memcpy(poses, h->h_iter, sizeof(double) * 7 * k);
    *n = k;
    return PF_OK;
}

// ES: the two kinds; BPF: n_edge_ds / map_edge describe the beam cloud, n_surf_ds / map_surf the facade cloud, residual counts are
// line-type (beam + pillar) and plane-type totals
extern "C" int pf_odom_get_stats(pf_odom* h, pf_odom_stats* s) {
    PF_REQUIRE(h && s, "bad argument");
    PF_CUDA(cudaSetDevice(h->device));
    PF_CUDA(cudaMemcpyAsync(h->h_counts, h->d_nds[h->cur ^ 1], sizeof(int) * kKinds, cudaMemcpyDeviceToHost, h->stream));
    PF_CUDA(cudaMemcpyAsync(h->h_counts + kKinds, h->d_nmap[h->cur], sizeof(int) * kKinds, cudaMemcpyDeviceToHost, h->stream));
    PF_CUDA(cudaMemcpyAsync(h->h_state, h->d_state, sizeof(LmState), cudaMemcpyDeviceToHost, h->stream));
    PF_CUDA(cudaStreamSynchronize(h->stream));
    const int plane = h->nk == 2 ? 1 : 2;
    s->n_edge_ds = h->h_counts[0]; s->n_surf_ds = h->h_counts[plane];
    s->map_edge = h->h_counts[kKinds + 0]; s->map_surf = h->h_counts[kKinds + plane];
    s->n_edge_res = h->h_state->n_edge_res; s->n_surf_res = h->h_state->n_surf_res;
    s->passes = h->last_passes;
    s->lm_iterations = h->h_state->iter;
    return PF_OK;
}

// Phase timing of the last update (needs PF_ODOM_TIMING=1 at create time): ms[0] predict + down-sample, ms[1] grid build,
// ms[2] passes up to the last association, ms[3] last pass's 5 LM evaluations, ms[4] append + map maintenance.
extern "C" int pf_odom_get_phase_ms(pf_odom* h, float ms[5]) {
    PF_REQUIRE(h && ms, "null argument");
    PF_REQUIRE(h->timing, "phase timing is off (set PF_ODOM_TIMING=1 before pf_odom_create)");
    PF_CUDA(cudaStreamSynchronize(h->stream));
    for (int i = 0; i < 5; ++i) PF_CUDA(cudaEventElapsedTime(&ms[i], h->tev[i], h->tev[i + 1]));
    return PF_OK;
}

extern "C" void* pf_odom_stream(pf_odom* h) { return h ? (void*)h->stream : nullptr; }

// how many times the steady-state update has been captured as a CUDA graph so far (0: graphs off or not yet in steady state)
extern "C" int pf_odom_graph_captures(pf_odom* h, int* n) {
    PF_REQUIRE(h && n, "null argument");
    *n = h->graph_captures;
    return PF_OK;
}

extern "C" int pf_odom_kernel_launches(pf_odom* h, uint64_t* launches) {
    PF_REQUIRE(h && launches, "null argument");
    *launches = h->ws.launches;
    return PF_OK;
}
