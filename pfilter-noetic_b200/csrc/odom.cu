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
    int n_app[2];          // map sizes after appending the new points
    int err;               // bit 0: map capacity exceeded; bits 1..2: map merge (voxel coordinate range, exception capacity)
    int pad;
    int n_map[2];          // map sizes after the last update (read back with the pose: tight launch bounds for the next frame)
    long long frame;       // frame index this block describes
};

__global__ void k_odom_reset(OdomShared* sh, LmState* S) {
    if (threadIdx.x != 0) return;
    for (int i = 0; i < 9; ++i) { sh->odom.R[i] = (i % 4 == 0) ? 1.0 : 0.0; sh->last_odom.R[i] = sh->odom.R[i]; }
    for (int i = 0; i < 3; ++i) { sh->odom.t[i] = 0; sh->last_odom.t[i] = 0; }
    const double id[7] = {0, 0, 0, 1, 0, 0, 0};
    for (int i = 0; i < 7; ++i) { sh->pose[i] = id[i]; S->x[i] = id[i]; }
    sh->n_app[0] = sh->n_app[1] = 0;
    sh->err = 0;
}

// odom_prediction = odom * (last_odom.inverse() * odom); last_odom = odom; odom = prediction;
// q_w_curr = Quaterniond(odom.rotation()); t_w_curr = odom.translation()                     (:235-240)
__global__ void k_predict(OdomShared* sh, LmState* S) {
    if (threadIdx.x != 0) return;
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
    OdomShared* sh; const LmState* S;
    int map_cap;      // capacity of the map buffers (points)
    double* pose_hist; int hist_slot;
};

// odom <- (q_w_curr, t_w_curr) (:278-280) and addPointsToMap's append loop (:592-604)
__global__ void __launch_bounds__(256) k_append(AppendParams A) {
    const int kind = blockIdx.y;
    __shared__ double s_pose[7];
    if (threadIdx.x < 7) s_pose[threadIdx.x] = A.S->x[threadIdx.x];
    __syncthreads();
    const int n = *A.n_ds[kind], m = *A.n_map[kind];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        int tot = m + n;
        if (tot > A.map_cap) { atomicOr(&A.sh->err, 1); tot = A.map_cap; }
        A.sh->n_app[kind] = tot;
        if (kind == 0) {
            quat_to_mat(s_pose, A.sh->odom.R);
            for (int i = 0; i < 3; ++i) A.sh->odom.t[i] = s_pose[4 + i];
            for (int i = 0; i < 7; ++i) { A.sh->pose[i] = s_pose[i]; A.pose_hist[7 * A.hist_slot + i] = s_pose[i]; }
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

struct InitParams { const float4* feat[2]; const int* n_feat[2]; Pt* map[2]; int* n_map[2]; int* n_sorted[2]; int map_cap; OdomShared* sh; };

// initMapWithPoints (:217-222): the raw first-frame clouds become the maps
__global__ void __launch_bounds__(256) k_init_map(InitParams I) {
    const int kind = blockIdx.y;
    int n = *I.n_feat[kind];
    if (n > I.map_cap) { n = I.map_cap; if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(&I.sh->err, 1); }
    if (blockIdx.x == 0 && threadIdx.x == 0) { *I.n_map[kind] = n; *I.n_sorted[kind] = 0; I.sh->n_map[kind] = n; I.sh->frame = 0; }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 f = I.feat[kind][i];
        Pt o;
        o.x = f.x; o.y = f.y; o.z = f.z; o.rgba = pack_rgba(0, 0, 0, 255);
        I.map[kind][i] = o;
    }
}

__global__ void k_check_map_cap(const int* n_map0, const int* n_map1, int cap, OdomShared* sh, long long frame, const unsigned* merge_err) {
    if (threadIdx.x != 0) return;
    if (*n_map0 > cap || *n_map1 > cap) atomicOr(&sh->err, 1);
    if (*merge_err) atomicOr(&sh->err, (int)(*merge_err & 14u));
    sh->n_map[0] = *n_map0;
    sh->n_map[1] = *n_map1;
    sh->frame = frame;
}

}  // namespace pf

using namespace pf;

// accessors implemented in extract.cu
void pf_extract_device_outputs(pf_extract* h, const float4** edge, const int** n_edge, const float4** surf, const int** n_surf,
                               cudaStream_t* stream, int* edge_cap, int* surf_cap);
int pf_extract_enqueue_single(pf_extract* h, const float* xyzi, int n, int device_input, int want_label);

struct pf_odom {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev = nullptr, ev_done = nullptr;
    pf_odom_params prm{};
    int fcap = 0, mcap = 0, bufcap = 0;
    Workspace ws;
    float4* d_feat[2] = {nullptr, nullptr};
    int* d_nfeat = nullptr;              // [2]
    Pt* d_ds[2] = {nullptr, nullptr};
    int* d_nds = nullptr;                // [2]
    Pt* d_map[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // [buffer][kind]
    int* d_nmap[2] = {nullptr, nullptr}; // [buffer] -> 4 ints: n_map[kind], n_sorted[kind] (merge.cuh: sorted prefix of the map)
    MapMergeScratch msc{};
    bool sorted_known = false;           // false until the first update after init: the raw first-frame maps are unsorted
    int cur = 0;
    float4* d_gpts[2] = {nullptr, nullptr};
    int *d_cs[2] = {nullptr, nullptr}, *d_ce[2] = {nullptr, nullptr}, *d_geom = nullptr;
    int *d_head[2] = {nullptr, nullptr}, *d_hits[2] = {nullptr, nullptr}, *d_next[2] = {nullptr, nullptr}, *d_nn[2] = {nullptr, nullptr};
    uint8_t* d_flag[2] = {nullptr, nullptr};
    double* d_g8[2] = {nullptr, nullptr};
    float* d_wobs[2] = {nullptr, nullptr};           // residual weights (weightType 1 / 2 / 12)
    double* d_wspa[2] = {nullptr, nullptr};
    unsigned long long* d_wminmax = nullptr;         // [2][4]
    LmState* d_state = nullptr;
    double* d_iter_poses = nullptr;
    OdomShared* d_sh = nullptr;
    double* d_pose_hist = nullptr;       // [kPoseHist][7] pose of every update (ring buffer)
    long long frame = 0;                 // frames processed (frame 0 = init)
    // pinned host mirrors
    OdomShared* h_sh = nullptr;
    LmState* h_state = nullptr;
    int* h_counts = nullptr;             // [8]
    double* h_iter = nullptr;            // [16*7]
    bool inited = false;
    int optimization_count = 2;          // :198
    int last_passes = 0;
    // host-side upper bounds of the device-resident counts (launch geometry only; the kernels read the exact counts)
    int map_ub[2] = {0, 0};
    static constexpr int kRing = 32;
    OdomShared* h_ring = nullptr;        // pinned [kRing]: asynchronous read-back of the shared block after every frame
    cudaEvent_t ring_ev[kRing] = {};
    long long ring_frame[kRing] = {};
    int ring_add[kRing][2] = {};
    long long ring_head = 0;             // frames recorded in the ring
    long long known_frame = -1;          // newest frame whose exact map sizes have been folded into map_ub
    // optional phase timing (PF_ODOM_TIMING=1): CUDA events at the phase boundaries of the last update
    bool timing = false;
    cudaEvent_t tev[8] = {};
    float phase_ms[8] = {};
};

namespace {

int odom_alloc(pf_odom* h) {
    const int fcap = h->fcap, bufcap = h->bufcap;
    PF_CHECK(workspace_create(h->ws, 2 * bufcap, h->stream));
    PF_CUDA(cudaMalloc(&h->d_nfeat, sizeof(int) * 2));
    PF_CUDA(cudaMalloc(&h->d_nds, sizeof(int) * 2));
    PF_CUDA(cudaMalloc(&h->d_geom, sizeof(int) * 12));
    for (int k = 0; k < 2; ++k) {
        PF_CUDA(cudaMalloc(&h->d_feat[k], sizeof(float4) * fcap));
        PF_CUDA(cudaMalloc(&h->d_ds[k], sizeof(Pt) * fcap));
        for (int b = 0; b < 2; ++b) PF_CUDA(cudaMalloc(&h->d_map[b][k], sizeof(Pt) * bufcap));
        PF_CUDA(cudaMalloc(&h->d_nmap[k], sizeof(int) * 4));
        PF_CUDA(cudaMemset(h->d_nmap[k], 0, sizeof(int) * 4));
        PF_CUDA(cudaMalloc(&h->d_gpts[k], sizeof(float4) * bufcap));
        PF_CUDA(cudaMalloc(&h->d_cs[k], sizeof(int) * (size_t)kGridCellCap));
        PF_CUDA(cudaMalloc(&h->d_ce[k], sizeof(int) * (size_t)kGridCellCap));
        PF_CUDA(cudaMalloc(&h->d_head[k], sizeof(int) * bufcap));
        PF_CUDA(cudaMalloc(&h->d_hits[k], sizeof(int) * bufcap));
        PF_CUDA(cudaMemset(h->d_head[k], 0xff, sizeof(int) * bufcap));
        PF_CUDA(cudaMemset(h->d_hits[k], 0, sizeof(int) * bufcap));
        PF_CUDA(cudaMalloc(&h->d_next[k], sizeof(int) * 5 * fcap));
        PF_CUDA(cudaMalloc(&h->d_nn[k], sizeof(int) * 5 * fcap));
        PF_CUDA(cudaMalloc(&h->d_flag[k], fcap));
        PF_CUDA(cudaMemset(h->d_flag[k], 0, fcap));
        PF_CUDA(cudaMalloc(&h->d_g8[k], sizeof(double) * 8 * fcap));
        PF_CUDA(cudaMalloc(&h->d_wobs[k], sizeof(float) * fcap));
        PF_CUDA(cudaMalloc(&h->d_wspa[k], sizeof(double) * fcap));
    }
    PF_CUDA(cudaMalloc(&h->d_wminmax, sizeof(unsigned long long) * 8));
    PF_CUDA(cudaMemset(h->d_wminmax, 0, sizeof(unsigned long long) * 8));
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
    PF_CUDA(cudaMallocHost(&h->h_counts, sizeof(int) * 8));
    PF_CUDA(cudaMallocHost(&h->h_iter, sizeof(double) * 16 * 7));
    PF_CUDA(cudaMallocHost(&h->h_ring, sizeof(OdomShared) * pf_odom::kRing));
    h->timing = getenv("PF_ODOM_TIMING") != nullptr;
    if (h->timing) for (int i = 0; i < 8; ++i) PF_CUDA(cudaEventCreate(&h->tev[i]));
    for (int i = 0; i < pf_odom::kRing; ++i) PF_CUDA(cudaEventCreateWithFlags(&h->ring_ev[i], cudaEventDisableTiming));
    k_odom_reset<<<1, 32, 0, h->stream>>>(h->d_sh, h->d_state);
    PF_CUDA(cudaStreamSynchronize(h->stream));
    return PF_OK;
}

int upload_features(pf_odom* h, const float* edge, int ne, const float* surf, int ns) {
    PF_REQUIRE(ne >= 0 && ns >= 0 && (edge || ne == 0) && (surf || ns == 0), "bad feature arrays");
    PF_REQUIRE(ne <= h->fcap && ns <= h->fcap, "feature cloud (%d / %d points) exceeds max_features %d", ne, ns, h->fcap);
    h->h_counts[0] = ne; h->h_counts[1] = ns;
    PF_CUDA(cudaMemcpyAsync(h->d_nfeat, h->h_counts, sizeof(int) * 2, cudaMemcpyHostToDevice, h->stream));
    if (ne) PF_CUDA(cudaMemcpyAsync(h->d_feat[0], edge, sizeof(float4) * ne, cudaMemcpyHostToDevice, h->stream));
    if (ns) PF_CUDA(cudaMemcpyAsync(h->d_feat[1], surf, sizeof(float4) * ns, cudaMemcpyHostToDevice, h->stream));
    return PF_OK;
}

// asynchronous read-back of the shared block: lets later frames size their launches from exact map counts
int ring_record(pf_odom* h, int add_e, int add_s) {
    const int slot = (int)(h->ring_head % pf_odom::kRing);
    PF_CUDA(cudaMemcpyAsync(h->h_ring + slot, h->d_sh, sizeof(OdomShared), cudaMemcpyDeviceToHost, h->stream));
    PF_CUDA(cudaEventRecord(h->ring_ev[slot], h->stream));
    h->ring_frame[slot] = h->frame;
    h->ring_add[slot][0] = add_e; h->ring_add[slot][1] = add_s;
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
        int ub[2] = {h->h_ring[slot].n_map[0], h->h_ring[slot].n_map[1]};
        for (long long j = i + 1; j < h->ring_head; ++j) {
            const int sj = (int)(j % pf_odom::kRing);
            ub[0] += h->ring_add[sj][0]; ub[1] += h->ring_add[sj][1];
        }
        for (int k = 0; k < 2; ++k) h->map_ub[k] = ub[k] < h->bufcap ? ub[k] : h->bufcap;
        h->known_frame = h->ring_frame[slot];
        break;
    }
    cudaGetLastError();   // cudaErrorNotReady from the queries is not an error
}

int enqueue_init(pf_odom* h, const float4* const feat[2], const int* const n_feat[2], int ub_e, int ub_s) {
    InitParams I{};
    for (int k = 0; k < 2; ++k) {
        I.feat[k] = feat[k]; I.n_feat[k] = n_feat[k]; I.map[k] = h->d_map[h->cur][k];
        I.n_map[k] = h->d_nmap[h->cur] + k; I.n_sorted[k] = h->d_nmap[h->cur] + 2 + k;
    }
    I.map_cap = h->mcap;
    I.sh = h->d_sh;
    k_init_map<<<dim3(2 * kSMs, 2), 256, 0, h->stream>>>(I);
    h->ws.launches += 1;
    PF_CUDA(cudaGetLastError());
    h->optimization_count = 12;   // :221
    h->inited = true;
    h->sorted_known = false;
    h->map_ub[0] = ub_e < h->mcap ? ub_e : h->mcap;
    h->map_ub[1] = ub_s < h->mcap ? ub_s : h->mcap;
    h->frame = 0;
    PF_CHECK(ring_record(h, 0, 0));
    h->frame = 1;
    return PF_OK;
}

int enqueue_update(pf_odom* h, const float4* const feat[2], const int* const n_feat[2], int ub_e, int ub_s) {
    if (h->optimization_count > 2) h->optimization_count--;   // :232-233
    ring_refresh(h);
    if (ub_e > h->fcap) ub_e = h->fcap;
    if (ub_s > h->fcap) ub_s = h->fcap;
    if (ub_e < 1) ub_e = 1;
    if (ub_s < 1) ub_s = 1;
    const int mub_e = h->map_ub[0] > 1 ? h->map_ub[0] : 1, mub_s = h->map_ub[1] > 1 ? h->map_ub[1] : 1;
    const int passes = h->optimization_count;
    h->last_passes = passes;
    Workspace& ws = h->ws;
    const int cur = h->cur, nxt = cur ^ 1;
    auto mark = [&](int i) { if (h->timing) cudaEventRecord(h->tev[i], h->stream); };
    mark(0);
    PF_CHECK(workspace_begin_step(ws));
    k_predict<<<1, 32, 0, h->stream>>>(h->d_sh, h->d_state);
    ws.launches += 1;
    // VoxelGrid down-sampling, leaf sizes as set by init (:189-190): setLeafSize takes floats
    VoxParams V{};
    V.mode = VOX_PCL;
    const float leaf[2] = {(float)h->prm.map_resolution, (float)(h->prm.map_resolution * 2)};
    for (int k = 0; k < 2; ++k)
        V.c[k] = VoxCloud{reinterpret_cast<const Pt*>(feat[k]), n_feat[k], h->d_ds[k], h->d_nds + k, leaf[k], 1};
    PF_CHECK(voxelize(ws, V, 0, ub_e, ub_s));
    mark(1);
    // search grids over the current maps
    GridBuild G{};
    for (int k = 0; k < 2; ++k) {
        G.map[k] = h->d_map[cur][k]; G.n_map[k] = h->d_nmap[cur] + k; G.pts[k] = h->d_gpts[k];
        G.cell_start[k] = h->d_cs[k]; G.cell_end[k] = h->d_ce[k]; G.geom[k] = h->d_geom + 6 * k;
    }
    PF_CHECK(build_grids(ws, G, 2, mub_e, mub_s));
    mark(2);
    // optimisation passes
    AssocParams A{};
    LmParams L{};
    for (int k = 0; k < 2; ++k) {
        A.c[k] = AssocCloud{h->d_ds[k], h->d_nds + k, h->d_map[cur][k], h->d_nmap[cur] + k,
                            KnnGrid{h->d_gpts[k], h->d_cs[k], h->d_ce[k], h->d_geom + 6 * k},
                            h->d_head[k], h->d_hits[k], h->d_next[k], h->d_nn[k], h->d_flag[k], h->d_g8[k], h->d_wobs[k], h->d_wspa[k]};
        L.src[k] = ResidualSrc{h->d_ds[k], nullptr, h->d_flag[k], h->d_g8[k], h->d_nds + k, h->d_wobs[k], h->d_wspa[k]};
    }
    A.pose = h->d_state->x;
    A.k_new = h->prm.k_new; A.theta_p = h->prm.theta_p; A.theta_max = h->prm.theta_max;
    A.min_edge_map = 10; A.min_surf_map = 50;   // :247
    A.weight_type = (int)h->prm.weight_type; A.w_minmax = h->d_wminmax;
    L.weight_type = (int)h->prm.weight_type; L.w_minmax = h->d_wminmax;
    L.state = h->d_state; L.iter_poses = h->d_iter_poses; L.eval_only = 0;
    for (int it = 0; it < passes; ++it) {
        PF_CHECK(associate_pass(h->stream, A, ub_e, ub_s, &ws.launches));
        if (it == passes - 1) mark(3);
        PF_CHECK(lm_solve(h->stream, L, nullptr, it == 0, &ws.launches, ub_e, ub_s));
    }
    mark(4);
    // append + map maintenance
    AppendParams P{};
    for (int k = 0; k < 2; ++k) { P.ds[k] = h->d_ds[k]; P.n_ds[k] = h->d_nds + k; P.map[k] = h->d_map[cur][k]; P.n_map[k] = h->d_nmap[cur] + k; }
    P.sh = h->d_sh; P.S = h->d_state; P.map_cap = h->bufcap;
    P.pose_hist = h->d_pose_hist; P.hist_slot = (int)(h->frame % kPoseHist);
    k_append<<<dim3(kSMs, 2), 256, 0, h->stream>>>(P);
    ws.launches += 1;
    // rgbds(tmpSurf, map_resolution * 2) / rgbds(tmpCorner, map_resolution) with the float member map_resolution (:625-626)
    const float mres = (float)h->prm.map_resolution;
    const float mleaf[2] = {mres, mres * 2};
    MapMergeParams M{};
    for (int k = 0; k < 2; ++k)
        M.c[k] = MapMergeCloud{h->d_map[cur][k], h->d_nmap[cur] + 2 + k, h->d_sh->n_app + k, h->d_map[nxt][k], h->d_nmap[nxt] + k,
                               h->d_nmap[nxt] + 2 + k, mleaf[k]};
    M.center = h->d_sh->odom.t;
    M.k_new = h->prm.k_new; M.theta_p = h->prm.theta_p; M.theta_max = h->prm.theta_max;
    M.s = h->msc;
    const int app_e = mub_e + ub_e < h->bufcap ? mub_e + ub_e : h->bufcap, app_s = mub_s + ub_s < h->bufcap ? mub_s + ub_s : h->bufcap;
    // unsorted part: everything on the first update (raw first-frame maps), later last update's exceptions + this frame's points
    const int capb_e = h->sorted_known ? (kMergeExcCap + ub_e < app_e ? kMergeExcCap + ub_e : app_e) : app_e;
    const int capb_s = h->sorted_known ? (kMergeExcCap + ub_s < app_s ? kMergeExcCap + ub_s : app_s) : app_s;
    PF_CHECK(map_merge(ws, M, capb_e, capb_s, mub_e, mub_s));
    h->sorted_known = true;
    k_check_map_cap<<<1, 32, 0, h->stream>>>(h->d_nmap[nxt], h->d_nmap[nxt] + 1, h->mcap, h->d_sh, h->frame, map_merge_error_word(ws));
    ws.launches += 1;
    PF_CUDA(cudaGetLastError());
    h->cur = nxt;
    mark(5);
    h->map_ub[0] = app_e; h->map_ub[1] = app_s;     // the update never grows a map beyond old + appended
    PF_CHECK(ring_record(h, ub_e, ub_s));
    h->frame += 1;
    return PF_OK;
}

int finish_frame(pf_odom* h, double pose_out[7]) {
    PF_CUDA(cudaMemcpyAsync(h->h_sh, h->d_sh, sizeof(OdomShared), cudaMemcpyDeviceToHost, h->stream));
    PF_CUDA(cudaStreamSynchronize(h->stream));
    if (h->h_sh->err & 1) {
        set_error("local map exceeded max_map_points = %d", h->mcap);
        return PF_ERR_CAPACITY;
    }
    if (h->h_sh->err & 14) {
        set_error("map update failed (bits %d: 2 = map_resolution below 0.2 m, 4 = more than %d centroids left their voxel)", h->h_sh->err & 6, kMergeExcCap);
        return PF_ERR_CAPACITY;
    }
    if (pose_out) memcpy(pose_out, h->h_sh->pose, sizeof(double) * 7);
    if (h->inited) {
        h->map_ub[0] = h->h_sh->n_map[0]; h->map_ub[1] = h->h_sh->n_map[1];
        h->known_frame = h->frame - 1;
    }
    return PF_OK;
}

}  // namespace

extern "C" int pf_odom_create(const pf_odom_params* p, int device, pf_odom** out) {
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
    h->device = device;
    h->prm = *p;
    h->fcap = p->max_features > 0 ? p->max_features : 131072;
    h->mcap = p->max_map_points > 0 ? p->max_map_points : (2 << 20);
    if (h->mcap < h->fcap) h->mcap = h->fcap;   // the raw first-frame clouds become the maps (:217-222)
    h->bufcap = h->mcap + h->fcap;
    PF_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    PF_CUDA(cudaEventCreateWithFlags(&h->ev, cudaEventDisableTiming));
    PF_CUDA(cudaEventCreateWithFlags(&h->ev_done, cudaEventDisableTiming));
    int rc = odom_alloc(h);
    if (rc != PF_OK) { pf_odom_destroy(h); return rc; }
    *out = h;
    return PF_OK;
}

extern "C" int pf_odom_destroy(pf_odom* h) {
    if (!h) return PF_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    workspace_destroy(h->ws);
    map_merge_scratch_destroy(h->msc);
    cudaFree(h->d_nfeat); cudaFree(h->d_nds); cudaFree(h->d_geom);
    for (int k = 0; k < 2; ++k) {
        cudaFree(h->d_feat[k]); cudaFree(h->d_ds[k]); cudaFree(h->d_map[0][k]); cudaFree(h->d_map[1][k]); cudaFree(h->d_nmap[k]);
        cudaFree(h->d_gpts[k]); cudaFree(h->d_cs[k]); cudaFree(h->d_ce[k]); cudaFree(h->d_head[k]); cudaFree(h->d_hits[k]);
        cudaFree(h->d_next[k]); cudaFree(h->d_nn[k]); cudaFree(h->d_flag[k]); cudaFree(h->d_g8[k]);
        cudaFree(h->d_wobs[k]); cudaFree(h->d_wspa[k]);
    }
    cudaFree(h->d_wminmax);
    cudaFree(h->d_state); cudaFree(h->d_iter_poses); cudaFree(h->d_sh); cudaFree(h->d_pose_hist);
    cudaFreeHost(h->h_sh); cudaFreeHost(h->h_state); cudaFreeHost(h->h_counts); cudaFreeHost(h->h_iter); cudaFreeHost(h->h_ring);
    for (int i = 0; i < pf_odom::kRing; ++i) if (h->ring_ev[i]) cudaEventDestroy(h->ring_ev[i]);
    if (h->ev) cudaEventDestroy(h->ev);
    if (h->ev_done) cudaEventDestroy(h->ev_done);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return PF_OK;
}

extern "C" int pf_odom_init_map(pf_odom* h, const float* edge, int n_edge, const float* surf, int n_surf) {
    PF_REQUIRE(h, "null handle");
    PF_CUDA(cudaSetDevice(h->device));
    PF_CHECK(upload_features(h, edge, n_edge, surf, n_surf));
    const float4* feat[2] = {h->d_feat[0], h->d_feat[1]};
    const int* nf[2] = {h->d_nfeat, h->d_nfeat + 1};
    PF_CHECK(enqueue_init(h, feat, nf, n_edge, n_surf));
    return finish_frame(h, nullptr);
}

extern "C" int pf_odom_update(pf_odom* h, const float* edge, int n_edge, const float* surf, int n_surf, double pose_out[7]) {
    PF_REQUIRE(h, "null handle");
    if (!h->inited) { set_error("pf_odom_update before pf_odom_init_map"); return PF_ERR_STATE; }
    PF_CUDA(cudaSetDevice(h->device));
    PF_CHECK(upload_features(h, edge, n_edge, surf, n_surf));
    const float4* feat[2] = {h->d_feat[0], h->d_feat[1]};
    const int* nf[2] = {h->d_nfeat, h->d_nfeat + 1};
    PF_CHECK(enqueue_update(h, feat, nf, n_edge, n_surf));
    return finish_frame(h, pose_out);
}

static int process_extracted(pf_odom* h, pf_extract* ex, double pose_out[7], bool sync) {
    PF_REQUIRE(h && ex, "null handle");
    PF_CUDA(cudaSetDevice(h->device));
    const float4* feat[2];
    const int* nf[2];
    cudaStream_t exs;
    int ecap, scap;
    pf_extract_device_outputs(ex, &feat[0], &nf[0], &feat[1], &nf[1], &exs, &ecap, &scap);
    PF_REQUIRE(scap <= h->fcap, "scan of %d points exceeds max_features %d", scap, h->fcap);
    PF_CUDA(cudaEventRecord(h->ev, exs));
    PF_CUDA(cudaStreamWaitEvent(h->stream, h->ev, 0));
    if (!h->inited) {
        PF_CHECK(enqueue_init(h, feat, nf, ecap, scap));
    } else {
        PF_CHECK(enqueue_update(h, feat, nf, ecap, scap));
    }
    // the extractor's output buffers are reused by the next frame: it must not start before this frame consumed them
    PF_CUDA(cudaEventRecord(h->ev_done, h->stream));
    PF_CUDA(cudaStreamWaitEvent(exs, h->ev_done, 0));
    return sync ? finish_frame(h, pose_out) : PF_OK;
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
    PF_CUDA(cudaStreamSynchronize(h->stream));
    for (int i = 0; i < count; ++i)
        PF_CUDA(cudaMemcpy(poses + 7 * i, h->d_pose_hist + 7 * ((first_frame + i) % kPoseHist), sizeof(double) * 7, cudaMemcpyDeviceToHost));
    return PF_OK;
}

extern "C" int pf_odom_get_pose(pf_odom* h, double pose[7]) {
    PF_REQUIRE(h && pose, "null argument");
    PF_CUDA(cudaSetDevice(h->device));
    return finish_frame(h, pose);
}

extern "C" int pf_odom_map_size(pf_odom* h, int which, int* n) {
    PF_REQUIRE(h && n && (which == 0 || which == 1), "bad argument");
    PF_CUDA(cudaSetDevice(h->device));
    PF_CUDA(cudaMemcpyAsync(h->h_counts, h->d_nmap[h->cur], sizeof(int) * 2, cudaMemcpyDeviceToHost, h->stream));
    PF_CUDA(cudaStreamSynchronize(h->stream));
    *n = h->h_counts[which];
    return PF_OK;
}

extern "C" int pf_odom_get_map_part(pf_odom* h, int which, pf_point* out, int cap, int* n) {
    PF_REQUIRE(h && out && n && (which == 0 || which == 1), "bad argument");
    int m = 0;
    PF_CHECK(pf_odom_map_size(h, which, &m));
    PF_REQUIRE(m <= cap, "map has %d points, buffer holds %d", m, cap);
    if (m) PF_CUDA(cudaMemcpy(out, h->d_map[h->cur][which], sizeof(Pt) * m, cudaMemcpyDeviceToHost));
    *n = m;
    return PF_OK;
}

// getMap (:210-215): surf map, then corner map
extern "C" int pf_odom_get_map(pf_odom* h, pf_point* out, int cap, int* n) {
    PF_REQUIRE(h && out && n, "bad argument");
    int ns = 0, ne = 0;
    PF_CHECK(pf_odom_map_size(h, 1, &ns));
    PF_CHECK(pf_odom_map_size(h, 0, &ne));
    PF_REQUIRE(ns + ne <= cap, "map has %d points, buffer holds %d", ns + ne, cap);
    if (ns) PF_CUDA(cudaMemcpy(out, h->d_map[h->cur][1], sizeof(Pt) * ns, cudaMemcpyDeviceToHost));
    if (ne) PF_CUDA(cudaMemcpy(out + ns, h->d_map[h->cur][0], sizeof(Pt) * ne, cudaMemcpyDeviceToHost));
    *n = ns + ne;
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

extern "C" int pf_odom_get_stats(pf_odom* h, pf_odom_stats* s) {
    PF_REQUIRE(h && s, "bad argument");
    PF_CUDA(cudaSetDevice(h->device));
    PF_CUDA(cudaMemcpyAsync(h->h_counts, h->d_nds, sizeof(int) * 2, cudaMemcpyDeviceToHost, h->stream));
    PF_CUDA(cudaMemcpyAsync(h->h_counts + 2, h->d_nmap[h->cur], sizeof(int) * 2, cudaMemcpyDeviceToHost, h->stream));
    PF_CUDA(cudaMemcpyAsync(h->h_state, h->d_state, sizeof(LmState), cudaMemcpyDeviceToHost, h->stream));
    PF_CUDA(cudaStreamSynchronize(h->stream));
    s->n_edge_ds = h->h_counts[0]; s->n_surf_ds = h->h_counts[1];
    s->map_edge = h->h_counts[2]; s->map_surf = h->h_counts[3];
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

extern "C" int pf_odom_kernel_launches(pf_odom* h, uint64_t* launches) {
    PF_REQUIRE(h && launches, "null argument");
    *launches = h->ws.launches;
    return PF_OK;
}
