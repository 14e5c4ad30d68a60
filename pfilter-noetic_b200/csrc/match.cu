// K4-K6: association pass.  See match.cuh; arithmetic spec in SURVEY.md appendix A.2.
//
// k_assoc_knn     half a warp per query: transform by the current pose (pointAssociateToMap :162-168), exact 5-NN in the
//                 1 m grid; leaves the five map indices (or -1).
// k_assoc_fit     one thread per query: the geometric fit in fp64 registers -- line: centroid + 3x3 covariance + symmetric
//                 eigen-solve, lambda2 > 3 lambda1 (:302-331); plane: 5x3 column-pivoted Householder least squares,
//                 5 x |n.p + d| <= 0.2 (:449-476).  Geometry-valid queries push their 5 hits on per-map-point lists.
// k_assoc_persist one thread per geometry-valid query: the reference updates the neighbours' observe counter g inside
//                 its serial query loop (:345-346), so query i sees the increments of queries 0..i-1 of the same pass.
//                 Geometry does not depend on g, hence the value query i saw is g0 + #(earlier valid queries that hit
//                 the same map point) = g0 + rank of its hit in that point's list (saturating at 255).  Then observe /
//                 round, the skip rule (:348-353) and the query's own counters (:354-355).  The last reader of a list
//                 commits g = min(255, g0 + hits) and clears the list.
#include "match.cuh"
#include "math.cuh"

namespace pf {

// The search and the fit are two kernels.  k_assoc_knn: a half warp per query (two searches side by side per warp: a search is a
// chain of dependent memory round trips, not a throughput problem), 40 registers, every query of the frame in flight at once; it
// leaves the five map indices (or -1: fewer than five neighbours within 1 m).  k_assoc_fit: one THREAD per query runs the fp64 line /
// plane fit (~3000 dependent instructions, 128 registers).  Fused in one kernel (round 1: a warp searched four queries, then four
// of its lanes fitted them) the fit's registers and its 12 us chain were paid by every search warp: 522 CTAs at two per SM held
// the whole GPU for 35 us per pass -- nothing for one sequence, but the ceiling of several sequences sharing the GPU.
__global__ void __launch_bounds__(256) k_assoc_knn(AssocParams P) {
    PF_PDL_ENTRY();
    const int kind = blockIdx.y;
    const AssocCloud& c = P.c[kind];
    const unsigned lane = lane_id();
    const int nq = *c.n_q;
    const int nhalf = (gridDim.x * blockDim.x) >> 4;
    const bool guard = P.guard == nullptr || *P.guard != 0;   // :247
    if (!guard) return;                                        // k_assoc_fit clears the flags
    for (int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 4; q < nq; q += nhalf) {      // uniform within the half warp
        const Pt qp = c.queries[q];
        const D3 pw = pose_apply(P.pose, d3((double)qp.x, (double)qp.y, (double)qp.z));
        int idx[5];
        float d2[5];
        const bool found = knn5_group<16>(c.grid, (float)pw.x, (float)pw.y, (float)pw.z, idx, d2);   // :299-300 / :447-451
        const int j = (int)(lane & 15u);
        if (j < 5) c.nn_idx[5 * q + j] = found ? idx[j] : -1;
    }
}

__global__ void __launch_bounds__(128) k_assoc_fit(AssocParams P) {
    PF_PDL_ENTRY();
    const int kind = blockIdx.y;
    const AssocCloud& c = P.c[kind];
    const int nq = *c.n_q;
    if (P.weight_type != 0 && blockIdx.x == 0 && threadIdx.x < 4)     // min / max slots of this pass (read by k_assoc_persist, which follows)
        P.w_minmax[4 * kind + threadIdx.x] = (threadIdx.x & 1) ? 0ull : ~0ull;
    const bool guard = P.guard == nullptr || *P.guard != 0;   // :247
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += gridDim.x * blockDim.x) {
        int my_idx[5] = {-1, -1, -1, -1, -1};
        if (guard) {
#pragma unroll
            for (int j = 0; j < 5; ++j) my_idx[j] = c.nn_idx[5 * q + j];
        }
        const bool my_found = my_idx[0] >= 0;
        {
            unsigned flag = 0;
            if (my_found) {
                D3 nb[5];
#pragma unroll
                for (int j = 0; j < 5; ++j) {
                    const Pt m = c.map[my_idx[j]];
                    nb[j] = d3((double)m.x, (double)m.y, (double)m.z);
                }
                double g8[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                bool valid;
                if (kind == 0) {
                    D3 ctr = d3(0, 0, 0);
#pragma unroll
                    for (int j = 0; j < 5; ++j) ctr = ctr + nb[j];
                    ctr = d3(ctr.x / 5.0, ctr.y / 5.0, ctr.z / 5.0);
                    double c00 = 0, c01 = 0, c02 = 0, c11 = 0, c12 = 0, c22 = 0;
#pragma unroll
                    for (int j = 0; j < 5; ++j) {
                        const D3 d = nb[j] - ctr;
                        c00 += d.x * d.x; c01 += d.x * d.y; c02 += d.x * d.z;
                        c11 += d.y * d.y; c12 += d.y * d.z; c22 += d.z * d.z;
                    }
                    double w[3];
                    D3 v;
                    eig3_sym(c00, c01, c02, c11, c12, c22, w, v);
                    valid = w[2] > 3 * w[1];                                   // :326
                    const D3 a = (0.1 * v) + ctr, b = (-0.1 * v) + ctr;        // :330-331
                    g8[0] = a.x; g8[1] = a.y; g8[2] = a.z; g8[3] = b.x; g8[4] = b.y; g8[5] = b.z;
                } else {
                    double A[3][5], rhs[5], n[3];
#pragma unroll
                    for (int j = 0; j < 5; ++j) { A[0][j] = nb[j].x; A[1][j] = nb[j].y; A[2][j] = nb[j].z; rhs[j] = -1.0; }
                    plane_lsq_5x3(A, rhs, n);                                  // :461
                    const double nn = sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
                    const double negOA = 1 / nn;                               // :462
                    n[0] = n[0] / nn; n[1] = n[1] / nn; n[2] = n[2] / nn;      // :463
                    valid = true;
#pragma unroll
                    for (int j = 0; j < 5; ++j)
                        if (fabs(n[0] * nb[j].x + n[1] * nb[j].y + n[2] * nb[j].z + negOA) > 0.2) valid = false;   // :466-476 (NaN -> stays valid, as in the reference)
                    g8[0] = n[0]; g8[1] = n[1]; g8[2] = n[2];
                    g8[3] = (double)(float)negOA;   // surfInfo::negative_OA_dot_norm is a float (include/odomEstimationClass.h:93-94)
                }
                if (valid) {
                    flag = 1;
#pragma unroll
                    for (int j = 0; j < 5; ++j) {
                        const int m = my_idx[j], hit = 5 * q + j;
                        c.nn_idx[hit] = m;
                        c.next[hit] = atomicExch(&c.head[m], hit);
                        atomicAdd(&c.hits[m], 1);
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) c.geom[8 * (size_t)q + j] = g8[j];
                }
            }
            c.flag[q] = (uint8_t)flag;
        }
    }
}

__global__ void __launch_bounds__(128) k_assoc_persist(AssocParams P) {
    PF_PDL_ENTRY();
    const int kind = blockIdx.y;
    const AssocCloud& c = P.c[kind];
    const int nq = *c.n_q;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += gridDim.x * blockDim.x) {
    if (c.flag[q] != 1) continue;
    int m[5], g0[5], len[5], rank[5], h[5];
    int sg = 0, sr = 0;
#pragma unroll
    for (int j = 0; j < 5; ++j) m[j] = c.nn_idx[5 * q + j];
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        const unsigned rgba = c.map[m[j]].rgba;
        g0[j] = (int)pt_g(rgba);
        sr += (int)pt_r(rgba);
        h[j] = c.head[m[j]];
        rank[j] = 0; len[j] = 0;
    }
    // the five hit lists are walked side by side: one hop of each per round, so the pointer chases overlap (a list is as long as the
    // number of valid queries that hit the map point in this pass; walking them one after the other was most of the kernel's time)
    while (true) {
        bool any = false;
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            if (h[j] >= 0) {
                rank[j] += (h[j] / 5 < q) ? 1 : 0;
                len[j] += 1;
                h[j] = c.next[h[j]];
                any = true;
            }
        }
        if (!any) break;
    }
#pragma unroll
    for (int j = 0; j < 5; ++j) sg += min(255, g0[j] + rank[j]);          // the counter value the serial loop would have read (:332-336)
    float observe = (float)((double)sg / 5.0 + 1);                           // :332-338
    const float round = (float)((double)sr / 5.0);                           // :339-344
    if (__fdiv_rn(observe, round) > 5) observe = 255;                        // :348-349 (round == 0 -> inf)
    const bool skip = (observe < __fmul_rn(round, P.theta_p)) && (round > (float)P.k_new) && (observe < (float)P.theta_max);   // :350
    if (!skip) {
        const unsigned r = (unsigned)min(255, (int)round), g = (unsigned)min(255, (int)observe);   // :354-355
        Pt* qp = c.queries + q;
        qp->rgba = (qp->rgba & 0xffff0000u) | r | (g << 8);
        c.flag[q] = 2;
        if (P.weight_type != 0) {     // weight inputs of the residual block (:360-385): observe, mean distance of the 5 neighbours to their centroid
            D3 nb[5], cn = d3(0, 0, 0);
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                const Pt mp = c.map[m[j]];
                nb[j] = d3((double)mp.x, (double)mp.y, (double)mp.z);
                cn = cn + nb[j];
            }
            cn = d3(cn.x / 5, cn.y / 5, cn.z / 5);
            float sum = 0.f;
#pragma unroll
            for (int j = 0; j < 5; ++j) sum = __fadd_rn(sum, (float)norm3(cn - nb[j]));      // float sum += double norm
            const double spa = (double)(float)((double)sum / 5.0);                            // sum /= 5.0 (float), pushed as double
            const double obs = (double)observe;
            c.w_obs[q] = observe;
            c.w_spa[q] = spa;
            unsigned long long* mm = P.w_minmax + 4 * kind;
            atomicMin(mm + 0, (unsigned long long)__double_as_longlong(obs));
            atomicMax(mm + 1, (unsigned long long)__double_as_longlong(obs));
            atomicMin(mm + 2, (unsigned long long)__double_as_longlong(spa));
            atomicMax(mm + 3, (unsigned long long)__double_as_longlong(spa));
        }
    }
    // release the lists; the last reader of a map point commits its counter and empties the list.  Every value this thread read
    // from the lists has been consumed above (the walk ends on them), so one fence orders all its reads before the five releases,
    // and the five decrements (five different map points: the neighbours of a query are distinct) travel together instead of
    // fence - round trip - fence - round trip ...: that serial tail was more than half of the kernel.
    __threadfence();
    int left[5];
#pragma unroll
    for (int j = 0; j < 5; ++j) left[j] = atomicSub(&c.hits[m[j]], 1);
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        if (left[j] == 1) {
            reinterpret_cast<uint8_t*>(c.map + m[j])[13] = (uint8_t)min(255, g0[j] + len[j]);   // g = min(255, g + 1) per hit (:345-346)
            c.head[m[j]] = -1;
        }
    }
    }
}

int associate_pass(cudaStream_t stream, const AssocParams& P, int qcap0, int qcap1, uint64_t* launches) {
    const int qcap = qcap0 > qcap1 ? qcap0 : qcap1;
    if (qcap <= 0) return PF_OK;
    int gk = div_up(qcap, 16), gp = div_up(qcap, 128);      // a half warp per query / a thread per query
    if (gk > 8 * kSMs) gk = 8 * kSMs;
    if (gp > 4 * kSMs) gp = 4 * kSMs;
    PF_CUDA(launch_pdl(k_assoc_knn, dim3(gk, 2), dim3(256), 0, stream, P));
    PF_CUDA(launch_pdl(k_assoc_fit, dim3(gp, 2), dim3(128), 0, stream, P));
    PF_CUDA(launch_pdl(k_assoc_persist, dim3(gp, 2), dim3(128), 0, stream, P));
    if (launches) *launches += 3;
    PF_CUDA(cudaGetLastError());
    return PF_OK;
}

}  // namespace pf

using namespace pf;

namespace {
struct AssocTap {
    cudaStream_t stream = nullptr;
    Workspace ws;
    Pt *d_map = nullptr, *d_q = nullptr;
    float4* d_pts = nullptr;
    int *d_cs = nullptr, *d_ce = nullptr, *d_geom = nullptr, *d_counts = nullptr, *d_head = nullptr, *d_hits = nullptr, *d_next = nullptr,
        *d_nn = nullptr;
    uint8_t* d_flag = nullptr;
    double *d_g8 = nullptr, *d_pose = nullptr;
    ~AssocTap() {
        workspace_destroy(ws);
        cudaFree(d_map); cudaFree(d_q); cudaFree(d_pts); cudaFree(d_cs); cudaFree(d_ce); cudaFree(d_geom); cudaFree(d_counts);
        cudaFree(d_head); cudaFree(d_hits); cudaFree(d_next); cudaFree(d_nn); cudaFree(d_flag); cudaFree(d_g8); cudaFree(d_pose);
        if (stream) cudaStreamDestroy(stream);
    }
};
}  // namespace

extern "C" int pf_associate(int device, int kind, pf_point* map, int m, pf_point* queries, int q, const double pose[7], int k_new,
                            float theta_p, int theta_max, uint8_t* flag, double* geom8) {
    PF_REQUIRE((kind == 0 || kind == 1) && m >= 0 && q >= 0 && (map || m == 0) && (queries || q == 0) && pose && flag && geom8, "bad argument");
    PF_CUDA(cudaSetDevice(device));
    AssocTap t;
    PF_CUDA(cudaStreamCreateWithFlags(&t.stream, cudaStreamNonBlocking));
    const int mc = m > 0 ? m : 1, qc = q > 0 ? q : 1;
    PF_CHECK(workspace_create(t.ws, mc, t.stream));
    PF_CUDA(cudaMalloc(&t.d_map, sizeof(Pt) * mc));
    PF_CUDA(cudaMalloc(&t.d_q, sizeof(Pt) * qc));
    PF_CUDA(cudaMalloc(&t.d_pts, sizeof(float4) * mc));
    PF_CUDA(cudaMalloc(&t.d_cs, sizeof(int) * (size_t)kGridCellCap));
    PF_CUDA(cudaMalloc(&t.d_ce, sizeof(int) * (size_t)kGridCellCap));
    PF_CUDA(cudaMalloc(&t.d_geom, sizeof(int) * 12));
    PF_CUDA(cudaMalloc(&t.d_counts, sizeof(int) * 4));
    PF_CUDA(cudaMalloc(&t.d_head, sizeof(int) * mc));
    PF_CUDA(cudaMalloc(&t.d_hits, sizeof(int) * mc));
    PF_CUDA(cudaMalloc(&t.d_next, sizeof(int) * 5 * qc));
    PF_CUDA(cudaMalloc(&t.d_nn, sizeof(int) * 5 * qc));
    PF_CUDA(cudaMalloc(&t.d_flag, qc));
    PF_CUDA(cudaMalloc(&t.d_g8, sizeof(double) * 8 * qc));
    PF_CUDA(cudaMalloc(&t.d_pose, sizeof(double) * 7));
    int counts[4] = {m, 0, q, 0};   // [0] map size, [1] empty map of the other kind, [2] queries, [3] no queries of the other kind
    PF_CUDA(cudaMemcpyAsync(t.d_counts, counts, sizeof(counts), cudaMemcpyHostToDevice, t.stream));
    if (m) PF_CUDA(cudaMemcpyAsync(t.d_map, map, sizeof(Pt) * m, cudaMemcpyHostToDevice, t.stream));
    if (q) PF_CUDA(cudaMemcpyAsync(t.d_q, queries, sizeof(Pt) * q, cudaMemcpyHostToDevice, t.stream));
    PF_CUDA(cudaMemcpyAsync(t.d_pose, pose, sizeof(double) * 7, cudaMemcpyHostToDevice, t.stream));
    PF_CUDA(cudaMemsetAsync(t.d_head, 0xff, sizeof(int) * mc, t.stream));
    PF_CUDA(cudaMemsetAsync(t.d_hits, 0, sizeof(int) * mc, t.stream));
    PF_CUDA(cudaMemsetAsync(t.d_g8, 0, sizeof(double) * 8 * qc, t.stream));
    GridBuild G{};
    G.map[0] = t.d_map; G.map[1] = t.d_map;
    G.n_map[0] = t.d_counts; G.n_map[1] = t.d_counts + 1;
    G.pts[0] = t.d_pts; G.pts[1] = t.d_pts;
    G.cell_start[0] = t.d_cs; G.cell_start[1] = t.d_cs;
    G.cell_end[0] = t.d_ce; G.cell_end[1] = t.d_ce;
    G.geom[0] = t.d_geom; G.geom[1] = t.d_geom + 6;
    PF_CHECK(workspace_begin_step(t.ws));
    PF_CHECK(build_grids(t.ws, G, 0, mc, 0));
    AssocParams P{};
    AssocCloud live{t.d_q, t.d_counts + 2, t.d_map, t.d_counts, KnnGrid{t.d_pts, t.d_cs, t.d_ce, t.d_geom}, t.d_head, t.d_hits, t.d_next, t.d_nn,
                    t.d_flag, t.d_g8, nullptr, nullptr};
    AssocCloud dead = live;
    dead.n_q = t.d_counts + 3;
    dead.n_map = t.d_counts + 1;
    P.c[kind] = live;
    P.c[1 - kind] = dead;
    P.pose = t.d_pose;
    P.k_new = k_new; P.theta_p = theta_p; P.theta_max = theta_max;
    P.guard = nullptr;
    P.weight_type = 0; P.w_minmax = nullptr;
    uint64_t launches = 0;
    PF_CHECK(associate_pass(t.stream, P, kind == 0 ? qc : 0, kind == 1 ? qc : 0, &launches));
    unsigned err = 0;
    PF_CUDA(cudaMemcpyAsync(&err, t.ws.ctrl + kSlotBase + 15, sizeof(unsigned), cudaMemcpyDeviceToHost, t.stream));
    if (m) PF_CUDA(cudaMemcpyAsync(map, t.d_map, sizeof(Pt) * m, cudaMemcpyDeviceToHost, t.stream));
    if (q) {
        PF_CUDA(cudaMemcpyAsync(queries, t.d_q, sizeof(Pt) * q, cudaMemcpyDeviceToHost, t.stream));
        PF_CUDA(cudaMemcpyAsync(flag, t.d_flag, q, cudaMemcpyDeviceToHost, t.stream));
        PF_CUDA(cudaMemcpyAsync(geom8, t.d_g8, sizeof(double) * 8 * q, cudaMemcpyDeviceToHost, t.stream));
    }
    PF_CUDA(cudaStreamSynchronize(t.stream));
    if (err) { set_error("map extent exceeds the search grid capacity"); return PF_ERR_CAPACITY; }
    return PF_OK;
}
