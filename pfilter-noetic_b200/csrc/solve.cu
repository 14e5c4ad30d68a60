// K7 / K8.  See solve.cuh.
//
// k_lm_solve: every thread walks a grid-stride slice of the residual blocks of both kinds, evaluates
//   edge (EdgeAnalyticCostFunction, src/lidarOptimization.cpp:12-46):  lp = q p + t, nu = (lp-a) x (lp-b), r = |nu| / |a-b|,
//        J = -nu^T/|nu| [a-b]x [ -[lp]x  I ] / |a-b|
//   surf (SurfNormAnalyticCostFunction, :56-78):                        r = n . lp + d,  J = n^T [ -[lp]x  I ]
//   applies ceres::HuberLoss(0.1) with Ceres' corrector (rho'' <= 0 -> scale r and J by sqrt(rho')), accumulates the 21
//   upper entries of J^T J, the 6 of J^T r and the cost in fp64 registers, reduces by warp shuffles + shared memory to one
//   partial per CTA; rank 0 of the 8-CTA cluster sums the partials in a fixed order through distributed shared memory
//   (deterministic) and runs the Levenberg-Marquardt state machine, so a whole solve (<= 5 evaluations) is one launch
//   and nothing returns to the host.
#include <cooperative_groups.h>

#include <vector>

#include "solve.cuh"
#include "math.cuh"

namespace pf {

constexpr int kAcc = 29;   // 21 H + 6 g + cost + count

__device__ __forceinline__ void eval_one(int kind, D3 p, const double* ge, const double* pose, double acc[kAcc]) {
    const D3 lp = pose_apply(pose, p);
    double r, J[6];
    if (kind == 0) {
        const D3 a = d3(ge[0], ge[1], ge[2]), b = d3(ge[3], ge[4], ge[5]);
        const D3 nu = cross3(lp - a, lp - b);
        const D3 de = a - b;
        const double de_norm = norm3(de), nu_norm = norm3(nu);
        r = nu_norm / de_norm;
        const D3 w = (-1.0 / nu_norm) * nu;
        const D3 m = cross3(w, de);          // w^T [de]x
        const D3 jr = cross3(lp, m);         // m^T (-[lp]x)
        J[0] = jr.x / de_norm; J[1] = jr.y / de_norm; J[2] = jr.z / de_norm;
        J[3] = m.x / de_norm; J[4] = m.y / de_norm; J[5] = m.z / de_norm;
    } else {
        const D3 n = d3(ge[0], ge[1], ge[2]);
        r = dot3(n, lp) + ge[3];
        const D3 jr = cross3(lp, n);
        J[0] = jr.x; J[1] = jr.y; J[2] = jr.z; J[3] = n.x; J[4] = n.y; J[5] = n.z;
    }
    // HuberLoss(0.1): rho(s), s = r^2
    const double s = r * r;
    double rho0, rho1;
    if (s > 0.01) {
        const double rs = sqrt(s);
        rho0 = 0.2 * rs - 0.01;
        rho1 = fmax(2.2250738585072014e-308, 0.1 / rs);
    } else {
        rho0 = s; rho1 = 1.0;
    }
    const double sc = sqrt(rho1);
    r *= sc;
#pragma unroll
    for (int k = 0; k < 6; ++k) J[k] *= sc;
    int t = 0;
#pragma unroll
    for (int a = 0; a < 6; ++a) {
#pragma unroll
        for (int b = a; b < 6; ++b) acc[t++] += J[a] * J[b];
    }
#pragma unroll
    for (int a = 0; a < 6; ++a) acc[21 + a] += J[a] * r;
    acc[27] += 0.5 * rho0;
    acc[28] += 1.0;
}

__device__ double vec_norm7(const double* x) {
    double s = 0;
    for (int k = 0; k < 7; ++k) s += x[k] * x[k];
    return sqrt(s);
}

// gradient_max_norm of Ceres: | x - Plus(x, -g) |_inf.  The tolerance is 1e-10: the exponential map is only evaluated
// when the plain max-norm of g is anywhere near it (|x - Plus(x,-g)| <= |g| (1 + |x|) for such tiny g).
__device__ double gradient_max_norm(const double* x, const double* g) {
    double gm = 0;
#pragma unroll
    for (int k = 0; k < 6; ++k) gm = fmax(gm, fabs(g[k]));
    if (gm > 1e-6) return gm;
    double ng[6], xp[7];
    for (int k = 0; k < 6; ++k) ng[k] = -g[k];
    se3_plus(x, ng, xp);
    double m = 0;
    for (int k = 0; k < 7; ++k) m = fmax(m, fabs(x[k] - xp[k]));
    return m;
}

__device__ void lm_finish(const LmParams& P, LmState* S) {
    S->phase = 2;
    if (P.iter_poses && S->pass < 16)
        for (int k = 0; k < 7; ++k) P.iter_poses[7 * S->pass + k] = S->x[k];
}

// Compute the next trust-region step from (H, g) at x; invalid steps shrink the radius without an evaluation.
__device__ void lm_propose(const LmParams& P, LmState* S) {
    while (true) {
        if (S->iter >= 4) { lm_finish(P, S); return; }               // max_num_iterations = 4 (:265)
        S->iter += 1;
        double Hs[21], gs[6], sc[6];
#pragma unroll
        for (int a = 0; a < 6; ++a) sc[a] = S->scale[a];
        {
            int t = 0;
#pragma unroll
            for (int a = 0; a < 6; ++a)
#pragma unroll
                for (int b = a; b < 6; ++b) { Hs[t] = S->H[t] * sc[a] * sc[b]; ++t; }
        }
#pragma unroll
        for (int a = 0; a < 6; ++a) gs[a] = S->g[a] * sc[a];
        constexpr int dpos[6] = {0, 6, 11, 15, 18, 20};
        if (!S->reuse_diag) {
#pragma unroll
            for (int a = 0; a < 6; ++a) S->diag[a] = fmin(fmax(Hs[dpos[a]], 1e-6), 1e32);   // min/max_lm_diagonal
        }
        double A[21];
        {
            const double inv_radius = 1.0 / S->radius;
#pragma unroll
            for (int k = 0; k < 21; ++k) A[k] = Hs[k];
#pragma unroll
            for (int a = 0; a < 6; ++a) A[dpos[a]] += S->diag[a] * inv_radius;
        }
        double y[6];
        bool ok = chol6_solve(A, gs, y);                              // (J^T J + D^T D) y = J^T r ; step = -y
        S->reuse_diag = 1;
        double mcc = 0;
        if (ok) {
            double st[6];
#pragma unroll
            for (int a = 0; a < 6; ++a) st[a] = -y[a];
            // model_cost_change = -(step . g_s + 1/2 step^T H_s step)
            double sg = 0, sHs = 0;
            int k = 0;
#pragma unroll
            for (int a = 0; a < 6; ++a) {
                sg += st[a] * gs[a];
#pragma unroll
                for (int b = a; b < 6; ++b) { sHs += (a == b ? 1.0 : 2.0) * st[a] * Hs[k] * st[b]; ++k; }
            }
            mcc = -(sg + 0.5 * sHs);
#pragma unroll
            for (int a = 0; a < 6; ++a) S->step[a] = st[a];
        }
        if (!ok || !(mcc > 0)) {                                      // invalid step == rejected with zero quality
            S->radius /= S->decrease;
            S->decrease *= 2.0;
            if (S->radius < 1e-32) { lm_finish(P, S); return; }
            continue;
        }
        S->model_cost_change = mcc;
        double delta[6];
        for (int a = 0; a < 6; ++a) delta[a] = S->step[a] * S->scale[a];
        se3_plus(S->x, delta, S->xc);
        S->phase = 1;
        return;
    }
}

__device__ __noinline__ void lm_advance(const LmParams& P, LmState* S, const double* sum) {
    for (int k = 0; k < 21; ++k) S->last_H[k] = sum[k];
    for (int k = 0; k < 6; ++k) S->last_g[k] = sum[21 + k];
    S->last_cost = sum[27];
    S->n_res = (int)sum[28];
    if (P.eval_only) { S->phase = 2; return; }
    if (S->phase == 0) {
        if (S->n_res == 0) { lm_finish(P, S); return; }               // no residual blocks: parameters untouched
        for (int k = 0; k < 21; ++k) S->H[k] = sum[k];
        for (int k = 0; k < 6; ++k) S->g[k] = sum[21 + k];
        S->cost = sum[27];
        const int dpos[6] = {0, 6, 11, 15, 18, 20};
        for (int a = 0; a < 6; ++a) S->scale[a] = 1.0 / (1.0 + sqrt(S->H[dpos[a]]));   // jacobi_scaling
        if (gradient_max_norm(S->x, S->g) <= 1e-10) { lm_finish(P, S); return; }
        S->radius = 1e4; S->decrease = 2.0; S->reuse_diag = 0; S->iter = 0;
        S->x_norm = vec_norm7(S->x);
        lm_propose(P, S);
        return;
    }
    // phase 1: the candidate xc has been evaluated
    const double cand = sum[27];
    double sn = 0;
    for (int k = 0; k < 7; ++k) sn += (S->x[k] - S->xc[k]) * (S->x[k] - S->xc[k]);
    if (sqrt(sn) <= 1e-8 * (S->x_norm + 1e-8)) { lm_finish(P, S); return; }            // parameter_tolerance
    const double cost_change = S->cost - cand;
    if (fabs(cost_change) <= 1e-6 * S->cost) { lm_finish(P, S); return; }               // function_tolerance
    const double rel = cost_change / S->model_cost_change;
    if (rel > 1e-3) {                                                                   // min_relative_decrease
        for (int k = 0; k < 7; ++k) S->x[k] = S->xc[k];
        S->x_norm = vec_norm7(S->x);
        for (int k = 0; k < 21; ++k) S->H[k] = sum[k];
        for (int k = 0; k < 6; ++k) S->g[k] = sum[21 + k];
        S->cost = cand;
        const double q = 2.0 * rel - 1.0;
        S->radius = fmin(1e16, S->radius / fmax(1.0 / 3.0, 1.0 - q * q * q));
        S->decrease = 2.0;
        S->reuse_diag = 0;
        if (gradient_max_norm(S->x, S->g) <= 1e-10) { lm_finish(P, S); return; }
    } else {
        S->radius /= S->decrease;
        S->decrease *= 2.0;
    }
    if (S->radius < 1e-32) { lm_finish(P, S); return; }
    lm_propose(P, S);
}

// One launch = one complete solve (ceres::Solve with max_num_iterations = 4): a cluster of kLmCluster CTAs evaluates the
// residual blocks (one grid-stride sweep per evaluation), reduces the 29 sums deterministically (warp shuffles -> shared
// memory -> rank 0 reads the CTA partials of its peers through distributed shared memory, fixed order) and thread 0 of
// rank 0 runs the trust-region state machine on a shared-memory copy of the state; the next candidate pose travels back to
// the peers through the same distributed shared memory.  Two cluster barriers per evaluation, nothing returns to the host.
__global__ void __cluster_dims__(kLmCluster, 1, 1) __launch_bounds__(kLmThreads) k_lm_solve(LmParams P, const double* pose_src, int first_pass) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned rank = cluster.block_rank();
    constexpr int NW = kLmThreads / 32;
    __shared__ double s_red[NW][kAcc];
    __shared__ int s_cnt[NW];
    __shared__ double s_part[32];          // this CTA's partial sums ([31] = edge count)
    __shared__ double s_sum[32];           // rank 0: cluster totals
    __shared__ LmState s_state;            // rank 0: the solver state
    __shared__ double s_pose[8];           // pose to evaluate at; [7] = phase (as double) -- peers read rank 0's copy
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;

    if (rank == 0) {
        const unsigned* src = reinterpret_cast<const unsigned*>(P.state);
        unsigned* dst = reinterpret_cast<unsigned*>(&s_state);
        for (unsigned i = tid; i < sizeof(LmState) / 4; i += blockDim.x) dst[i] = src[i];
        __syncthreads();
        if (tid == 0) {   // lm_begin
            if (pose_src) for (int k = 0; k < 7; ++k) s_state.x[k] = pose_src[k];
            s_state.phase = 0;
            s_state.iter = 0;
            s_state.reuse_diag = 0;
            s_state.pass = first_pass ? 0 : s_state.pass + 1;
            for (int k = 0; k < 7; ++k) s_pose[k] = s_state.x[k];
            s_pose[7] = 0.0;
        }
    }
    cluster.sync();
    const double* pose0 = cluster.map_shared_rank(s_pose, 0);
    const double* part_of[kLmCluster];
#pragma unroll
    for (int r = 0; r < kLmCluster; ++r) part_of[r] = cluster.map_shared_rank(s_part, r);

    const int gtid = rank * blockDim.x + tid, gstride = kLmCluster * blockDim.x;
    for (int round = 0; round < 8; ++round) {
        double pose[7];
#pragma unroll
        for (int k = 0; k < 7; ++k) pose[k] = pose0[k];
        if (pose0[7] != 0.0) break;           // finished (uniform across the cluster: read after a cluster barrier)
        double acc[kAcc];
#pragma unroll
        for (int k = 0; k < kAcc; ++k) acc[k] = 0.0;
        int cnt_edge = 0;
#pragma unroll
        for (int kind = 0; kind < 2; ++kind) {
            const ResidualSrc& R = P.src[kind];
            const int n = R.n ? *R.n : 0;
            for (int i = gtid; i < n; i += gstride) {
                if (R.flag[i] != 2) continue;
                D3 p;
                if (R.p_override) p = d3(R.p_override[3 * i], R.p_override[3 * i + 1], R.p_override[3 * i + 2]);
                else { const Pt q = R.queries[i]; p = d3((double)q.x, (double)q.y, (double)q.z); }
                eval_one(kind, p, R.geom + 8 * (size_t)i, pose, acc);
                if (kind == 0) ++cnt_edge;
            }
        }
        // warp reduction, fixed tree (skipped by warps that evaluated nothing: their partial is exactly zero)
        if (__any_sync(0xffffffffu, acc[28] != 0.0)) {
#pragma unroll
            for (int k = 0; k < kAcc; ++k) {
#pragma unroll
                for (int o = 16; o >= 1; o >>= 1) acc[k] += __shfl_down_sync(0xffffffffu, acc[k], o);
            }
        }
        cnt_edge = __reduce_add_sync(0xffffffffu, cnt_edge);
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < kAcc; ++k) s_red[w][k] = acc[k];
            s_cnt[w] = cnt_edge;
        }
        __syncthreads();
        if (tid < kAcc) {
            double s = 0;
#pragma unroll
            for (int k = 0; k < NW; ++k) s += s_red[k][tid];
            s_part[tid] = s;
        }
        if (tid == 31) {
            int c = 0;
#pragma unroll
            for (int k = 0; k < NW; ++k) c += s_cnt[k];
            s_part[31] = (double)c;
        }
        cluster.sync();                        // all partials visible
        if (rank == 0) {
            if (tid < 32 && (tid < kAcc || tid == 31)) {
                double s = 0;
#pragma unroll
                for (int r = 0; r < kLmCluster; ++r) s += part_of[r][tid];   // fixed order: deterministic
                s_sum[tid] = s;
            }
            __syncthreads();
            if (tid == 0) {
                s_state.n_edge_res = (int)s_sum[31];
                s_state.n_surf_res = (int)s_sum[28] - (int)s_sum[31];
                lm_advance(P, &s_state, s_sum);
                const double* nx = s_state.phase == 1 ? s_state.xc : s_state.x;
                for (int k = 0; k < 7; ++k) s_pose[k] = nx[k];
                s_pose[7] = s_state.phase == 2 ? 1.0 : 0.0;
            }
        }
        cluster.sync();                        // next pose / phase visible; partials may be overwritten
    }
    if (rank == 0) {
        __syncthreads();
        unsigned* dst = reinterpret_cast<unsigned*>(P.state);
        const unsigned* src = reinterpret_cast<const unsigned*>(&s_state);
        for (unsigned i = tid; i < sizeof(LmState) / 4; i += blockDim.x) dst[i] = src[i];
    }
    cluster.sync();                            // peers must not exit while rank 0 may still read their shared memory
}

int lm_solve(cudaStream_t stream, const LmParams& P, const double* pose_src, int first_pass, uint64_t* launches) {
    k_lm_solve<<<kLmCluster, kLmThreads, 0, stream>>>(P, pose_src, first_pass);
    if (launches) *launches += 1;
    PF_CUDA(cudaGetLastError());
    return PF_OK;
}

}  // namespace pf

// ------------------------------------------------------------------------------------------------------------
// stage taps
// ------------------------------------------------------------------------------------------------------------
using namespace pf;

namespace {
struct SolveTap {
    cudaStream_t stream = nullptr;
    double *d_p[2] = {nullptr, nullptr}, *d_geom[2] = {nullptr, nullptr};
    uint8_t* d_flag[2] = {nullptr, nullptr};
    int* d_n = nullptr;
    LmState* d_state = nullptr;
    ~SolveTap() {
        for (int k = 0; k < 2; ++k) { cudaFree(d_p[k]); cudaFree(d_geom[k]); cudaFree(d_flag[k]); }
        cudaFree(d_n); cudaFree(d_state);
        if (stream) cudaStreamDestroy(stream);
    }
};

int solve_tap_setup(SolveTap& t, int device, const double pose[7], const double* edge9, int ne, const double* surf7, int ns, LmParams& P) {
    PF_REQUIRE(pose && ne >= 0 && ns >= 0 && (edge9 || ne == 0) && (surf7 || ns == 0), "bad argument");
    PF_CUDA(cudaSetDevice(device));
    PF_CUDA(cudaStreamCreateWithFlags(&t.stream, cudaStreamNonBlocking));
    const int n[2] = {ne, ns};
    std::vector<double> hp[2], hg[2];
    for (int i = 0; i < ne; ++i) {
        const double* e = edge9 + 9 * (size_t)i;
        hp[0].insert(hp[0].end(), e, e + 3);
        const double g8[8] = {e[3], e[4], e[5], e[6], e[7], e[8], 0, 0};
        hg[0].insert(hg[0].end(), g8, g8 + 8);
    }
    for (int i = 0; i < ns; ++i) {
        const double* s = surf7 + 7 * (size_t)i;
        hp[1].insert(hp[1].end(), s, s + 3);
        const double g8[8] = {s[3], s[4], s[5], s[6], 0, 0, 0, 0};
        hg[1].insert(hg[1].end(), g8, g8 + 8);
    }
    PF_CUDA(cudaMalloc(&t.d_n, sizeof(int) * 2));
    PF_CUDA(cudaMemcpyAsync(t.d_n, n, sizeof(n), cudaMemcpyHostToDevice, t.stream));
    for (int k = 0; k < 2; ++k) {
        const int c = n[k] > 0 ? n[k] : 1;
        PF_CUDA(cudaMalloc(&t.d_p[k], sizeof(double) * 3 * c));
        PF_CUDA(cudaMalloc(&t.d_geom[k], sizeof(double) * 8 * c));
        PF_CUDA(cudaMalloc(&t.d_flag[k], c));
        PF_CUDA(cudaMemsetAsync(t.d_flag[k], 2, c, t.stream));
        if (n[k]) {
            PF_CUDA(cudaMemcpyAsync(t.d_p[k], hp[k].data(), sizeof(double) * 3 * n[k], cudaMemcpyHostToDevice, t.stream));
            PF_CUDA(cudaMemcpyAsync(t.d_geom[k], hg[k].data(), sizeof(double) * 8 * n[k], cudaMemcpyHostToDevice, t.stream));
        }
        P.src[k] = ResidualSrc{nullptr, t.d_p[k], t.d_flag[k], t.d_geom[k], t.d_n + k};
    }
    PF_CUDA(cudaMalloc(&t.d_state, sizeof(LmState)));
    PF_CUDA(cudaMemsetAsync(t.d_state, 0, sizeof(LmState), t.stream));
    PF_CUDA(cudaMemcpyAsync(t.d_state->x, pose, sizeof(double) * 7, cudaMemcpyHostToDevice, t.stream));
    P.state = t.d_state;
    P.iter_poses = nullptr;
    PF_CUDA(cudaStreamSynchronize(t.stream));   // the staging vectors go out of scope
    return PF_OK;
}
}  // namespace

extern "C" int pf_eval_normal_eq(int device, const double pose[7], const double* edge9, int n_edge, const double* surf7, int n_surf,
                                 double H21[21], double g6[6], double* cost) {
    PF_REQUIRE(H21 && g6 && cost, "null output");
    SolveTap t;
    LmParams P{};
    PF_CHECK(solve_tap_setup(t, device, pose, edge9, n_edge, surf7, n_surf, P));
    P.eval_only = 1;
    PF_CHECK(lm_solve(t.stream, P, nullptr, 1, nullptr));
    LmState S;
    PF_CUDA(cudaMemcpyAsync(&S, t.d_state, sizeof(S), cudaMemcpyDeviceToHost, t.stream));
    PF_CUDA(cudaStreamSynchronize(t.stream));
    memcpy(H21, S.last_H, sizeof(double) * 21);
    memcpy(g6, S.last_g, sizeof(double) * 6);
    *cost = S.last_cost;
    return PF_OK;
}

extern "C" int pf_lm_solve(int device, double pose_io[7], const double* edge9, int n_edge, const double* surf7, int n_surf, int* iterations,
                           double* final_cost) {
    PF_REQUIRE(pose_io, "null pose");
    SolveTap t;
    LmParams P{};
    PF_CHECK(solve_tap_setup(t, device, pose_io, edge9, n_edge, surf7, n_surf, P));
    P.eval_only = 0;
    PF_CHECK(lm_solve(t.stream, P, nullptr, 1, nullptr));
    LmState S;
    PF_CUDA(cudaMemcpyAsync(&S, t.d_state, sizeof(S), cudaMemcpyDeviceToHost, t.stream));
    PF_CUDA(cudaStreamSynchronize(t.stream));
    memcpy(pose_io, S.x, sizeof(double) * 7);
    if (iterations) *iterations = S.iter;
    if (final_cost) *final_cost = S.cost;
    return PF_OK;
}
