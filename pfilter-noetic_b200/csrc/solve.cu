// K7 / K8.  See solve.cuh.
//
// k_lm_solve: every thread walks a grid-stride slice of the residual blocks of both kinds, evaluates
//   edge (EdgeAnalyticCostFunction, src/lidarOptimization.cpp:12-46):  lp = q p + t, nu = (lp-a) x (lp-b), r = |nu| / |a-b|,
//        J = -nu^T/|nu| [a-b]x [ -[lp]x  I ] / |a-b|
//   surf (SurfNormAnalyticCostFunction, :56-78):                        r = n . lp + d,  J = n^T [ -[lp]x  I ]
//   applies ceres::HuberLoss(0.1) with Ceres' corrector (rho'' <= 0 -> scale r and J by sqrt(rho')), accumulates the 21
//   upper entries of J^T J, the 6 of J^T r and the cost in fp64 registers, reduces by warp shuffles + shared memory to one
//   partial per CTA; every CTA of the 12-CTA cluster sums the partials in a fixed order through distributed shared memory
//   (deterministic) and runs the Levenberg-Marquardt state machine, so a whole solve (<= 5 evaluations) is one launch
//   and nothing returns to the host.
#include <cooperative_groups.h>

#include <vector>

#include "solve.cuh"
#include "math.cuh"

namespace pf {

constexpr int kAcc = 29;   // 21 H + 6 g + cost + count

// One residual block at the pose (R, t).  Edge: with u = (a - b) / |a - b| the reference's nu = (lp-a) x (lp-b) equals
// |a-b| (lp-a) x u, so r = |(lp-a) x u| and J = w^T [u]x [ -[lp]x  I ] with w = -nu/|nu|: the same numbers as
// src/lidarOptimization.cpp:18-41 with one square root and two reciprocals on the critical path instead of eleven fp64
// divisions / roots.  Huber: sqrt(s) = |r|.
// The residual weights of weightType 1 / 2 / 12: min-max normalised observe (observeMean, src/odomEstimationClass.cpp:136-160) and
// point sparsity (pointSparsityMean, include/odomEstimationClass.h:111-126).  mm = {min obs, max obs, min spa, max spa}.
__device__ __forceinline__ double residual_weight(int weight_type, double obs, double spa, const double* mm) {
    if (weight_type == 1 || weight_type == 12) {
        const double len = mm[1] - mm[0];
        if (len != 0) { obs = (obs - mm[0]) / len; obs -= 1.0; obs = fabs(obs); obs *= 2.0; obs = fmax(0.1, obs); }
    }
    if (weight_type == 2 || weight_type == 12) {
        const double len = mm[3] - mm[2];
        if (len != 0) { spa = (spa - mm[2]) / len; spa -= 1.0; spa = fabs(spa); spa *= 2.0; }
    }
    return weight_type == 1 ? obs : weight_type == 2 ? spa : (spa + obs) / 2;
}

// `weight` scales the RESIDUAL only, never the Jacobian, and only when the functor's own test passes: the edge functor
// wants exactly 1, 2 or 12 (src/lidarOptimization.cpp:25-28), the surf functor anything but 0 (:62-63).
// kFma: accumulate J^T J and J^T r with fused multiply-adds (the library is built with -fmad=false so that every product rounds like
// the CPU reference; the map-sweep kernel, whose sums are compared at 1e-10 and whose fp64 issue rate matters, may fuse the 27
// accumulations -- the default there, -DPF_NE_NOFMA builds it without)
template <bool kFma = false>
__device__ __forceinline__ void eval_one(int kind, D3 p, const double* ge, const double* Rm, const double* tv, double weight, double acc[kAcc]) {
    const D3 lp = d3(Rm[0] * p.x + Rm[1] * p.y + Rm[2] * p.z + tv[0], Rm[3] * p.x + Rm[4] * p.y + Rm[5] * p.z + tv[1],
                     Rm[6] * p.x + Rm[7] * p.y + Rm[8] * p.z + tv[2]);
    double r, ar, J[6];
    if (kind == 0) {
        const D3 a = d3(ge[0], ge[1], ge[2]), b = d3(ge[3], ge[4], ge[5]);
        const D3 de = a - b;
        const double inv_de = 1.0 / norm3(de);
        const D3 u = inv_de * de;
        const D3 nu = cross3(lp - a, u);
        r = norm3(nu);
        const D3 w = (-1.0 / r) * nu;
        const D3 m = cross3(w, u);           // w^T [u]x
        const D3 jr = cross3(lp, m);         // m^T (-[lp]x)
        J[0] = jr.x; J[1] = jr.y; J[2] = jr.z; J[3] = m.x; J[4] = m.y; J[5] = m.z;
        if (weight == 1 || weight == 2 || weight == 12) r = weight * r;
    } else {
        const D3 n = d3(ge[0], ge[1], ge[2]);
        r = dot3(n, lp) + ge[3];
        if (weight != 0) r = weight * r;
        const D3 jr = cross3(lp, n);
        J[0] = jr.x; J[1] = jr.y; J[2] = jr.z; J[3] = n.x; J[4] = n.y; J[5] = n.z;
    }
    ar = fabs(r);
    // HuberLoss(0.1): rho(s), s = r^2
    const double s = r * r;
    double rho0, sc;
    if (s > 0.01) {
        rho0 = 0.2 * ar - 0.01;
        sc = sqrt(fmax(2.2250738585072014e-308, 0.1 / ar));
    } else {
        rho0 = s; sc = 1.0;
    }
    r *= sc;
#pragma unroll
    for (int k = 0; k < 6; ++k) J[k] *= sc;
    int t = 0;
#pragma unroll
    for (int a = 0; a < 6; ++a) {
#pragma unroll
        for (int b = a; b < 6; ++b) { acc[t] = kFma ? fma(J[a], J[b], acc[t]) : acc[t] + J[a] * J[b]; ++t; }
    }
#pragma unroll
    for (int a = 0; a < 6; ++a) acc[21 + a] = kFma ? fma(J[a], r, acc[21 + a]) : acc[21 + a] + J[a] * r;
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
// one copy of the exponential map (two sincos, ~10 KB inlined) for the step and the gradient test
__device__ __noinline__ void se3_plus_once(const double* x, const double* d, double* out) { se3_plus(x, d, out); }

__device__ __noinline__ double gradient_max_norm(const double* x, const double* g) {
    double gm = 0;
#pragma unroll
    for (int k = 0; k < 6; ++k) gm = fmax(gm, fabs(g[k]));
    if (gm > 1e-6) return gm;
    double ng[6], xp[7];
    for (int k = 0; k < 6; ++k) ng[k] = -g[k];
    se3_plus_once(x, ng, xp);
    double m = 0;
    for (int k = 0; k < 7; ++k) m = fmax(m, fabs(x[k] - xp[k]));
    return m;
}

__device__ __noinline__ void lm_finish(const LmParams& P, LmState* S, bool writer = true) {
    S->phase = 2;
    if (writer && P.iter_poses && S->pass < 16)
        for (int k = 0; k < 7; ++k) P.iter_poses[7 * S->pass + k] = S->x[k];
}

// Compute the next trust-region step from (H, g) at x; invalid steps shrink the radius without an evaluation.
// Out of line: the step is called from two places of lm_advance and run by ONE thread while the cluster waits; one copy of its
// ~1000 straight-line fp64 instructions measured 1-2 % faster end to end than two (a rolled, local-memory Cholesky: 10 % slower).
__device__ __noinline__ void lm_propose(const LmParams& P, LmState* S, bool writer) {
    while (true) {
        if (S->iter >= 4) { lm_finish(P, S, writer); return; }               // max_num_iterations = 4 (:265)
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
            if (S->radius < 1e-32) { lm_finish(P, S, writer); return; }
            continue;
        }
        S->model_cost_change = mcc;
        double delta[6];
        for (int a = 0; a < 6; ++a) delta[a] = S->step[a] * S->scale[a];
        se3_plus_once(S->x, delta, S->xc);
        S->phase = 1;
        return;
    }
}

__device__ __noinline__ void lm_advance(const LmParams& P, LmState* S, const double* sum, bool writer) {
    for (int k = 0; k < 21; ++k) S->last_H[k] = sum[k];
    for (int k = 0; k < 6; ++k) S->last_g[k] = sum[21 + k];
    S->last_cost = sum[27];
    S->n_res = (int)sum[28];
    if (P.eval_only) { S->phase = 2; return; }
    if (S->phase == 0) {
        if (S->n_res == 0) { lm_finish(P, S, writer); return; }               // no residual blocks: parameters untouched
        for (int k = 0; k < 21; ++k) S->H[k] = sum[k];
        for (int k = 0; k < 6; ++k) S->g[k] = sum[21 + k];
        S->cost = sum[27];
        const int dpos[6] = {0, 6, 11, 15, 18, 20};
        for (int a = 0; a < 6; ++a) S->scale[a] = 1.0 / (1.0 + sqrt(S->H[dpos[a]]));   // jacobi_scaling
        if (gradient_max_norm(S->x, S->g) <= 1e-10) { lm_finish(P, S, writer); return; }
        S->radius = 1e4; S->decrease = 2.0; S->reuse_diag = 0; S->iter = 0;
        S->x_norm = vec_norm7(S->x);
        lm_propose(P, S, writer);
        return;
    }
    // phase 1: the candidate xc has been evaluated
    const double cand = sum[27];
    double sn = 0;
    for (int k = 0; k < 7; ++k) sn += (S->x[k] - S->xc[k]) * (S->x[k] - S->xc[k]);
    if (sqrt(sn) <= 1e-8 * (S->x_norm + 1e-8)) { lm_finish(P, S, writer); return; }            // parameter_tolerance
    const double cost_change = S->cost - cand;
    if (fabs(cost_change) <= 1e-6 * S->cost) { lm_finish(P, S, writer); return; }               // function_tolerance
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
        if (gradient_max_norm(S->x, S->g) <= 1e-10) { lm_finish(P, S, writer); return; }
    } else {
        S->radius /= S->decrease;
        S->decrease *= 2.0;
    }
    if (S->radius < 1e-32) { lm_finish(P, S, writer); return; }
    lm_propose(P, S, writer);
}

// One launch = one complete solve (ceres::Solve with max_num_iterations = 4) by a cluster of kLmCluster CTAs.
//   * CTA r owns the r-th slice of both query clouds and compacts the indices of its residual blocks (flag == 2) into shared
//     memory once, in query order, so every evaluation is a dense loop over them.
//   * One evaluation = every thread sums its blocks, warp shuffle tree, shared memory, ONE cluster barrier; then every CTA
//     adds up the partials of all CTAs in the same fixed order through distributed shared memory (deterministic, identical
//     in every CTA) and runs the trust-region state machine on its own copy of the state -- the next candidate pose never
//     has to travel, so there is one cluster barrier per evaluation (the partials are double buffered by round parity).
// Nothing returns to the host; rank 0 writes the state back at the end.
__global__ void __cluster_dims__(kLmCluster, 1, 1) __launch_bounds__(kLmThreads) k_lm_solve(LmParams P, const double* pose_src, int first_pass,
                                                                                                int list_cap) {
    PF_PDL_ENTRY();
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned rank = cluster.block_rank();
    constexpr int NW = kLmThreads / 32;
    extern __shared__ int s_list[];          // [list_cap] compacted residual blocks of this CTA: kind << 30 | query index
    __shared__ double s_red[NW][kAcc];
    __shared__ int s_cnt[NW];
    __shared__ double s_part[2][32];       // this CTA's partial sums, by round parity ([31] = edge count)
    __shared__ double s_sum[32];           // cluster totals
    __shared__ LmState s_state;            // every CTA keeps the whole solver state
    __shared__ int s_n, s_scan[NW + 1];
    __shared__ double s_mm[4 * kLmMaxSrc]; // weightType != 0: min / max of observe and sparsity per source
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (P.weight_type != 0 && tid < 4 * P.nsrc) s_mm[tid] = __longlong_as_double((long long)P.src[tid >> 2].w_minmax[tid & 3]);

    {
        const unsigned* src = reinterpret_cast<const unsigned*>(P.state);
        unsigned* dst = reinterpret_cast<unsigned*>(&s_state);
        for (unsigned i = tid; i < sizeof(LmState) / 4; i += blockDim.x) dst[i] = src[i];
        if (tid == 0) s_n = 0;
        __syncthreads();
        if (tid == 0) {   // lm_begin
            if (pose_src) for (int k = 0; k < 7; ++k) s_state.x[k] = pose_src[k];
            s_state.phase = 0;
            s_state.iter = 0;
            s_state.reuse_diag = 0;
            s_state.pass = first_pass ? 0 : s_state.pass + 1;
        }
    }
    // compaction of this CTA's slices, order preserving (chunks of kLmThreads entries: ballot + warp counts)
    int overflow = 0;
    for (int kind = 0; kind < P.nsrc; ++kind) {
        const ResidualSrc& R = P.src[kind];
        const int n = R.n ? *R.n : 0;
        const int per = (n + kLmCluster - 1) / kLmCluster;
        const int lo = min(n, (int)rank * per), hi = min(n, lo + per);
        for (int base = lo; base < hi; base += kLmThreads) {
            const int i = base + tid;
            const bool v = i < hi && R.flag[i] == 2;
            const unsigned m = __ballot_sync(0xffffffffu, v);
            if (lane == 0) s_scan[w] = __popc(m);
            __syncthreads();
            int off = s_n, tot = 0;
#pragma unroll
            for (int k = 0; k < NW; ++k) { const int c = s_scan[k]; if (k < w) off += c; tot += c; }
            if (v) {
                const int pos = off + __popc(m & lanemask_lt());
                if (pos < list_cap) s_list[pos] = (int)(((unsigned)kind << 30) | (unsigned)i); else overflow = 1;
            }
            __syncthreads();
            if (tid == 0) s_n += tot;
            __syncthreads();
        }
    }
    const int n_list = min(s_n, list_cap);
    const int my_edges = [&] { int c = 0; for (int k = tid; k < n_list; k += kLmThreads) c += P.src[(unsigned)s_list[k] >> 30].type == 0 ? 1 : 0; return c; }();
    (void)overflow;     // cannot happen: list_cap covers the whole slice (host side)
    const double* part_of[kLmCluster];
#pragma unroll
    for (int r = 0; r < kLmCluster; ++r) part_of[r] = cluster.map_shared_rank(&s_part[0][0], r);
    __syncthreads();

#pragma unroll 1     // (left to itself nvcc unrolls the eight rounds: eight copies of the evaluation, 150 KB of code on a 32 KB instruction cache)
    for (int round = 0; round < 8; ++round) {
        if (s_state.phase == 2) break;        // identical in every CTA
        const double* x = s_state.phase == 1 ? s_state.xc : s_state.x;
        double Rm[9], tv[3];
        quat_to_mat(x, Rm);
        tv[0] = x[4]; tv[1] = x[5]; tv[2] = x[6];
        double acc[kAcc];
#pragma unroll
        for (int k = 0; k < kAcc; ++k) acc[k] = 0.0;
#pragma unroll 1
        for (int k = tid; k < n_list; k += kLmThreads) {
            const unsigned e = (unsigned)s_list[k];
            const int src = (int)(e >> 30), i = (int)(e & 0x3fffffffu);
            const ResidualSrc& R = P.src[src];
            const int kind = R.type;
            D3 p;
            if (R.p_override) p = d3(R.p_override[3 * i], R.p_override[3 * i + 1], R.p_override[3 * i + 2]);
            else { const Pt q = R.queries[i]; p = d3((double)q.x, (double)q.y, (double)q.z); }
            const double weight = P.weight_type != 0 ? residual_weight(P.weight_type, (double)R.w_obs[i], R.w_spa[i], s_mm + 4 * src) : 0.0;
            eval_one(kind, p, R.geom + 8 * (size_t)i, Rm, tv, weight, acc);
        }
        // Warp reduction, fixed tree (skipped by warps that evaluated nothing: their partial is exactly zero).  Transposed: at
        // distance 16 a lane keeps one half of the 32 slots and hands the other half to its partner, at distance 8 a quarter, ...:
        // 31 exchanges instead of 29 x 5, and lane l ends up with the total of slot l.  The pairing of the partial sums is that of
        // the shuffle-down tree this replaces ((l, l ^ 16), then (l, l ^ 8), ...), so the sums are the same bit for bit -- but the
        // unrolled 145-step tree was 55 KB of the kernel's code and a third of the instructions it executes.
        double tot = 0.0;
        if (__any_sync(0xffffffffu, acc[28] != 0.0)) {
            double v16[16], v8[8], v4[4], v2[2];
            const bool b4 = (lane & 16) != 0, b3 = (lane & 8) != 0, b2 = (lane & 4) != 0, b1 = (lane & 2) != 0, b0 = (lane & 1) != 0;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const double lo = acc[i], hi = i + 16 < kAcc ? acc[i + 16] : 0.0;
                v16[i] = (b4 ? hi : lo) + __shfl_xor_sync(0xffffffffu, b4 ? lo : hi, 16);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) v8[i] = (b3 ? v16[i + 8] : v16[i]) + __shfl_xor_sync(0xffffffffu, b3 ? v16[i] : v16[i + 8], 8);
#pragma unroll
            for (int i = 0; i < 4; ++i) v4[i] = (b2 ? v8[i + 4] : v8[i]) + __shfl_xor_sync(0xffffffffu, b2 ? v8[i] : v8[i + 4], 4);
#pragma unroll
            for (int i = 0; i < 2; ++i) v2[i] = (b1 ? v4[i + 2] : v4[i]) + __shfl_xor_sync(0xffffffffu, b1 ? v4[i] : v4[i + 2], 2);
            tot = (b0 ? v2[1] : v2[0]) + __shfl_xor_sync(0xffffffffu, b0 ? v2[0] : v2[1], 1);
        }
        const int cnt_edge = __reduce_add_sync(0xffffffffu, my_edges);
        if (lane < kAcc) s_red[w][lane] = tot;
        if (lane == 0) s_cnt[w] = cnt_edge;
        __syncthreads();
        const int par = round & 1;
        if (tid < kAcc) {
            double sum = 0;
#pragma unroll
            for (int k = 0; k < NW; ++k) sum += s_red[k][tid];
            s_part[par][tid] = sum;
        }
        if (tid == 31) {
            int c = 0;
#pragma unroll
            for (int k = 0; k < NW; ++k) c += s_cnt[k];
            s_part[par][31] = (double)c;
        }
        cluster.sync();                        // all partials of this round visible
        if (tid < 32 && (tid < kAcc || tid == 31)) {
            double sum = 0;
#pragma unroll
            for (int r = 0; r < kLmCluster; ++r) sum += part_of[r][par * 32 + tid];   // fixed order: deterministic
            s_sum[tid] = sum;
        }
        __syncthreads();
        if (tid == 0) {
            s_state.n_edge_res = (int)s_sum[31];
            s_state.n_surf_res = (int)s_sum[28] - (int)s_sum[31];
            lm_advance(P, &s_state, s_sum, rank == 0);
        }
        __syncthreads();
    }
    if (rank == 0) {
        unsigned* dst = reinterpret_cast<unsigned*>(P.state);
        const unsigned* src = reinterpret_cast<const unsigned*>(&s_state);
        for (unsigned i = tid; i < sizeof(LmState) / 4; i += blockDim.x) dst[i] = src[i];
    }
    cluster.sync();                            // peers must not exit while others may still read their shared memory
}

int lm_solve(cudaStream_t stream, const LmParams& P, const double* pose_src, int first_pass, uint64_t* launches, const int* ub) {
    // shared-memory list: the largest slice a CTA can own
    int list_cap = 8;
    for (int k = 0; k < P.nsrc; ++k) list_cap += (ub[k] + kLmCluster - 1) / kLmCluster;
    const size_t smem = sizeof(int) * (size_t)list_cap;
    PF_REQUIRE(P.nsrc >= 1 && P.nsrc <= kLmMaxSrc, "lm_solve: %d residual sources", P.nsrc);
    PF_REQUIRE(smem <= 160 * 1024, "lm_solve: %d residual candidates per CTA exceed the solver's shared-memory list", list_cap);
    // function attributes are per device: a process may hold handles on several
    int dev = 0;
    PF_CUDA(cudaGetDevice(&dev));
    static size_t smem_set[64] = {};
    static bool cluster_set[64] = {};
    const int slot = dev & 63;
    if (kLmCluster > 8 && !cluster_set[slot]) {       // more than the portable cluster size
        PF_CUDA(cudaFuncSetAttribute(k_lm_solve, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        cluster_set[slot] = true;
    }
    if (smem > smem_set[slot]) {
        PF_CUDA(cudaFuncSetAttribute(k_lm_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem > 48 * 1024 ? smem : 48 * 1024)));
        smem_set[slot] = smem > 48 * 1024 ? smem : 48 * 1024;
    }
    PF_CUDA(launch_pdl(k_lm_solve, dim3(kLmCluster), dim3(kLmThreads), smem, stream, P, pose_src, first_pass, list_cap));
    if (launches) *launches += 1;
    PF_CUDA(cudaGetLastError());
    return PF_OK;
}


// ------------------------------------------------------------------------------------------------------------
// K7 at map-sweep sizes (BASELINE.json configs[4]: 10^6 .. 5 x 10^7 residual blocks): the same residual / Jacobian / Huber / J^T J
// arithmetic as one evaluation of k_lm_solve, as a grid-wide streaming kernel.  The per-frame solver is a 12-CTA cluster because a
// frame has ~5 k residual blocks and needs five dependent evaluations; at 10^6 blocks one evaluation is an HBM stream of
// 72 B (edge: p, a, b) / 56 B (surf: p, n, d) per block (src/lidarOptimization.cpp:12-78) and wants every SM.
// Persistent CTAs, static round-robin over tiles of 256 blocks; a tile is one contiguous 18 KB / 14 KB run of the packed input,
// brought into shared memory by ONE bulk (TMA) copy per tile into a 3-deep ring (mbarrier completion), so the loads of the next two
// tiles are in flight while the CTA evaluates the current one; thread i reads block i from shared memory (stride 72 / 56 B:
// conflict-free for 8-byte accesses), accumulates the 29 sums in fp64 registers; shuffle tree + shared memory per CTA, one partial
// per CTA in global memory, and the last CTA to finish (ticket) adds the partials in CTA order -- deterministic.
#ifndef PF_NE_STAGES
#define PF_NE_STAGES 5
#endif
#ifndef PF_NE_CTAS
#define PF_NE_CTAS 2
#endif
constexpr int kNeTile = 256, kNeStages = PF_NE_STAGES, kNeCtasPerSm = PF_NE_CTAS;
constexpr int kNeStageBytes = kNeTile * 72;

struct NeStreamParams {
    const double* edge9; const double* surf7;
    int n_edge, n_surf;
    const double* pose;        // [7] device
    double* partial;           // [grid][32]
    unsigned* ticket;
    double* out;               // [32]: H21, g6, cost, count
};

__global__ void __launch_bounds__(kNeTile, kNeCtasPerSm) k_normal_eq_stream(NeStreamParams P) {
    extern __shared__ __align__(128) unsigned char s_stage[];      // [kNeStages][kNeStageBytes]
    __shared__ __align__(8) unsigned long long s_bar[kNeStages];
    __shared__ double s_red[kNeTile / 32][kAcc];
    __shared__ int s_last;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int tiles_e = (P.n_edge + kNeTile - 1) / kNeTile, tiles_s = (P.n_surf + kNeTile - 1) / kNeTile, ntiles = tiles_e + tiles_s;
    const int G = gridDim.x;
    const int mine = (int)blockIdx.x < ntiles ? (ntiles - 1 - (int)blockIdx.x) / G + 1 : 0;
    // tile j of this CTA -> (kind, first block, blocks)
    auto desc = [&](int j, int& kind, const double*& src, int& cnt) {
        const int t = (int)blockIdx.x + j * G;
        if (t < tiles_e) { kind = 0; src = P.edge9 + (size_t)t * kNeTile * 9; cnt = min(kNeTile, P.n_edge - t * kNeTile); }
        else { const int u = t - tiles_e; kind = 1; src = P.surf7 + (size_t)u * kNeTile * 7; cnt = min(kNeTile, P.n_surf - u * kNeTile); }
    };
    auto issue = [&](int j) {      // one thread; only full tiles travel by bulk copy (their byte count is a multiple of 16)
        int kind, cnt; const double* src;
        desc(j, kind, src, cnt);
        if (cnt == kNeTile) bulk_load(s_stage + (size_t)(j % kNeStages) * kNeStageBytes, src, (unsigned)(kNeTile * (kind == 0 ? 72 : 56)), &s_bar[j % kNeStages]);
    };
    if (tid == 0) {
        for (int b = 0; b < kNeStages; ++b) mbar_init(&s_bar[b], 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int j = 0; j < kNeStages - 1 && j < mine; ++j) issue(j);
    }
    double Rm[9], tv[3];
    {
        double x[7];
#pragma unroll
        for (int k = 0; k < 7; ++k) x[k] = P.pose[k];
        quat_to_mat(x, Rm);
        tv[0] = x[4]; tv[1] = x[5]; tv[2] = x[6];
    }
    double acc[kAcc];
#pragma unroll
    for (int k = 0; k < kAcc; ++k) acc[k] = 0.0;
    __syncthreads();
    unsigned phases = 0u;      // bit b: parity the next wait on stage b expects.  Only FULL tiles travel through a stage's barrier (a ragged
                               // tile -- the last edge tile may sit in the middle of a CTA's sequence -- neither arrives nor waits), so the
                               // parity is counted per completed wait, not derived from the tile number
    for (int j = 0; j < mine; ++j) {
        if (tid == 0 && j + kNeStages - 1 < mine) issue(j + kNeStages - 1);      // its buffer was released by the barrier below
        int kind, cnt; const double* src;
        desc(j, kind, src, cnt);
        const int nd = kind == 0 ? 9 : 7;
        double v[9];
        if (cnt == kNeTile) {
            const int st = j % kNeStages;
            mbar_wait(&s_bar[st], (phases >> st) & 1u);
            phases ^= 1u << st;
            const double* sp = reinterpret_cast<const double*>(s_stage + (size_t)(j % kNeStages) * kNeStageBytes) + (size_t)tid * nd;
#pragma unroll
            for (int k = 0; k < 9; ++k) v[k] = k < nd ? sp[k] : 0.0;
        } else if (tid < cnt) {      // the ragged last tile of a kind: plain loads
#pragma unroll
            for (int k = 0; k < 9; ++k) v[k] = k < nd ? src[(size_t)tid * nd + k] : 0.0;
        }
#ifdef PF_NE_NOFMA
        if (tid < cnt) eval_one<false>(kind, d3(v[0], v[1], v[2]), v + 3, Rm, tv, 0.0, acc);
#else       // fused accumulation of the 27 sums: 71 % of the measured HBM peak against 65 % without (profiles/k7_stream_variants_r2.txt)
        if (tid < cnt) eval_one<true>(kind, d3(v[0], v[1], v[2]), v + 3, Rm, tv, 0.0, acc);
#endif
        __syncthreads();
    }
#pragma unroll
    for (int k = 0; k < kAcc; ++k) {
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) acc[k] += __shfl_down_sync(0xffffffffu, acc[k], o);
    }
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < kAcc; ++k) s_red[w][k] = acc[k];
    }
    __syncthreads();
    if (tid < kAcc) {
        double sum = 0;
#pragma unroll
        for (int k = 0; k < kNeTile / 32; ++k) sum += s_red[k][tid];
        P.partial[(size_t)blockIdx.x * 32 + tid] = sum;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(P.ticket, 1u) == (unsigned)G - 1u;
    __syncthreads();
    if (s_last) {
        __threadfence();
        if (tid < kAcc) {
            double sum = 0;
            for (int b = 0; b < G; ++b) sum += __ldcg(P.partial + (size_t)b * 32 + tid);      // CTA order: deterministic
            P.out[tid] = sum;
        }
        if (tid == 0) *P.ticket = 0u;
    }
}

int normal_eq_stream(cudaStream_t stream, const double* d_pose, const double* d_edge9, int n_edge, const double* d_surf7, int n_surf,
                     double* d_partial, unsigned* d_ticket, double* d_out, int grid) {
    static bool attr_set[64] = {};
    int dev = 0;
    PF_CUDA(cudaGetDevice(&dev));
    if (!attr_set[dev & 63]) {
        PF_CUDA(cudaFuncSetAttribute(k_normal_eq_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, kNeStages * kNeStageBytes));
        attr_set[dev & 63] = true;
    }
    NeStreamParams P{d_edge9, d_surf7, n_edge, n_surf, d_pose, d_partial, d_ticket, d_out};
    k_normal_eq_stream<<<grid, kNeTile, kNeStages * kNeStageBytes, stream>>>(P);
    PF_CUDA(cudaGetLastError());
    return PF_OK;
}

}  // namespace pf

// ------------------------------------------------------------------------------------------------------------
// stage taps
// ------------------------------------------------------------------------------------------------------------
using namespace pf;

extern "C" int pf_eval_normal_eq_timed(int device, const double pose[7], const double* edge9, int n_edge, const double* surf7, int n_surf,
                                       int reps, double H21[21], double g6[6], double* cost, float* ms_kernel);
namespace {
constexpr long long kLmTapMax = 262144;    // residual blocks up to which pf_eval_normal_eq runs the per-frame cluster kernel
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
        P.src[k] = ResidualSrc{k, nullptr, t.d_p[k], t.d_flag[k], t.d_geom[k], t.d_n + k, nullptr, nullptr, nullptr};
    }
    PF_CUDA(cudaMalloc(&t.d_state, sizeof(LmState)));
    PF_CUDA(cudaMemsetAsync(t.d_state, 0, sizeof(LmState), t.stream));
    PF_CUDA(cudaMemcpyAsync(t.d_state->x, pose, sizeof(double) * 7, cudaMemcpyHostToDevice, t.stream));
    P.state = t.d_state;
    P.iter_poses = nullptr;
    P.weight_type = 0; P.nsrc = 2;
    PF_CUDA(cudaStreamSynchronize(t.stream));   // the staging vectors go out of scope
    return PF_OK;
}
}  // namespace

extern "C" int pf_eval_normal_eq(int device, const double pose[7], const double* edge9, int n_edge, const double* surf7, int n_surf,
                                 double H21[21], double g6[6], double* cost) {
    PF_REQUIRE(H21 && g6 && cost, "null output");
    if ((long long)n_edge + n_surf > kLmTapMax)       // beyond the cluster kernel's shared-memory lists: the grid-wide kernel
        return pf_eval_normal_eq_timed(device, pose, edge9, n_edge, surf7, n_surf, 1, H21, g6, cost, nullptr);
    SolveTap t;
    LmParams P{};
    PF_CHECK(solve_tap_setup(t, device, pose, edge9, n_edge, surf7, n_surf, P));
    P.eval_only = 1;
    { const int ub[2] = {n_edge, n_surf}; PF_CHECK(lm_solve(t.stream, P, nullptr, 1, nullptr, ub)); }
    LmState S;
    PF_CUDA(cudaMemcpyAsync(&S, t.d_state, sizeof(S), cudaMemcpyDeviceToHost, t.stream));
    PF_CUDA(cudaStreamSynchronize(t.stream));
    memcpy(H21, S.last_H, sizeof(double) * 21);
    memcpy(g6, S.last_g, sizeof(double) * 6);
    *cost = S.last_cost;
    return PF_OK;
}


// Grid-wide evaluation (k_normal_eq_stream): packed host arrays in, sums out; `reps` timed repetitions on the device (the first is a
// warm-up when reps > 1), CUDA events on the tap's stream around the kernel alone.
extern "C" int pf_eval_normal_eq_timed(int device, const double pose[7], const double* edge9, int n_edge, const double* surf7, int n_surf,
                                       int reps, double H21[21], double g6[6], double* cost, float* ms_kernel) {
    PF_REQUIRE(pose && H21 && g6 && cost && n_edge >= 0 && n_surf >= 0 && (edge9 || n_edge == 0) && (surf7 || n_surf == 0) && reps >= 1, "bad argument");
    PF_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    PF_CUDA(cudaGetDeviceProperties(&prop, device));
    PF_REQUIRE(prop.major == 10, "pfilter_b200 needs an sm_100a device, found sm_%d%d", prop.major, prop.minor);
    struct Bufs {
        cudaStream_t stream = nullptr; double *e = nullptr, *s = nullptr, *pose = nullptr, *partial = nullptr, *out = nullptr; unsigned* ticket = nullptr;
        cudaEvent_t ev[2] = {nullptr, nullptr};
        ~Bufs() { cudaFree(e); cudaFree(s); cudaFree(pose); cudaFree(partial); cudaFree(out); cudaFree(ticket);
                  for (auto x : ev) if (x) cudaEventDestroy(x);
                  if (stream) cudaStreamDestroy(stream); }
    } b;
    PF_CUDA(cudaStreamCreateWithFlags(&b.stream, cudaStreamNonBlocking));
    const int grid = prop.multiProcessorCount * kNeCtasPerSm;
    PF_CUDA(cudaMalloc(&b.e, sizeof(double) * 9 * (size_t)(n_edge > 0 ? n_edge : 1)));
    PF_CUDA(cudaMalloc(&b.s, sizeof(double) * 7 * (size_t)(n_surf > 0 ? n_surf : 1)));
    PF_CUDA(cudaMalloc(&b.pose, sizeof(double) * 7));
    PF_CUDA(cudaMalloc(&b.partial, sizeof(double) * 32 * grid));
    PF_CUDA(cudaMalloc(&b.out, sizeof(double) * 32));
    PF_CUDA(cudaMalloc(&b.ticket, sizeof(unsigned)));
    PF_CUDA(cudaMemsetAsync(b.ticket, 0, sizeof(unsigned), b.stream));
    PF_CUDA(cudaMemcpyAsync(b.pose, pose, sizeof(double) * 7, cudaMemcpyHostToDevice, b.stream));
    if (n_edge) PF_CUDA(cudaMemcpyAsync(b.e, edge9, sizeof(double) * 9 * (size_t)n_edge, cudaMemcpyHostToDevice, b.stream));
    if (n_surf) PF_CUDA(cudaMemcpyAsync(b.s, surf7, sizeof(double) * 7 * (size_t)n_surf, cudaMemcpyHostToDevice, b.stream));
    for (auto& x : b.ev) PF_CUDA(cudaEventCreate(&x));
    float sum = 0.f;
    for (int r = 0; r < reps; ++r) {
        PF_CUDA(cudaEventRecord(b.ev[0], b.stream));
        PF_CHECK(normal_eq_stream(b.stream, b.pose, b.e, n_edge, b.s, n_surf, b.partial, b.ticket, b.out, grid));
        PF_CUDA(cudaEventRecord(b.ev[1], b.stream));
        PF_CUDA(cudaStreamSynchronize(b.stream));
        float ms = 0.f;
        PF_CUDA(cudaEventElapsedTime(&ms, b.ev[0], b.ev[1]));
        if (r > 0 || reps == 1) sum += ms;
    }
    double out[32];
    PF_CUDA(cudaMemcpy(out, b.out, sizeof(out), cudaMemcpyDeviceToHost));
    memcpy(H21, out, sizeof(double) * 21);
    memcpy(g6, out + 21, sizeof(double) * 6);
    *cost = out[27];
    if (ms_kernel) *ms_kernel = sum / (reps > 1 ? reps - 1 : 1);
    return PF_OK;
}

extern "C" int pf_lm_solve(int device, double pose_io[7], const double* edge9, int n_edge, const double* surf7, int n_surf, int* iterations,
                           double* final_cost) {
    PF_REQUIRE(pose_io, "null pose");
    SolveTap t;
    LmParams P{};
    PF_CHECK(solve_tap_setup(t, device, pose_io, edge9, n_edge, surf7, n_surf, P));
    P.eval_only = 0;
    { const int ub[2] = {n_edge, n_surf}; PF_CHECK(lm_solve(t.stream, P, nullptr, 1, nullptr, ub)); }
    LmState S;
    PF_CUDA(cudaMemcpyAsync(&S, t.d_state, sizeof(S), cudaMemcpyDeviceToHost, t.stream));
    PF_CUDA(cudaStreamSynchronize(t.stream));
    memcpy(pose_io, S.x, sizeof(double) * 7);
    if (iterations) *iterations = S.iter;
    if (final_cost) *final_cost = S.cost;
    return PF_OK;
}
