// K7 / K8: fused residual + Jacobian + Huber + 6x6 normal-equation reduction and the on-device Levenberg-Marquardt
// state machine that replaces ceres::Solve (/root/reference/src/odomEstimationClass.cpp:254-271,
// src/lidarOptimization.cpp:12-104; Ceres behaviour restated in SURVEY.md appendix A.3).
#pragma once
#include "common.cuh"

namespace pf {

constexpr int kLmMaxSrc = 3;  // ES path: edge, surf; BPF path: beam, pillar, facade

struct ResidualSrc {          // one feature kind
    int type;                 // 0 = point-to-line residual (edge / beam / pillar), 1 = point-to-plane (surf / facade)
    const Pt* queries;        // sensor-frame points (float), used when p_override == null
    const double* p_override; // [3 n] points in double (stage taps)
    const uint8_t* flag;      // 2 = residual block present
    const double* geom;       // [8 n] edge: a[3] b[3]; surf: n[3] d
    const int* n;             // device count
    const float* w_obs;       // weightType != 0: observe value of the residual block
    const double* w_spa;      // weightType != 0: point sparsity of the residual block
    const unsigned long long* w_minmax;   // weightType != 0: [4] min / max of observe, min / max of sparsity of this kind (match.cuh)
};

struct LmState {
    double x[7];              // accepted pose [qx qy qz qw tx ty tz]
    double xc[7];             // candidate pose under evaluation
    double cost;              // 1/2 sum rho at x
    double H[21], g[6];       // sum J^T J (upper, row-major) and sum J^T r at x (robustified)
    double scale[6];          // Jacobi scaling fixed at iteration 0
    double diag[6];           // LM diagonal (scaled space)
    double step[6];           // last step (scaled space)
    double radius, decrease, x_norm, model_cost_change;
    double last_H[21], last_g[6], last_cost;   // sums of the most recent evaluation (stage taps read these)
    int phase;                // 0 initial evaluation pending, 1 candidate evaluation pending, 2 finished
    int iter;                 // LM step attempts so far (<= 4)
    int reuse_diag;
    int n_res;                // residual blocks in the last evaluation
    int pass;                 // outer iteration index within the frame
    int n_edge_res, n_surf_res;
    int pad;
};

struct LmParams {
    ResidualSrc src[kLmMaxSrc];   // in the order the reference adds the residual blocks
    int nsrc;
    LmState* state;
    double* iter_poses;       // [16][7] pose after every outer iteration (may be null)
    int eval_only;            // stage tap: evaluate at state->x and stop
    int weight_type;          // 0, 1, 2, 12: residual weights (src/odomEstimationClass.cpp:389-423, src/lidarOptimization.cpp:25-28, :62-63)
};

#ifndef PF_LM_CLUSTER
#define PF_LM_CLUSTER 12
#endif
#ifndef PF_LM_THREADS
#define PF_LM_THREADS 256
#endif
constexpr int kLmCluster = PF_LM_CLUSTER;    // CTAs of the solver cluster (distributed shared memory reduction); 12 needs the non-portable cluster attribute, measured: 4: 2870, 8: 3125, 12: 3223, 16: 3215 scans/s
constexpr int kLmThreads = PF_LM_THREADS;  // 255 registers per thread keep the serial state machine out of local memory

// One launch = one complete solve: at most 1 + 4 evaluations (max_num_iterations = 4) and the trust-region state machine.
// pose_src: device pose to start from (null: keep state->x); first_pass resets the outer-iteration counter.
// ub[s]: upper bound of the query count of source s (sizes the shared-memory list of residual blocks).
int lm_solve(cudaStream_t stream, const LmParams& P, const double* pose_src, int first_pass, uint64_t* launches, const int* ub);   // ub[nsrc]

}  // namespace pf
