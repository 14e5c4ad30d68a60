// K4-K6: scan-to-map association (addEdgeCostFactor / addSurfCostFactor,
// /root/reference/src/odomEstimationClass.cpp:284-432, :434-578) for both feature kinds in the same launches.
#pragma once
#include "knn.cuh"

namespace pf {

struct AssocCloud {
    Pt* queries;           // down-sampled feature cloud (sensor frame); r, g written for accepted queries (:354-355)
    const int* n_q;        // device count
    Pt* map;               // local map; g counters updated with the reference's sequential semantics (:345-346)
    const int* n_map;
    KnnGrid grid;
    int* head;             // [map capacity] per-map-point list of hits, -1 = empty (self-cleaning)
    int* hits;             // [map capacity] pending readers of the list, 0 (self-cleaning)
    int* next;             // [5 * query capacity]
    int* nn_idx;           // [5 * query capacity]
    uint8_t* flag;         // [query capacity] 0 none, 1 fit ok but skipped by the persistence rule, 2 residual block
    double* geom;          // [8 * query capacity] edge: a[3] b[3]; surf: n[3] d
    float* w_obs;          // [query capacity] weightType != 0: the residual's observe value (:360 / :509)
    double* w_spa;         // [query capacity] weightType != 0: point sparsity (:367-385 / :513-531)
};

struct AssocParams {
    AssocCloud c[2];       // 0 = edge (line fit), 1 = surf (plane fit)
    const double* pose;    // device [qx qy qz qw tx ty tz]
    int k_new; float theta_p; int theta_max;
    const int* guard;      // device flag: the maps hold enough points to associate (:247); null = always (stage tap)
    int weight_type;       // 0, 1, 2 or 12 (src/odomEstimationClass.cpp:389-423)
    unsigned long long* w_minmax;   // [2][4] per kind: min / max of observe, min / max of sparsity over the pass's residual blocks
                                    // (bit patterns of non-negative doubles: ordered as unsigned integers)
};

// Enqueues the two association kernels (match, persist) for one pass on `stream`.
int associate_pass(cudaStream_t stream, const AssocParams& P, int qcap0, int qcap1, uint64_t* launches);

}  // namespace pf
