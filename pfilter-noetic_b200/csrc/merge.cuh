// K9 (streaming form): map maintenance of addPointsToMap (/root/reference/src/odomEstimationClass.cpp:606-647) as a MERGE.
//
// The reference re-voxelises the whole local map every frame: CropBox, rgbds (sort all M + n points by voxel id, :34-134),
// extractstablepoint, r += 2.  But rgbds emits its voxels in ascending id = lexicographic (kz, ky, kx) order, the grid is
// globally anchored (floor(x / leaf)), and a centroid stays inside the voxel it was averaged in, so the map that comes out
// of one update is already sorted for the next one.  Only the points appended by the current frame (and the rare centroids
// that float rounding pushed across a voxel face, kept as "exceptions" behind the sorted part) are unsorted.
//
// Map buffer layout:  [0, n_sorted) strictly ascending voxel key, one point per voxel
//                     [n_sorted, n_map) exceptions of the last update
//                     [n_map, n_app)    points appended by this frame (k_append)
// One update = sort B = [n_sorted, n_app) (small), find B's voxel runs ("heads"), and ONE streaming pass over the sorted
// part that crops, merges the matching runs, inserts the new voxels in key order, applies the PFilter delete rule and the
// r update and writes the next map: 16 B read per map point, 16 B written per surviving voxel.
// With n_sorted = 0 (first update after initMapWithPoints: the raw first-frame clouds) everything is B and the pipeline
// degenerates to the full sort.  Summation order inside a voxel: map point first, then B in buffer order = the reference's
// ascending index order, so centroids are bit-identical.
#pragma once
#include "primitives.cuh"

namespace pf {

constexpr int kMergeMaxGrid = 16 * kSMs;   // CTAs per cloud of the two streaming passes

struct MapMergeCloud {
    const Pt* buf;           // current map buffer (layout above)
    const int* n_sorted;     // device counts
    const int* n_app;
    Pt* out;                 // next map buffer
    int* n_out;              // next n_map (sorted + exceptions)
    int* n_sorted_out;       // next n_sorted
    float leaf;
};

struct MapMergeScratch {     // sized by the capacity of B (both clouds together)
    int* m_ra;    Pt* m_pt;      int* m_delta;    // matched voxels: map index, finished point (alpha 0 = not kept), count change
    int* i_ra;    Pt* i_pt;                       // inserted voxels (new, kept): insertion index, finished point
    Pt* exc;      int exc_cap;                    // [2][exc_cap] exceptions of this update
    int cap;                                      // entries per list and cloud half: cloud c uses [c * cap, (c + 1) * cap)
    int* tile_m;  int* tile_i;   int tile_cap;    // [2][tile_cap] first matched head / insert of every 1024-point map tile
    int* tile_agg; int* cta_sum;                  // [2][tile_cap] points a tile emits; [2][kMergeMaxGrid] per CTA of the count pass
};

struct MapMergeParams {
    MapMergeCloud c[2];
    const double* center;    // device: crop centre (pose translation), 3 doubles
    int k_new; float theta_p; int theta_max;
    MapMergeScratch s;
    unsigned* state;         // filled by map_merge(): state slot
    int hint;                // filled by map_merge(): L2 eviction-priority experiment switch (PF_MM_HINT)
};

// capB0 / capB1: upper bounds of (n_app - n_sorted); capA0 / capA1: upper bounds of n_sorted.
int map_merge(Workspace& ws, const MapMergeParams& P, int capB0, int capB1, int capA0, int capA1);
// device word holding the error bits of the last map_merge of `ws` (2 = voxel coordinates out of range, 4 = exceptions overflowed)
inline const unsigned* map_merge_error_word(const Workspace& ws) { return ws.ctrl + kSlotBase + 3 * kSlotWords + 15; }
// cap_b: capacity of the unsorted part per cloud; cap_a: capacity of the sorted part (map buffer) per cloud
int map_merge_scratch_create(MapMergeScratch& s, int cap_b, int exc_cap, int cap_a);
void map_merge_scratch_destroy(MapMergeScratch& s);

}  // namespace pf
