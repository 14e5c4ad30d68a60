// Interface of the voxel-grid kernels (voxel.cu).
#pragma once
#include "primitives.cuh"

namespace pf {

enum { VOX_PCL = 0, VOX_MAP = 1 };

struct VoxCloud {
    const Pt* in;        // device points (16 B); when in_is_xyzi the 4th word is ignored (r=g=b=0, a=255)
    const int* n_in;     // device count (null = empty cloud)
    Pt* out;             // device output, ascending voxel key
    int* n_out;          // device count of the output
    float leaf;
    int in_is_xyzi;
};

struct VoxParams {
    VoxCloud c[2];       // the two clouds (edge, surf) processed by the same launches
    int mode;            // VOX_PCL or VOX_MAP
    const double* center;   // VOX_MAP: device pointer to the crop centre (pose translation), 3 doubles
    int k_new; float theta_p; int theta_max;   // VOX_MAP: PFilter delete rule
    unsigned* state;     // filled by voxelize(): per-step state slot
};

// Enqueues bounds -> keys -> stable radix sort -> ordered segment reduction on ws.stream.
// slot (0 or 1) selects the state slot / scan site, so two voxelisations can be issued in the same pipeline step.
// cap0 / cap1: capacities (upper bounds of *n_in) of the two clouds.
int voxelize(Workspace& ws, const VoxParams& P, int slot, int cap0, int cap1);

}  // namespace pf
