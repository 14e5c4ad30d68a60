/* Synthetic LiDAR sequence generator (test / bench data only). */
#ifndef PF_SYNTH_H_
#define PF_SYNTH_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { PF_SYNTH_SCENE_STREET = 0, PF_SYNTH_SCENE_CAMPUS = 1, PF_SYNTH_SCENE_CAMPUS_DENSE = 2, PF_SYNTH_SCENE_STREET_DENSE = 3 };
enum { PF_SYNTH_TRAJ_STREET = 0, PF_SYNTH_TRAJ_LOOP = 1 };

typedef struct pf_synth_params {
    int32_t sensor_lines;    /* 64 (HDL-64E shaped), 32 (VLP-32 shaped) or 16 */
    int32_t azimuth_steps;   /* 1800 = 0.2 deg */
    uint64_t seed;
    int32_t scene;           /* PF_SYNTH_SCENE_* */
    int32_t trajectory;      /* PF_SYNTH_TRAJ_* */
    double speed;            /* metres per frame */
    double range_sigma;      /* gaussian range noise along the ray, metres */
    double elev_jitter_deg;  /* uniform +- jitter of the ring elevation, degrees */
    double min_range;        /* horizontal range gate of the generator (inside the 3-90 m gate of the extractor) */
    double max_range;
} pf_synth_params;

void pf_synth_default_params(pf_synth_params* p);
/* Ground-truth sensor pose of a frame as [qx qy qz qw tx ty tz]. */
void pf_synth_pose(const pf_synth_params* p, int frame, double pose[7]);
/* Writes up to cap points (x,y,z,intensity float32) in the sensor frame; returns the count or -1 on overflow. */
int pf_synth_scan(const pf_synth_params* p, int frame, float* out_xyzi, int cap);

#ifdef __cplusplus
}
#endif
#endif
