// ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into the product library.
//
// C wrapper around the REAL reference extraction class.  It is compiled together with the
// reference's own, unmodified sources where they lie (/root/reference/src/laserProcessingClass.cpp,
// /root/reference/src/lidar.cpp; header /root/reference/include/laserProcessingClass.h:32-41)
// against the PCL stand-in under oracle/shim/.  Nothing from the reference is copied.
// Output goes to oracle/_ref/libpf_ref_extract.so (git-ignored, travels to the GPU box).
//
// The input index of every point is carried through the reference code in the `intensity`
// field (exact for n < 2^24) so that the emitted clouds can be mapped back to input indices.
#include "laserProcessingClass.h"

extern "C" int pfref_extract(const float* xyzi, int n, int num_lines, double min_distance, double max_distance,
                             int* edge_idx, int* n_edge, int* surf_idx, int* n_surf) {
    if (n >= (1 << 24)) return -1;
    lidar::Lidar lp;
    lp.setLines(num_lines);
    lp.setMinDistance(min_distance);
    lp.setMaxDistance(max_distance);
    LaserProcessingClass proc;
    proc.init(lp);
    pcl::PointCloud<pcl::PointXYZI>::Ptr in(new pcl::PointCloud<pcl::PointXYZI>());
    pcl::PointCloud<pcl::PointXYZI>::Ptr edge(new pcl::PointCloud<pcl::PointXYZI>());
    pcl::PointCloud<pcl::PointXYZI>::Ptr surf(new pcl::PointCloud<pcl::PointXYZI>());
    in->points.resize(n);
    for (int i = 0; i < n; ++i) {
        in->points[i].x = xyzi[4 * i + 0];
        in->points[i].y = xyzi[4 * i + 1];
        in->points[i].z = xyzi[4 * i + 2];
        in->points[i].intensity = (float)i;
    }
    proc.featureExtraction(in, edge, surf);
    *n_edge = (int)edge->points.size();
    *n_surf = (int)surf->points.size();
    for (int i = 0; i < *n_edge; ++i) edge_idx[i] = (int)edge->points[i].intensity;
    for (int i = 0; i < *n_surf; ++i) surf_idx[i] = (int)surf->points[i].intensity;
    return 0;
}

// Timing entry: runs the reference extraction `reps` times on the same scan (for the CPU baseline).
extern "C" int pfref_extract_time(const float* xyzi, int n, int num_lines, double min_distance, double max_distance,
                                  int reps, int* n_edge, int* n_surf) {
    lidar::Lidar lp;
    lp.setLines(num_lines);
    lp.setMinDistance(min_distance);
    lp.setMaxDistance(max_distance);
    LaserProcessingClass proc;
    proc.init(lp);
    pcl::PointCloud<pcl::PointXYZI>::Ptr in(new pcl::PointCloud<pcl::PointXYZI>());
    in->points.resize(n);
    for (int i = 0; i < n; ++i) {
        in->points[i].x = xyzi[4 * i + 0];
        in->points[i].y = xyzi[4 * i + 1];
        in->points[i].z = xyzi[4 * i + 2];
        in->points[i].intensity = xyzi[4 * i + 3];
    }
    for (int r = 0; r < reps; ++r) {
        pcl::PointCloud<pcl::PointXYZI>::Ptr edge(new pcl::PointCloud<pcl::PointXYZI>());
        pcl::PointCloud<pcl::PointXYZI>::Ptr surf(new pcl::PointCloud<pcl::PointXYZI>());
        proc.featureExtraction(in, edge, surf);
        *n_edge = (int)edge->points.size();
        *n_surf = (int)surf->points.size();
    }
    return 0;
}
