// ROS-free harness: the call order of the reference's three node loops (src/laserProcessingNode.cpp:43-106,
// src/odomEstimationNode copy.cpp:54-130, src/laserMappingNode.cpp:60-95) with the reference's class names, driven by the
// synthetic LiDAR generator instead of ROS topics.  The classes come from include/pfilter_b200/*.h and run on the GPU through
// libpfilter_b200.so -- swap the include directory and this is the reference's own code path.
//
//   g++ -std=c++17 -O2 -I include examples/harness.cpp -o examples/harness \
//       pfilter-noetic_b200/libpfilter_b200.so pfilter-noetic_b200/libpf_synth.so -Wl,-rpath,'$ORIGIN/../pfilter-noetic_b200'
//   examples/harness [frames]
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <memory>

#include "pf_synth.h"
#include "pfilter_b200/laserMappingClass.h"
#include "pfilter_b200/laserProcessingClass.h"
#include "pfilter_b200/odomEstimationClass.h"

int main(int argc, char** argv) {
    const int frames = argc > 1 ? std::atoi(argv[1]) : 20;
    pf_synth_params sp;
    pf_synth_default_params(&sp);
    sp.sensor_lines = 64;
    sp.seed = 2022;

    lidar::Lidar lidar_param;                 // src/laserProcessingNode.cpp:120-134
    lidar_param.setScanPeriod(0.1);
    lidar_param.setLines(64);
    lidar_param.setMaxDistance(90.0);
    lidar_param.setMinDistance(3.0);

    LaserProcessingClass laserProcessing;
    laserProcessing.init(lidar_param);
    OdomEstimationClass odomEstimation;       // src/odomEstimationNode copy.cpp:176-188: map_resolution 0.4, PFilter 0 / 0.4 / 75
    odomEstimation.init(lidar_param, 0.4, 0, 0.4f, 75, 0.0);
    LaserMappingClass laserMapping;           // src/laserMappingNode.cpp:122
    laserMapping.init(0.4);
    if (laserProcessing.status() != PF_OK || odomEstimation.status() != PF_OK || laserMapping.status() != PF_OK) return 2;

    using CloudI = LaserProcessingClass::Cloud;
    using CloudRGB = OdomEstimationClass::Cloud;
    bool is_odom_inited = false;
    double worst = 0, total_ms = 0;
    for (int f = 0; f < frames; ++f) {
        auto pointcloud_in = std::make_shared<CloudI>();
        pointcloud_in->points.resize((size_t)sp.sensor_lines * sp.azimuth_steps);
        const int n = pf_synth_scan(&sp, f, reinterpret_cast<float*>(pointcloud_in->points.data()), (int)pointcloud_in->points.size());
        if (n < 0) return 3;
        pointcloud_in->points.resize(n);
        const auto t0 = std::chrono::steady_clock::now();

        auto pointcloud_edge = std::make_shared<CloudI>(), pointcloud_surf = std::make_shared<CloudI>();
        laserProcessing.featureExtraction(pointcloud_in, pointcloud_edge, pointcloud_surf);      // src/laserProcessingNode.cpp:73

        auto edge = std::make_shared<CloudRGB>(), surf = std::make_shared<CloudRGB>();           // copyPointCloud XYZI -> XYZRGB (:74-80)
        for (const auto& p : pointcloud_edge->points) edge->points.push_back({p.x, p.y, p.z, 0, 0, 0, 255});
        for (const auto& p : pointcloud_surf->points) surf->points.push_back({p.x, p.y, p.z, 0, 0, 0, 255});
        if (!is_odom_inited) { odomEstimation.initMapWithPoints(edge, surf); is_odom_inited = true; }   // :87-95
        else odomEstimation.updatePointsToMap(edge, surf);
        if (odomEstimation.status() != PF_OK) return 4;

        double R[9], rt[12];
        odomEstimation.odom.rotation_matrix(R);
        const double* t = odomEstimation.odom.translation();
        for (int r = 0; r < 3; ++r) { rt[4 * r] = R[3 * r]; rt[4 * r + 1] = R[3 * r + 1]; rt[4 * r + 2] = R[3 * r + 2]; rt[4 * r + 3] = t[r]; }
        laserMapping.updateCurrentPointsToMap(pointcloud_surf, rt);                               // src/laserMappingNode.cpp:84
        if (laserMapping.status() != PF_OK) return 5;
        total_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();

        double gt[7], gt0[7];
        pf_synth_pose(&sp, f, gt);
        pf_synth_pose(&sp, 0, gt0);
        const double dx = t[0] - (gt[4] - gt0[4]), dy = t[1] - (gt[5] - gt0[5]), dz = t[2] - (gt[6] - gt0[6]);
        worst = std::fmax(worst, std::sqrt(dx * dx + dy * dy + dz * dz));
    }
    auto local_map = std::make_shared<CloudRGB>();
    odomEstimation.getMap(local_map);
    auto global_map = laserMapping.getMap();
    std::printf("harness: %d frames, %.3f ms/frame, final t = (%.3f %.3f %.3f), max |t - ground truth| = %.3f m, local map %zu pts, global map %zu pts\n",
                frames, total_ms / frames, odomEstimation.odom.t[0], odomEstimation.odom.t[1], odomEstimation.odom.t[2], worst,
                local_map->points.size(), global_map->points.size());
    return worst < 0.3 && !local_map->points.empty() && !global_map->points.empty() ? 0 : 1;
}
