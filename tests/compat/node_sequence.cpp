// TEST: the call sequences of the reference's three node loops, typed exactly as the reference types them (pcl::PointCloud<...>::Ptr,
// Eigen::Isometry3d, public members `odom`, `laserCloudCornerMap`, `laserCloudSurfMap`), compiled against
// include/pfilter_b200/compat_eigen_pcl.h.  Eigen and PCL come from the minimal stand-ins under oracle/shim (neither library exists in
// this image); with the real headers on the include path the same source compiles against them.
//   extraction  /root/reference/src/laserProcessingNode.cpp:62-73
//   odometry    /root/reference/src/odomEstimationNode copy.cpp:74-106, src/odomEstimationNode.cpp:144-184
//   mapping     /root/reference/src/laserMappingNode.cpp:75-87
// Prints one line per frame: pose, map sizes; exit code 0 when every call succeeded.
#include <Eigen/Dense>
#include <Eigen/Geometry>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>

#include <cstdio>
#include <cstdlib>

#include "pf_synth.h"
#include "pfilter_b200/compat_eigen_pcl.h"

using namespace pfilter_b200::compat;

LaserProcessingClass laserProcessing;
Odom_ES_EstimationClass odom_ES_Estimation;
LaserMappingClass laserMapping;
lidar::Lidar lidar_param;

int main(int argc, char** argv) {
    const int frames = argc > 1 ? std::atoi(argv[1]) : 5;
    pf_synth_params sp;
    pf_synth_default_params(&sp);

    lidar_param.setScanPeriod(0.1);
    lidar_param.setVerticalAngle(2.0);
    lidar_param.setLines(64);
    lidar_param.setMaxDistance(90.0);
    lidar_param.setMinDistance(3.0);
    laserProcessing.init(lidar_param);
    odom_ES_Estimation.init(lidar_param, 0.4, 0, 0.4f, 75, 0.0);
    laserMapping.init(0.4);
    if (laserProcessing.status() != PF_OK || odom_ES_Estimation.status() != PF_OK || laserMapping.status() != PF_OK) return 2;

    bool is_odom_inited = false;
    for (int f = 0; f < frames; ++f) {
        pcl::PointCloud<pcl::PointXYZI>::Ptr pointcloud_in(new pcl::PointCloud<pcl::PointXYZI>());
        {   // stands in for pcl::fromROSMsg
            std::vector<float> raw((size_t)sp.sensor_lines * sp.azimuth_steps * 4);
            const int n = pf_synth_scan(&sp, f, raw.data(), sp.sensor_lines * sp.azimuth_steps);
            if (n < 0) return 3;
            for (int i = 0; i < n; ++i) {
                pcl::PointXYZI p;
                p.x = raw[4 * i]; p.y = raw[4 * i + 1]; p.z = raw[4 * i + 2]; p.intensity = raw[4 * i + 3];
                pointcloud_in->push_back(p);
            }
        }
        // ---- laserProcessingNode
        pcl::PointCloud<pcl::PointXYZI>::Ptr pointcloud_edge(new pcl::PointCloud<pcl::PointXYZI>());
        pcl::PointCloud<pcl::PointXYZI>::Ptr pointcloud_surf(new pcl::PointCloud<pcl::PointXYZI>());
        laserProcessing.featureExtraction(pointcloud_in, pointcloud_edge, pointcloud_surf);

        // ---- odomEstimationNode
        pcl::PointCloud<pcl::PointXYZRGB>::Ptr pointcloud_edge_in(new pcl::PointCloud<pcl::PointXYZRGB>());
        pcl::PointCloud<pcl::PointXYZRGB>::Ptr pointcloud_surf_in(new pcl::PointCloud<pcl::PointXYZRGB>());
        pcl::copyPointCloud(*pointcloud_edge, *pointcloud_edge_in);
        pcl::copyPointCloud(*pointcloud_surf, *pointcloud_surf_in);
        if (is_odom_inited == false) {
            odom_ES_Estimation.initMapWithPoints(pointcloud_edge_in, pointcloud_surf_in);
            is_odom_inited = true;
        } else {
            odom_ES_Estimation.updatePointsToMap(pointcloud_edge_in, pointcloud_surf_in);
        }
        if (odom_ES_Estimation.status() != PF_OK) return 4;
        Eigen::Quaterniond q_current(odom_ES_Estimation.odom.rotation());
        Eigen::Vector3d t_current = odom_ES_Estimation.odom.translation();
        const size_t n_corner = (*odom_ES_Estimation.laserCloudCornerMap).points.size();     // pcl::toROSMsg(*...laserCloudCornerMap, cloudMsg)
        const size_t n_surf = (*odom_ES_Estimation.laserCloudSurfMap).points.size();
        pcl::PointCloud<PointType>::Ptr whole(new pcl::PointCloud<PointType>());
        odom_ES_Estimation.getMap(whole);

        // ---- laserMappingNode
        Eigen::Isometry3d current_pose = Eigen::Isometry3d::Identity();
        current_pose.rotate(Eigen::Quaterniond(q_current.w(), q_current.x(), q_current.y(), q_current.z()));
        current_pose.pretranslate(Eigen::Vector3d(t_current.x(), t_current.y(), t_current.z()));
        laserMapping.updateCurrentPointsToMap(pointcloud_in, current_pose);
        pcl::PointCloud<pcl::PointXYZI>::Ptr pc_map = laserMapping.getMap();
        if (laserMapping.status() != PF_OK) return 5;

        std::printf("frame %d pose %.9f %.9f %.9f %.9f %.9f %.9f %.9f edge %zu surf %zu corner_map %zu surf_map %zu whole %zu global %zu\n", f, q_current.x(),
                    q_current.y(), q_current.z(), q_current.w(), t_current.x(), t_current.y(), t_current.z(), pointcloud_edge->points.size(),
                    pointcloud_surf->points.size(), n_corner, n_surf, whole->points.size(), pc_map->points.size());
        if (whole->points.size() != n_corner + n_surf) return 6;
    }
    return 0;
}
