// Stand-in for <pcl/filters/filter.h> (oracle only). The index-only overload of
// removeNaNFromPointCloud used at /root/reference/src/laserProcessingClass.cpp:13 only fills `index`.
#pragma once
#include "../point_cloud.h"
namespace pcl {
template <class PointT>
void removeNaNFromPointCloud(const PointCloud<PointT>& cloud, std::vector<int>& index) {
    index.clear();
    index.reserve(cloud.points.size());
    for (std::size_t i = 0; i < cloud.points.size(); ++i)
        if (std::isfinite(cloud.points[i].x) && std::isfinite(cloud.points[i].y) && std::isfinite(cloud.points[i].z))
            index.push_back((int)i);
}
}  // namespace pcl
