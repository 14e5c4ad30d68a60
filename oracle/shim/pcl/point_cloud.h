// Minimal stand-in for <pcl/point_cloud.h> (oracle only): points vector, push_back, Ptr.
#pragma once
#include "point_types.h"
namespace pcl {
template <class PointT>
struct PointCloud {
    using Ptr = std::shared_ptr<PointCloud<PointT>>;
    using ConstPtr = std::shared_ptr<const PointCloud<PointT>>;
    std::vector<PointT> points;
    unsigned width = 0, height = 1;
    bool is_dense = true;
    void push_back(const PointT& p) { points.push_back(p); width = (unsigned)points.size(); }
    std::size_t size() const { return points.size(); }
};
}  // namespace pcl
