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
// pcl::copyPointCloud between point types: copies the fields both types have (x, y, z here), the rest keep their defaults
template <class A, class B>
void copyPointCloud(const PointCloud<A>& in, PointCloud<B>& out) {
    out.points.resize(in.points.size());
    for (std::size_t i = 0; i < in.points.size(); ++i) { B q; q.x = in.points[i].x; q.y = in.points[i].y; q.z = in.points[i].z; out.points[i] = q; }
    out.width = (unsigned)out.points.size();
}
}  // namespace pcl
