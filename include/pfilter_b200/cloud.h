// Minimal point / cloud types for the ROS-free harness.  They mirror the parts of pcl::PointXYZI,
// pcl::PointXYZRGB and pcl::PointCloud<T> that the reference's callers touch (points, push_back, size, Ptr),
// with the 16-byte layouts of the C ABI so that no repacking is needed.  A PCL build can convert with two
// loops.
#pragma once
#include <cstdint>
#include <memory>
#include <vector>

namespace pfilter_b200 {

struct PointXYZI { float x = 0, y = 0, z = 0, intensity = 0; };
struct PointXYZRGB { float x = 0, y = 0, z = 0; uint8_t r = 0, g = 0, b = 0, a = 255; };
static_assert(sizeof(PointXYZI) == 16 && sizeof(PointXYZRGB) == 16, "ABI layout");

template <class PointT>
struct PointCloud {
    using Ptr = std::shared_ptr<PointCloud<PointT>>;
    using ConstPtr = std::shared_ptr<const PointCloud<PointT>>;
    std::vector<PointT> points;
    void push_back(const PointT& p) { points.push_back(p); }
    std::size_t size() const { return points.size(); }
    void clear() { points.clear(); }
    PointCloud& operator+=(const PointCloud& o) { points.insert(points.end(), o.points.begin(), o.points.end()); return *this; }
};

}  // namespace pfilter_b200
