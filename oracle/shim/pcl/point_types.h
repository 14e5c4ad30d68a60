// Minimal stand-in for <pcl/point_types.h> so that the reference's extraction sources
// (/root/reference/src/laserProcessingClass.cpp, src/lidar.cpp) compile unmodified, in place.
// TEST INFRASTRUCTURE ONLY (oracle). Layout mirrors PCL's 32-byte, 16-byte aligned PointXYZI.
#pragma once
#include <math.h>      // the real PCL/boost header tree reaches <math.h>: unqualified sqrt(float) -> float overload
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <memory>
#include <algorithm>
#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
namespace pcl {
struct alignas(16) PointXYZI {
    float x = 0.f, y = 0.f, z = 0.f, w_ = 1.f;
    float intensity = 0.f, pad_[3] = {0.f, 0.f, 0.f};
};
// 32-byte layout of pcl::PointXYZRGB (xyz + padding, then b g r a packed in one word); used by the typed-adapter compile test only
struct alignas(16) PointXYZRGB {
    float x = 0.f, y = 0.f, z = 0.f, w_ = 1.f;
    unsigned char b = 0, g = 0, r = 0, a = 255;
    float pad_[3] = {0.f, 0.f, 0.f};
};
}  // namespace pcl
