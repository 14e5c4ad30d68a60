// LaserMappingClass with the reference's interface (/root/reference/include/laserMappingClass.h:32-58), implemented by the
// CUDA global map behind the C ABI (pf_mapping_*).  Header-only.
//
//   LaserMappingClass m;  m.init(map_resolution);
//   m.updateCurrentPointsToMap(pc_in, pose);      // pose: row-major 3x4 [R | t] doubles (the reference passes an
//   auto cloud = m.getMap();                      //       Eigen::Isometry3d; isometry.matrix().topRows<3>() row-major)
//
// The map lives in HBM; getMap() copies it out on demand (the reference concatenates every cell on every call,
// src/laserMappingClass.cpp:196-208).  Errors are printed and the call carries on, like the rest of the reference.
#pragma once
#include <cstdio>

#include "../pfilter_b200.h"
#include "cloud.h"

class LaserMappingClass {
   public:
    using Cloud = pfilter_b200::PointCloud<pfilter_b200::PointXYZI>;

    LaserMappingClass() {}
    ~LaserMappingClass() { if (h_) pf_mapping_destroy(h_); }
    LaserMappingClass(const LaserMappingClass&) = delete;
    LaserMappingClass& operator=(const LaserMappingClass&) = delete;

    // device / capacities are extensions with defaults; the reference signature is init(double map_resolution)
    void init(double map_resolution, int device = 0, int max_map_points = 0, int max_points = 0) {
        if (h_) { pf_mapping_destroy(h_); h_ = nullptr; }
        status_ = pf_mapping_create(map_resolution, max_map_points, max_points, device, &h_);
        if (status_ != PF_OK) std::fprintf(stderr, "LaserMappingClass::init: %s\n", pf_last_error());
    }

    void updateCurrentPointsToMap(const Cloud::Ptr& pc_in, const double pose_current_rt[12]) {
        if (!h_) return;
        status_ = pf_mapping_update(h_, reinterpret_cast<const float*>(pc_in->points.data()), (int)pc_in->points.size(), pose_current_rt);
        if (status_ != PF_OK) std::fprintf(stderr, "updateCurrentPointsToMap: %s\n", pf_last_error());
    }

    Cloud::Ptr getMap() {
        Cloud::Ptr c(new Cloud());
        if (!h_) return c;
        int n = 0;
        status_ = pf_mapping_map_size(h_, &n);
        if (status_ != PF_OK || n <= 0) return c;
        c->points.resize(n);
        status_ = pf_mapping_get_map(h_, reinterpret_cast<float*>(c->points.data()), n, &n);
        c->points.resize(status_ == PF_OK ? n : 0);
        return c;
    }

    int status() const { return status_; }
    pf_mapping* handle() { return h_; }

   private:
    pf_mapping* h_ = nullptr;
    int status_ = PF_OK;
};
