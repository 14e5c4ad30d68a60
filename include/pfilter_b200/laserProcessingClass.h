// LaserProcessingClass with the reference's interface (/root/reference/include/laserProcessingClass.h:32-41),
// implemented by the CUDA extractor behind the C ABI (pf_extract_*).  Header-only.
//
//   LaserProcessingClass p;  p.init(lidar_param);
//   p.featureExtraction(pc_in, pc_out_edge, pc_out_surf);      // appends to the two output clouds, like the reference
//
// Error behaviour follows the reference: nothing is thrown and nothing is returned; a failure is printed to stderr
// and the outputs are left untouched (status() returns the last pf_status for callers that want it).
#pragma once
#include <cstdio>

#include "cloud.h"
#include "lidar.h"

class LaserProcessingClass {
   public:
    using Cloud = pfilter_b200::PointCloud<pfilter_b200::PointXYZI>;

    LaserProcessingClass() {}
    ~LaserProcessingClass() { if (h_) pf_extract_destroy(h_); }
    LaserProcessingClass(const LaserProcessingClass&) = delete;
    LaserProcessingClass& operator=(const LaserProcessingClass&) = delete;

    // device / max_points are extensions with defaults; the reference signature is init(lidar::Lidar)
    void init(lidar::Lidar lidar_param_in, int device = 0, int max_points = 262144) {
        lidar_param = lidar_param_in;
        if (h_) { pf_extract_destroy(h_); h_ = nullptr; }
        pf_lidar_params lp = lidar_param.c_params();
        pf_extract_config cfg{max_points, 1, 0};
        status_ = pf_extract_create(&lp, &cfg, device, &h_);
        if (status_ != PF_OK) std::fprintf(stderr, "LaserProcessingClass::init: %s\n", pf_last_error());
        max_points_ = max_points;
    }

    void featureExtraction(const Cloud::Ptr& pc_in, Cloud::Ptr& pc_out_edge, Cloud::Ptr& pc_out_surf) {
        if (!h_) { std::fprintf(stderr, "LaserProcessingClass: init() has not succeeded\n"); return; }
        const int n = (int)pc_in->points.size();
        edge_.resize((size_t)120 * lidar_param.num_lines);
        surf_.resize(n > 0 ? n : 1);
        int ne = 0, ns = 0;
        status_ = pf_extract_run(h_, reinterpret_cast<const float*>(pc_in->points.data()), n, reinterpret_cast<float*>(edge_.data()), &ne,
                                 reinterpret_cast<float*>(surf_.data()), &ns, nullptr);
        if (status_ != PF_OK) { std::fprintf(stderr, "LaserProcessingClass::featureExtraction: %s\n", pf_last_error()); return; }
        pc_out_edge->points.insert(pc_out_edge->points.end(), edge_.begin(), edge_.begin() + ne);
        pc_out_surf->points.insert(pc_out_surf->points.end(), surf_.begin(), surf_.begin() + ns);
    }

    int status() const { return status_; }
    pf_extract* handle() { return h_; }

   private:
    lidar::Lidar lidar_param;
    pf_extract* h_ = nullptr;
    int status_ = PF_OK, max_points_ = 0;
    std::vector<pfilter_b200::PointXYZI> edge_, surf_;
};
