// Typed drop-in surface for callers written against the reference's headers: the same class names, the same member and method
// signatures, Eigen and PCL types included, so that the node sources compile UNCHANGED against this header
//   LaserProcessingClass      /root/reference/include/laserProcessingClass.h:32-41
//   Odom_ES_EstimationClass   /root/reference/include/odomEstimationClass.h:140-167  (public `Eigen::Isometry3d odom`, :57; public
//                             `laserCloudCornerMap` / `laserCloudSurfMap`, :151-152, read by src/odomEstimationNode.cpp:144-146, :171, :179)
//   LaserMappingClass         /root/reference/include/laserMappingClass.h:32-58  (updateCurrentPointsToMap takes an Eigen::Isometry3d)
// It is compiled only where Eigen and PCL exist: include <Eigen/Geometry>, <pcl/point_cloud.h> and <pcl/point_types.h> BEFORE this
// header (the reference's node sources already do).  The classes live in namespace pfilter_b200::compat; a translation unit that
// wants them under the reference's global names adds `using namespace pfilter_b200::compat;`.
//
// pcl::PointXYZI / pcl::PointXYZRGB are 32-byte structs; the C ABI moves 16-byte points.  The conversion is two loops per call
// (the host side of a frame is ~0.1 ms for 115 k points); nothing else is copied.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../pfilter_b200.h"
#include "lidar.h"

namespace pfilter_b200 {
namespace compat {

typedef pcl::PointXYZRGB PointType;      // include/odomEstimationClass.h:38

namespace detail {
struct P16 { float x, y, z, w; };
template <class PclPoint>
inline void pack_xyz(const pcl::PointCloud<PclPoint>& in, std::vector<P16>& out, bool intensity) {
    out.resize(in.points.size());
    for (std::size_t i = 0; i < in.points.size(); ++i) {
        out[i].x = in.points[i].x; out[i].y = in.points[i].y; out[i].z = in.points[i].z; out[i].w = 0.f;
    }
    (void)intensity;
}
inline void pack_xyzi(const pcl::PointCloud<pcl::PointXYZI>& in, std::vector<P16>& out) {
    out.resize(in.points.size());
    for (std::size_t i = 0; i < in.points.size(); ++i) {
        out[i].x = in.points[i].x; out[i].y = in.points[i].y; out[i].z = in.points[i].z; out[i].w = in.points[i].intensity;
    }
}
inline void append_xyzi(const P16* p, int n, pcl::PointCloud<pcl::PointXYZI>& out) {
    for (int i = 0; i < n; ++i) {
        pcl::PointXYZI q;
        q.x = p[i].x; q.y = p[i].y; q.z = p[i].z; q.intensity = p[i].w;
        out.push_back(q);
    }
}
inline void append_rgb(const pf_point* p, int n, pcl::PointCloud<PointType>& out) {
    for (int i = 0; i < n; ++i) {
        PointType q;
        q.x = p[i].x; q.y = p[i].y; q.z = p[i].z; q.r = p[i].r; q.g = p[i].g; q.b = p[i].b; q.a = p[i].a;
        out.push_back(q);
    }
}
}  // namespace detail

class LaserProcessingClass {
   public:
    LaserProcessingClass() {}
    ~LaserProcessingClass() { if (h_) pf_extract_destroy(h_); }
    LaserProcessingClass(const LaserProcessingClass&) = delete;
    LaserProcessingClass& operator=(const LaserProcessingClass&) = delete;

    void init(lidar::Lidar lidar_param_in) {                                                   // src/laserProcessingClass.cpp:4-8
        lidar_param = lidar_param_in;
        if (h_) { pf_extract_destroy(h_); h_ = nullptr; }
        pf_lidar_params lp = lidar_param.c_params();
        pf_extract_config cfg{262144, 1, 0, 1};     // surf_order 1: the reference's emission order (ascending curvature in a sector)
        status_ = pf_extract_create(&lp, &cfg, 0, &h_);
        if (status_ != PF_OK) std::fprintf(stderr, "LaserProcessingClass::init: %s\n", pf_last_error());
    }
    void featureExtraction(const pcl::PointCloud<pcl::PointXYZI>::Ptr& pc_in, pcl::PointCloud<pcl::PointXYZI>::Ptr& pc_out_edge,
                           pcl::PointCloud<pcl::PointXYZI>::Ptr& pc_out_surf) {                // :10-96: appends, input untouched
        if (!h_) { std::fprintf(stderr, "LaserProcessingClass: init() has not succeeded\n"); return; }
        detail::pack_xyzi(*pc_in, in_);
        edge_.resize((std::size_t)120 * lidar_param.num_lines);
        surf_.resize(in_.size() ? in_.size() : 1);
        int ne = 0, ns = 0;
        status_ = pf_extract_run(h_, reinterpret_cast<const float*>(in_.data()), (int)in_.size(), reinterpret_cast<float*>(edge_.data()), &ne,
                                 reinterpret_cast<float*>(surf_.data()), &ns, nullptr);
        if (status_ != PF_OK) { std::fprintf(stderr, "LaserProcessingClass::featureExtraction: %s\n", pf_last_error()); return; }
        detail::append_xyzi(edge_.data(), ne, *pc_out_edge);
        detail::append_xyzi(surf_.data(), ns, *pc_out_surf);
    }
    int status() const { return status_; }

   private:
    lidar::Lidar lidar_param;
    pf_extract* h_ = nullptr;
    int status_ = PF_OK;
    std::vector<detail::P16> in_, edge_, surf_;
};

class Odom_ES_EstimationClass;

// Stands where the reference has `pcl::PointCloud<PointType>::Ptr laserCloudCornerMap`: dereferencing it (`*m`, `m->points`)
// fetches the map from HBM if an update has happened since the last fetch -- a node that never publishes the maps never pays for the
// copy (the reference rebuilds the cloud on the host every frame either way).
class LazyMapPtr {
   public:
    pcl::PointCloud<PointType>& operator*() { refresh(); return *cloud_; }
    pcl::PointCloud<PointType>* operator->() { refresh(); return cloud_.get(); }
    operator typename pcl::PointCloud<PointType>::Ptr() { refresh(); return cloud_; }
    typename pcl::PointCloud<PointType>::Ptr get() { refresh(); return cloud_; }

   private:
    friend class Odom_ES_EstimationClass;
    void refresh();
    Odom_ES_EstimationClass* owner_ = nullptr;
    int which_ = 0;
    long long fetched_version_ = -1;
    typename pcl::PointCloud<PointType>::Ptr cloud_{new pcl::PointCloud<PointType>()};
};

class Odom_ES_EstimationClass {
   public:
    Odom_ES_EstimationClass() {
        odom = Eigen::Isometry3d::Identity();
        laserCloudCornerMap.owner_ = this; laserCloudCornerMap.which_ = 0;
        laserCloudSurfMap.owner_ = this; laserCloudSurfMap.which_ = 1;
    }
    ~Odom_ES_EstimationClass() { if (h_) pf_odom_destroy(h_); }
    Odom_ES_EstimationClass(const Odom_ES_EstimationClass&) = delete;
    Odom_ES_EstimationClass& operator=(const Odom_ES_EstimationClass&) = delete;

    // src/odomEstimationClass.cpp:182-208
    void init(lidar::Lidar /*lidar_param*/, double map_resolution_in, int k_new_para, float theta_p_para, int theta_max_para, double weightType_para) {
        if (h_) { pf_odom_destroy(h_); h_ = nullptr; }
        pf_odom_params p{map_resolution_in, k_new_para, theta_p_para, theta_max_para, weightType_para, 0, 0};
        status_ = pf_odom_create(&p, 0, &h_);
        if (status_ != PF_OK) std::fprintf(stderr, "Odom_ES_EstimationClass::init: %s\n", pf_last_error());
        odom = Eigen::Isometry3d::Identity();
    }
    void initMapWithPoints(const pcl::PointCloud<PointType>::Ptr& edge_in, const pcl::PointCloud<PointType>::Ptr& surf_in) {     // :217-222
        if (!h_) return;
        detail::pack_xyz(*edge_in, e_, false);
        detail::pack_xyz(*surf_in, s_, false);
        status_ = pf_odom_init_map(h_, reinterpret_cast<const float*>(e_.data()), (int)e_.size(), reinterpret_cast<const float*>(s_.data()), (int)s_.size());
        if (status_ != PF_OK) std::fprintf(stderr, "initMapWithPoints: %s\n", pf_last_error());
        ++version_;
    }
    void updatePointsToMap(const pcl::PointCloud<PointType>::Ptr& edge_in, const pcl::PointCloud<PointType>::Ptr& surf_in) {     // :229-282
        if (!h_) return;
        detail::pack_xyz(*edge_in, e_, false);
        detail::pack_xyz(*surf_in, s_, false);
        double pose[7];
        status_ = pf_odom_update(h_, reinterpret_cast<const float*>(e_.data()), (int)e_.size(), reinterpret_cast<const float*>(s_.data()), (int)s_.size(), pose);
        if (status_ != PF_OK) { std::fprintf(stderr, "updatePointsToMap: %s\n", pf_last_error()); return; }
        // odom = Identity; odom.linear() = q_w_curr.toRotationMatrix(); odom.translation() = t_w_curr   (:278-280)
        Eigen::Quaterniond q(pose[3], pose[0], pose[1], pose[2]);
        odom = Eigen::Isometry3d::Identity();
        odom.linear() = q.toRotationMatrix();
        odom.translation() = Eigen::Vector3d(pose[4], pose[5], pose[6]);
        ++version_;
    }
    void getMap(pcl::PointCloud<PointType>::Ptr& laserCloudMap) {                                                                // :210-215
        if (!h_) return;
        int ne = 0, ns = 0, n = 0;
        pf_odom_map_size(h_, 0, &ne);
        pf_odom_map_size(h_, 1, &ns);
        buf_.resize((std::size_t)(ne + ns > 0 ? ne + ns : 1));
        status_ = pf_odom_get_map(h_, buf_.data(), ne + ns, &n);
        if (status_ == PF_OK) detail::append_rgb(buf_.data(), n, *laserCloudMap);
    }

    Eigen::Isometry3d odom;                                  // include/odomEstimationClass.h:57
    LazyMapPtr laserCloudCornerMap, laserCloudSurfMap;       // :151-152
    int status() const { return status_; }
    pf_odom* handle() { return h_; }

   private:
    friend class LazyMapPtr;
    pf_odom* h_ = nullptr;
    int status_ = PF_OK;
    long long version_ = 0;
    std::vector<detail::P16> e_, s_;
    std::vector<pf_point> buf_;
};

inline void LazyMapPtr::refresh() {
    if (!owner_ || !owner_->h_ || fetched_version_ == owner_->version_) return;
    int n = 0;
    pf_odom_map_size(owner_->h_, which_, &n);
    owner_->buf_.resize((std::size_t)(n > 0 ? n : 1));
    cloud_.reset(new pcl::PointCloud<PointType>());
    if (n > 0 && pf_odom_get_map_part(owner_->h_, which_, owner_->buf_.data(), n, &n) == PF_OK) detail::append_rgb(owner_->buf_.data(), n, *cloud_);
    fetched_version_ = owner_->version_;
}

typedef Odom_ES_EstimationClass OdomEstimationClass;

class LaserMappingClass {
   public:
    LaserMappingClass() {}
    ~LaserMappingClass() { if (h_) pf_mapping_destroy(h_); }
    LaserMappingClass(const LaserMappingClass&) = delete;
    LaserMappingClass& operator=(const LaserMappingClass&) = delete;

    void init(double map_resolution) {                                                           // src/laserMappingClass.cpp:7-32
        if (h_) { pf_mapping_destroy(h_); h_ = nullptr; }
        status_ = pf_mapping_create(map_resolution, 0, 0, 0, &h_);
        if (status_ != PF_OK) std::fprintf(stderr, "LaserMappingClass::init: %s\n", pf_last_error());
    }
    void updateCurrentPointsToMap(const pcl::PointCloud<pcl::PointXYZI>::Ptr& pc_in, const Eigen::Isometry3d& pose_current) {    // :152-191
        if (!h_) return;
        detail::pack_xyzi(*pc_in, in_);
        double rt[12];
        const auto R = pose_current.linear();
        const auto t = pose_current.translation();
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 3; ++j) rt[4 * i + j] = R(i, j);
            rt[4 * i + 3] = t(i);
        }
        status_ = pf_mapping_update(h_, reinterpret_cast<const float*>(in_.data()), (int)in_.size(), rt);
        if (status_ != PF_OK) std::fprintf(stderr, "updateCurrentPointsToMap: %s\n", pf_last_error());
    }
    pcl::PointCloud<pcl::PointXYZI>::Ptr getMap(void) {                                          // :196-208
        pcl::PointCloud<pcl::PointXYZI>::Ptr c(new pcl::PointCloud<pcl::PointXYZI>());
        if (!h_) return c;
        int n = 0;
        status_ = pf_mapping_map_size(h_, &n);
        if (status_ != PF_OK || n <= 0) return c;
        out_.resize(n);
        status_ = pf_mapping_get_map(h_, reinterpret_cast<float*>(out_.data()), n, &n);
        if (status_ == PF_OK) detail::append_xyzi(out_.data(), n, *c);
        return c;
    }
    int status() const { return status_; }

   private:
    pf_mapping* h_ = nullptr;
    int status_ = PF_OK;
    std::vector<detail::P16> in_, out_;
};

}  // namespace compat
}  // namespace pfilter_b200
