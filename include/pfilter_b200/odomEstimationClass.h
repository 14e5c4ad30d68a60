// Odom_ES_EstimationClass with the reference's interface (/root/reference/include/odomEstimationClass.h:140-167),
// implemented by the CUDA odometry behind the C ABI (pf_odom_*).  Header-only.
//
// The reference exposes the pose as the public member `Eigen::Isometry3d odom` and the maps as the public members
// laserCloudCornerMap / laserCloudSurfMap, which its node reads after every call
// (src/odomEstimationNode.cpp:144-146, :171, :179).  Here `odom` is a small POD with the same accessors the node uses
// (rotation as quaternion, translation()); the maps stay in HBM and are fetched on demand by the two accessor
// functions of the same names (copying ~10^4..10^6 points to the host every frame is exactly the cost the
// device-resident design removes).
#pragma once
#include <cstdio>
#include <cstring>

#include "cloud.h"
#include "lidar.h"

typedef pfilter_b200::PointXYZRGB PointType;

struct OdomPose {
    double q[4] = {0, 0, 0, 1};   // x y z w  (Eigen::Quaterniond(odom.rotation()))
    double t[3] = {0, 0, 0};      // odom.translation()
    const double* translation() const { return t; }
    const double* rotation_quaternion() const { return q; }
    void rotation_matrix(double R[9]) const {   // row-major, Eigen::Quaterniond::toRotationMatrix
        const double tx = 2 * q[0], ty = 2 * q[1], tz = 2 * q[2];
        const double twx = tx * q[3], twy = ty * q[3], twz = tz * q[3], txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
        const double tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
        R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy; R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
        R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
    }
};

class Odom_ES_EstimationClass {
   public:
    using Cloud = pfilter_b200::PointCloud<PointType>;

    Odom_ES_EstimationClass() {}
    ~Odom_ES_EstimationClass() { if (h_) pf_odom_destroy(h_); }
    Odom_ES_EstimationClass(const Odom_ES_EstimationClass&) = delete;
    Odom_ES_EstimationClass& operator=(const Odom_ES_EstimationClass&) = delete;

    void init(lidar::Lidar /*lidar_param*/, double map_resolution_in, int k_new_para, float theta_p_para, int theta_max_para,
              double weightType_para, int device = 0, int max_map_points = 0) {
        if (h_) { pf_odom_destroy(h_); h_ = nullptr; }
        pf_odom_params p{map_resolution_in, k_new_para, theta_p_para, theta_max_para, weightType_para, max_map_points, 0};
        status_ = pf_odom_create(&p, device, &h_);
        if (status_ != PF_OK) std::fprintf(stderr, "Odom_ES_EstimationClass::init: %s\n", pf_last_error());
    }

    void initMapWithPoints(const Cloud::Ptr& edge_in, const Cloud::Ptr& surf_in) {
        if (!h_) return;
        status_ = pf_odom_init_map(h_, reinterpret_cast<const float*>(edge_in->points.data()), (int)edge_in->points.size(),
                                   reinterpret_cast<const float*>(surf_in->points.data()), (int)surf_in->points.size());
        if (status_ != PF_OK) std::fprintf(stderr, "initMapWithPoints: %s\n", pf_last_error());
    }

    void updatePointsToMap(const Cloud::Ptr& edge_in, const Cloud::Ptr& surf_in) {
        if (!h_) return;
        double pose[7];
        status_ = pf_odom_update(h_, reinterpret_cast<const float*>(edge_in->points.data()), (int)edge_in->points.size(),
                                 reinterpret_cast<const float*>(surf_in->points.data()), (int)surf_in->points.size(), pose);
        if (status_ != PF_OK) { std::fprintf(stderr, "updatePointsToMap: %s\n", pf_last_error()); return; }
        std::memcpy(odom.q, pose, sizeof(double) * 4);
        std::memcpy(odom.t, pose + 4, sizeof(double) * 3);
    }

    // *laserCloudMap += surf map; += corner map  (src/odomEstimationClass.cpp:210-215)
    void getMap(Cloud::Ptr& laserCloudMap) {
        if (!h_) return;
        int ne = 0, ns = 0, n = 0;
        pf_odom_map_size(h_, 0, &ne);
        pf_odom_map_size(h_, 1, &ns);
        const size_t old = laserCloudMap->points.size();
        laserCloudMap->points.resize(old + ne + ns);
        status_ = pf_odom_get_map(h_, reinterpret_cast<pf_point*>(laserCloudMap->points.data() + old), ne + ns, &n);
        laserCloudMap->points.resize(old + (status_ == PF_OK ? n : 0));
    }

    Cloud::Ptr laserCloudCornerMap() { return fetch(0); }
    Cloud::Ptr laserCloudSurfMap() { return fetch(1); }

    OdomPose odom;
    int status() const { return status_; }
    pf_odom* handle() { return h_; }

   private:
    Cloud::Ptr fetch(int which) {
        Cloud::Ptr c(new Cloud());
        if (!h_) return c;
        int n = 0;
        pf_odom_map_size(h_, which, &n);
        c->points.resize(n > 0 ? n : 0);
        if (n > 0) pf_odom_get_map_part(h_, which, reinterpret_cast<pf_point*>(c->points.data()), n, &n);
        return c;
    }
    pf_odom* h_ = nullptr;
    int status_ = PF_OK;
};

typedef Odom_ES_EstimationClass OdomEstimationClass;

// Odom_BPF_EstimationClass (/root/reference/include/odomEstimationClass.h:169-205): beam / pillar / facade feature kinds.
class Odom_BPF_EstimationClass {
   public:
    using Cloud = pfilter_b200::PointCloud<PointType>;

    Odom_BPF_EstimationClass() {}
    ~Odom_BPF_EstimationClass() { if (h_) pf_odom_destroy(h_); }
    Odom_BPF_EstimationClass(const Odom_BPF_EstimationClass&) = delete;
    Odom_BPF_EstimationClass& operator=(const Odom_BPF_EstimationClass&) = delete;

    void init(lidar::Lidar /*lidar_param*/, double map_resolution_in, int k_new_para, float theta_p_para, int theta_max_para,
              double weightType_para, int device = 0, int max_map_points = 0) {
        if (h_) { pf_odom_destroy(h_); h_ = nullptr; }
        pf_odom_params p{map_resolution_in, k_new_para, theta_p_para, theta_max_para, weightType_para, max_map_points, 0};
        status_ = pf_odom_bpf_create(&p, device, &h_);
        if (status_ != PF_OK) std::fprintf(stderr, "Odom_BPF_EstimationClass::init: %s\n", pf_last_error());
    }

    void initMapWithPoints(const Cloud::Ptr& beam_in, const Cloud::Ptr& pillar_in, const Cloud::Ptr& facade_in) {
        if (!h_) return;
        status_ = pf_odom_bpf_init_map(h_, data(beam_in), size(beam_in), data(pillar_in), size(pillar_in), data(facade_in), size(facade_in));
        if (status_ != PF_OK) std::fprintf(stderr, "initMapWithPoints: %s\n", pf_last_error());
    }

    void updatePointsToMap(const Cloud::Ptr& beam_in, const Cloud::Ptr& pillar_in, const Cloud::Ptr& facade_in) {
        if (!h_) return;
        double pose[7];
        status_ = pf_odom_bpf_update(h_, data(beam_in), size(beam_in), data(pillar_in), size(pillar_in), data(facade_in), size(facade_in), pose);
        if (status_ != PF_OK) { std::fprintf(stderr, "updatePointsToMap: %s\n", pf_last_error()); return; }
        std::memcpy(odom.q, pose, sizeof(double) * 4);
        std::memcpy(odom.t, pose + 4, sizeof(double) * 3);
    }

    // *laserCloudMap += beam map; += pillar map; += facade map  (src/odomEstimationClass.cpp:683-689)
    void getMap(Cloud::Ptr& laserCloudMap) {
        if (!h_) return;
        int total = 0, n = 0;
        for (int which = 0; which < 3; ++which) { int m = 0; pf_odom_map_size(h_, which, &m); total += m; }
        const size_t old = laserCloudMap->points.size();
        laserCloudMap->points.resize(old + total);
        status_ = pf_odom_get_map(h_, reinterpret_cast<pf_point*>(laserCloudMap->points.data() + old), total, &n);
        laserCloudMap->points.resize(old + (status_ == PF_OK ? n : 0));
    }

    Cloud::Ptr laserCloudBeamMap() { return fetch(0); }
    Cloud::Ptr laserCloudPillarMap() { return fetch(1); }
    Cloud::Ptr laserCloudFacadeMap() { return fetch(2); }

    OdomPose odom;
    int status() const { return status_; }
    pf_odom* handle() { return h_; }

   private:
    static const float* data(const Cloud::Ptr& c) { return reinterpret_cast<const float*>(c->points.data()); }
    static int size(const Cloud::Ptr& c) { return (int)c->points.size(); }
    Cloud::Ptr fetch(int which) {
        Cloud::Ptr c(new Cloud());
        if (!h_) return c;
        int n = 0;
        pf_odom_map_size(h_, which, &n);
        c->points.resize(n > 0 ? n : 0);
        if (n > 0) pf_odom_get_map_part(h_, which, reinterpret_cast<pf_point*>(c->points.data()), n, &n);
        return c;
    }
    pf_odom* h_ = nullptr;
    int status_ = PF_OK;
};
