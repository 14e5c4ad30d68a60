// ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into the product library.
//
// Plain C++ (no Eigen / PCL / FLANN / Ceres: none of them exists in this image) restatement of the reference's
// odometry + persistence filter + local-map path, following line by line
//   /root/reference/src/odomEstimationClass.cpp:7-25    extractstablepoint (PFilter delete rule)
//   /root/reference/src/odomEstimationClass.cpp:34-134  rgbds (voxel centroid, max r / max g)
//   /root/reference/src/odomEstimationClass.cpp:162-174 pointAssociateToMap
//   /root/reference/src/odomEstimationClass.cpp:182-282 init / initMapWithPoints / updatePointsToMap
//   /root/reference/src/odomEstimationClass.cpp:284-432 addEdgeCostFactor
//   /root/reference/src/odomEstimationClass.cpp:434-578 addSurfCostFactor
//   /root/reference/src/odomEstimationClass.cpp:589-647 addPointsToMap
//   /root/reference/src/lidarOptimization.cpp:12-156    cost functions, SE3 Plus, exp map
// plus restated semantics of the un-vendored third parties the reference calls:
//   PCL 1.10   VoxelGrid / CropBox / ExtractIndices / getMinMax3D / KdTreeFLANN(k=5)
//   FLANN 1.9.1 L2_Simple<float>, exact search, sorted results
//   Eigen 3.3.9 Quaternion <-> matrix, SelfAdjointEigenSolver<Matrix3d>, ColPivHouseholderQR<5x3>
//   Ceres (<= 2.1, version unrecorded) TrustRegionMinimizer + LevenbergMarquardtStrategy + DENSE_QR + HuberLoss(0.1)
//
// PARITY UNPINNED for everything that lives in those third parties: the reference ships no test, golden vector
// or fixture for this path and its odometry sources cannot be compiled here.
// What IS pinned: tests/test_oracle_odom.py checks the numerical kernels of this file against numpy.linalg
// (eigh, lstsq) and finite differences, and the kd-tree search against brute force.
//
// Conventions for what the reference leaves unspecified (std::sort is unstable, FLANN tie order):
//   - points of one voxel are summed in ascending input index          
//   - equal kNN distances are ordered by lower map index               (H4)
// Compile with -ffp-contract=off.
#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <chrono>
#include <vector>

namespace {

struct OPoint {
    float x, y, z;
    uint8_t r, g, b, a;
};
static_assert(sizeof(OPoint) == 16, "16-byte point");

// ---------------------------------------------------------------------------------------------------------
// small dense math
// ---------------------------------------------------------------------------------------------------------
struct V3 { double x, y, z; };
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline double norm(V3 a) { return std::sqrt(dot(a, a)); }

struct Quat { double x, y, z, w; };
// Eigen: q * v  =  v + w*uv + qv x uv, uv = 2 (qv x v)
inline V3 rotate(const Quat& q, V3 v) {
    V3 qv{q.x, q.y, q.z};
    V3 uv = cross(qv, v);
    uv = uv + uv;
    return v + (q.w * uv) + cross(qv, uv);
}
inline Quat qmul(const Quat& a, const Quat& b) {   // Eigen quaternion product
    return {a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y, a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z,
            a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x, a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z};
}
// Eigen Quaternion::toRotationMatrix
inline void quat_to_mat(const Quat& q, double R[9]) {
    const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
    const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
    const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
    const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
    R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
    R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}
// Eigen quaternion from rotation matrix (internal::quaternionbase_assign_impl<.,3,3>)
inline Quat mat_to_quat(const double R[9]) {
    auto m = [&](int i, int j) { return R[3 * i + j]; };
    Quat q;
    double t = m(0, 0) + m(1, 1) + m(2, 2);
    if (t > 0) {
        t = std::sqrt(t + 1.0);
        q.w = 0.5 * t;
        t = 0.5 / t;
        q.x = (m(2, 1) - m(1, 2)) * t;
        q.y = (m(0, 2) - m(2, 0)) * t;
        q.z = (m(1, 0) - m(0, 1)) * t;
    } else {
        int i = 0;
        if (m(1, 1) > m(0, 0)) i = 1;
        if (m(2, 2) > m(i, i)) i = 2;
        int j = (i + 1) % 3, k = (j + 1) % 3;
        t = std::sqrt(m(i, i) - m(j, j) - m(k, k) + 1.0);
        double c[3];
        c[i] = 0.5 * t;
        t = 0.5 / t;
        q.w = (m(k, j) - m(j, k)) * t;
        c[j] = (m(j, i) + m(i, j)) * t;
        c[k] = (m(k, i) + m(i, k)) * t;
        q.x = c[0]; q.y = c[1]; q.z = c[2];
    }
    return q;
}

struct Iso { double R[9]; double t[3]; };   // Eigen::Isometry3d (rotation row-major here, translation)
inline Iso iso_identity() { return {{1, 0, 0, 0, 1, 0, 0, 0, 1}, {0, 0, 0}}; }
inline Iso iso_mul(const Iso& a, const Iso& b) {
    Iso c;
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) c.R[3 * i + j] = a.R[3 * i] * b.R[j] + a.R[3 * i + 1] * b.R[3 + j] + a.R[3 * i + 2] * b.R[6 + j];
        c.t[i] = a.R[3 * i] * b.t[0] + a.R[3 * i + 1] * b.t[1] + a.R[3 * i + 2] * b.t[2] + a.t[i];
    }
    return c;
}
inline Iso iso_inv(const Iso& a) {   // Transform::inverse(Isometry): R^T, -R^T t
    Iso c;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) c.R[3 * i + j] = a.R[3 * j + i];
    for (int i = 0; i < 3; ++i) c.t[i] = -(c.R[3 * i] * a.t[0] + c.R[3 * i + 1] * a.t[1] + c.R[3 * i + 2] * a.t[2]);
    return c;
}

// Symmetric 3x3 eigen decomposition, ascending eigenvalues (Eigen::SelfAdjointEigenSolver<Matrix3d> contract,
// src/odomEstimationClass.cpp:321-326).  Cyclic Jacobi in double; eigenvectors in columns of V (row-major).
void eig3_sym(const double A_in[9], double w[3], double V[9]) {
    double A[9];
    std::memcpy(A, A_in, sizeof(A));
    for (int i = 0; i < 9; ++i) V[i] = (i % 4 == 0) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 64; ++sweep) {
        double off = A[1] * A[1] + A[2] * A[2] + A[5] * A[5];
        double diag = A[0] * A[0] + A[4] * A[4] + A[8] * A[8];
        if (off <= 1e-40 * diag || off == 0.0) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                double apq = A[3 * p + q];
                if (apq == 0.0) continue;
                double theta = (A[3 * q + q] - A[3 * p + p]) / (2.0 * apq);
                double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 3; ++k) {   // A <- A J
                    double akp = A[3 * k + p], akq = A[3 * k + q];
                    A[3 * k + p] = c * akp - s * akq;
                    A[3 * k + q] = s * akp + c * akq;
                }
                for (int k = 0; k < 3; ++k) {   // A <- J^T A
                    double apk = A[3 * p + k], aqk = A[3 * q + k];
                    A[3 * p + k] = c * apk - s * aqk;
                    A[3 * q + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 3; ++k) {
                    double vkp = V[3 * k + p], vkq = V[3 * k + q];
                    V[3 * k + p] = c * vkp - s * vkq;
                    V[3 * k + q] = s * vkp + c * vkq;
                }
            }
    }
    int idx[3] = {0, 1, 2};
    double d[3] = {A[0], A[4], A[8]};
    std::sort(idx, idx + 3, [&](int a, int b) { return d[a] < d[b]; });
    double Vs[9];
    for (int j = 0; j < 3; ++j) {
        w[j] = d[idx[j]];
        for (int k = 0; k < 3; ++k) Vs[3 * k + j] = V[3 * k + idx[j]];
    }
    std::memcpy(V, Vs, sizeof(Vs));
}

// x = argmin |A x - b|, A is rows x 3 (row-major), column-pivoted Householder QR as in
// Eigen::ColPivHouseholderQR::compute/solve (src/odomEstimationClass.cpp:461).
void colpiv_qr_solve(const double* A_in, const double* b_in, int rows, double x[3]) {
    const int cols = 3;
    std::vector<double> A(A_in, A_in + rows * cols), c(b_in, b_in + rows);
    auto a = [&](int i, int j) -> double& { return A[i * cols + j]; };
    double normsU[3], normsD[3], hcoef[3];
    int perm[3] = {0, 1, 2};
    for (int k = 0; k < cols; ++k) {
        double s = 0;
        for (int i = 0; i < rows; ++i) s += a(i, k) * a(i, k);
        normsU[k] = normsD[k] = std::sqrt(s);
    }
    const double eps = std::numeric_limits<double>::epsilon();
    double mx = std::max(normsU[0], std::max(normsU[1], normsU[2]));
    double thr_helper = (mx * eps / rows) * (mx * eps / rows);
    const double downdate_thr = std::sqrt(eps);
    int nonzero = cols;
    for (int k = 0; k < cols; ++k) {
        int big = k;
        for (int j = k + 1; j < cols; ++j) if (normsU[j] > normsU[big]) big = j;
        double big_sq = normsU[big] * normsU[big];
        if (nonzero == cols && big_sq < thr_helper * (double)(rows - k)) nonzero = k;
        if (big != k) {
            for (int i = 0; i < rows; ++i) std::swap(a(i, k), a(i, big));
            std::swap(normsU[k], normsU[big]);
            std::swap(normsD[k], normsD[big]);
            std::swap(perm[k], perm[big]);
        }
        // Householder of column k, rows k..rows-1
        double tail = 0;
        for (int i = k + 1; i < rows; ++i) tail += a(i, k) * a(i, k);
        double c0 = a(k, k), tau, beta;
        if (tail <= std::numeric_limits<double>::min()) {
            tau = 0; beta = c0;
            for (int i = k + 1; i < rows; ++i) a(i, k) = 0;
        } else {
            beta = std::sqrt(c0 * c0 + tail);
            if (c0 >= 0) beta = -beta;
            for (int i = k + 1; i < rows; ++i) a(i, k) /= (c0 - beta);
            tau = (beta - c0) / beta;
        }
        hcoef[k] = tau;
        a(k, k) = beta;
        for (int j = k + 1; j < cols; ++j) {   // apply H = I - tau v v^T to remaining columns
            double s = a(k, j);
            for (int i = k + 1; i < rows; ++i) s += a(i, k) * a(i, j);
            s *= tau;
            a(k, j) -= s;
            for (int i = k + 1; i < rows; ++i) a(i, j) -= s * a(i, k);
        }
        for (int j = k + 1; j < cols; ++j) {   // LAPACK working note 176 norm down-date
            if (normsU[j] != 0) {
                double temp = std::fabs(a(k, j)) / normsU[j];
                temp = (1.0 + temp) * (1.0 - temp);
                temp = temp < 0 ? 0 : temp;
                double r2 = normsU[j] / normsD[j];
                double temp2 = temp * r2 * r2;
                if (temp2 <= downdate_thr) {
                    double s = 0;
                    for (int i = k + 1; i < rows; ++i) s += a(i, j) * a(i, j);
                    normsD[j] = std::sqrt(s);
                    normsU[j] = normsD[j];
                } else {
                    normsU[j] *= std::sqrt(temp);
                }
            }
        }
    }
    x[0] = x[1] = x[2] = 0;
    if (nonzero == 0) return;
    for (int k = 0; k < nonzero; ++k) {   // c <- Q^T c
        double s = c[k];
        for (int i = k + 1; i < rows; ++i) s += a(i, k) * c[i];
        s *= hcoef[k];
        c[k] -= s;
        for (int i = k + 1; i < rows; ++i) c[i] -= s * a(i, k);
    }
    double y[3] = {0, 0, 0};
    for (int i = nonzero - 1; i >= 0; --i) {
        double s = c[i];
        for (int j = i + 1; j < nonzero; ++j) s -= a(i, j) * y[j];
        y[i] = s / a(i, i);
    }
    for (int i = 0; i < nonzero; ++i) x[perm[i]] = y[i];
}

// ---------------------------------------------------------------------------------------------------------
// PCL restatements
// ---------------------------------------------------------------------------------------------------------
struct KeyIdx { unsigned key; unsigned idx; };   // same layout as pcl's cloud_point_index_idx {idx, cloud_point_index}

// Order of the points inside a voxel = order of the centroid's float sums.  The reference sorts (voxel, point) pairs with std::sort
// on the voxel index alone (src/odomEstimationClass.cpp:74, pcl/filters/impl/voxel_grid.hpp), which is not stable: the order is
// whatever libstdc++'s introsort leaves.  Mode 0 (default, the convention shared with the GPU kernels): stable, ascending
// input index.  Mode 1 ("literal"): the reference's own call, std::sort with the comparator on the voxel index only -- with the same
// libstdc++ algorithm this is the reference's summation order; used to measure how far the convention moves the trajectory.
static int g_sort_literal = 0;
static void sort_voxel_pairs(std::vector<KeyIdx>& iv) {
    if (g_sort_literal) std::sort(iv.begin(), iv.end(), [](const KeyIdx& a, const KeyIdx& b) { return a.key < b.key; });
    else std::stable_sort(iv.begin(), iv.end(), [](const KeyIdx& a, const KeyIdx& b) { return a.key < b.key; });
}

// pcl::VoxelGrid<PointXYZRGB>::applyFilter, downsample_all_data = true, no filter field, dense input
// (called through downSamplingToMap, src/odomEstimationClass.cpp:176-180).
void voxel_grid_pcl(const std::vector<OPoint>& in, float leaf, std::vector<OPoint>& out) {
    out.clear();
    if (in.empty()) return;
    const float inv = 1.0f / leaf;   // inverse_leaf_size_ = Array4f::Ones() / leaf_size_
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (const OPoint& p : in) {
        mn[0] = std::min(mn[0], p.x); mn[1] = std::min(mn[1], p.y); mn[2] = std::min(mn[2], p.z);
        mx[0] = std::max(mx[0], p.x); mx[1] = std::max(mx[1], p.y); mx[2] = std::max(mx[2], p.z);
    }
    int64_t d[3];
    for (int a = 0; a < 3; ++a) d[a] = (int64_t)((mx[a] - mn[a]) * inv) + 1;
    if (d[0] * d[1] * d[2] > (int64_t)INT_MAX) { out = in; return; }   // "Leaf size is too small": output = input
    int minb[3], maxb[3], divb[3];
    for (int a = 0; a < 3; ++a) {
        minb[a] = (int)std::floor(mn[a] * inv);
        maxb[a] = (int)std::floor(mx[a] * inv);
        divb[a] = maxb[a] - minb[a] + 1;
    }
    const int mul[3] = {1, divb[0], divb[0] * divb[1]};
    std::vector<KeyIdx> iv(in.size());
    for (size_t i = 0; i < in.size(); ++i) {
        int i0 = (int)(std::floor(in[i].x * inv) - (float)minb[0]);
        int i1 = (int)(std::floor(in[i].y * inv) - (float)minb[1]);
        int i2 = (int)(std::floor(in[i].z * inv) - (float)minb[2]);
        iv[i] = {(unsigned)(i0 * mul[0] + i1 * mul[1] + i2 * mul[2]), (unsigned)i};
    }
    sort_voxel_pairs(iv);
    for (size_t s = 0; s < iv.size();) {
        size_t e = s + 1;
        while (e < iv.size() && iv[e].key == iv[s].key) ++e;
        float sx = 0, sy = 0, sz = 0, sr = 0, sg = 0, sb = 0, sa = 0;   // CentroidPoint accumulators (float)
        for (size_t k = s; k < e; ++k) {
            const OPoint& p = in[iv[k].idx];
            sx += p.x; sy += p.y; sz += p.z;
            sr += (float)p.r; sg += (float)p.g; sb += (float)p.b; sa += (float)p.a;
        }
        float n = (float)(e - s);
        OPoint o;
        o.x = sx / n; o.y = sy / n; o.z = sz / n;
        o.r = (uint8_t)(uint32_t)(sr / n); o.g = (uint8_t)(uint32_t)(sg / n);
        o.b = (uint8_t)(uint32_t)(sb / n); o.a = (uint8_t)(uint32_t)(sa / n);
        out.push_back(o);
        s = e;
    }
}

// OdomBaseClass::rgbds, src/odomEstimationClass.cpp:34-134
void rgbds(const std::vector<OPoint>& in, float leaf, std::vector<OPoint>& out) {
    out.clear();
    if (in.empty()) return;
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (const OPoint& p : in) {   // pcl::getMinMax3D (:43)
        mn[0] = std::min(mn[0], p.x); mn[1] = std::min(mn[1], p.y); mn[2] = std::min(mn[2], p.z);
        mx[0] = std::max(mx[0], p.x); mx[1] = std::max(mx[1], p.y); mx[2] = std::max(mx[2], p.z);
    }
    int minb[3], divb[3];
    for (int a = 0; a < 3; ++a) {   // :46-54, float division
        minb[a] = (int)std::floor(mn[a] / leaf);
        int maxb = (int)std::floor(mx[a] / leaf);
        divb[a] = maxb - minb[a] + 1;
    }
    const int mul[3] = {1, divb[0], divb[0] * divb[1]};
    std::vector<KeyIdx> iv(in.size());
    for (size_t i = 0; i < in.size(); ++i) {   // :61-70
        int i0 = (int)(std::floor(in[i].x / leaf) - (float)minb[0]);
        int i1 = (int)(std::floor(in[i].y / leaf) - (float)minb[1]);
        int i2 = (int)(std::floor(in[i].z / leaf) - (float)minb[2]);
        iv[i] = {(unsigned)(i0 * mul[0] + i1 * mul[1] + i2 * mul[2]), (unsigned)i};
    }
    sort_voxel_pairs(iv);   // :74
    for (size_t s = 0; s < iv.size();) {   // :86-131 (min_points_per_voxel_ = 0)
        size_t e = s + 1;
        while (e < iv.size() && iv[e].key == iv[s].key) ++e;
        float c[4] = {0, 0, 0, 0};
        int r_max = -1;
        float g_max = -1;
        for (size_t k = s; k < e; ++k) {
            const OPoint& p = in[iv[k].idx];
            c[0] += p.x; c[1] += p.y; c[2] += p.z; c[3] += 1.0f;
            if (p.r > r_max) r_max = p.r;
            if (p.g > g_max) g_max = p.g;
        }
        float n = (float)(e - s);
        OPoint o;
        o.x = c[0] / n; o.y = c[1] / n; o.z = c[2] / n;
        o.r = (uint8_t)r_max; o.g = (uint8_t)g_max; o.b = 0; o.a = 255;
        out.push_back(o);
        s = e;
    }
}

// extractstablepoint, src/odomEstimationClass.cpp:7-25 (ExtractIndices keeps order)
void extract_stable(std::vector<OPoint>& map, int k_new, float theta_p, int theta_max) {
    size_t w = 0;
    for (size_t i = 0; i < map.size(); ++i) {
        const OPoint& p = map[i];
        if ((float)p.g < (float)p.r * theta_p && (int)p.r > k_new && (int)p.g < theta_max + 1) continue;
        map[w++] = p;
    }
    map.resize(w);
}

// CropBox + rgbds + extractstablepoint + r update of addPointsToMap, src/odomEstimationClass.cpp:606-647
void map_maintain(std::vector<OPoint>& map, const double center[3], float leaf, int k_new, float theta_p, int theta_max) {
    const float lo[3] = {(float)(center[0] - 100), (float)(center[1] - 100), (float)(center[2] - 100)};
    const float hi[3] = {(float)(center[0] + 100), (float)(center[1] + 100), (float)(center[2] + 100)};
    std::vector<OPoint> tmp;
    tmp.reserve(map.size());
    for (const OPoint& p : map) {   // pcl::CropBox::applyFilter, negative = false
        if ((p.x < lo[0] || p.y < lo[1] || p.z < lo[2]) || (p.x > hi[0] || p.y > hi[1] || p.z > hi[2])) continue;
        tmp.push_back(p);
    }
    rgbds(tmp, leaf, map);
    extract_stable(map, k_new, theta_p, theta_max);
    for (OPoint& p : map) p.r = p.r > 250 ? 255 : (uint8_t)(p.r + 2);   // :634-646
}

// ---------------------------------------------------------------------------------------------------------
// exact 5-NN (FLANN L2_Simple<float>: ((dx*dx)+(dy*dy))+(dz*dz), diff = query - point; ties -> lower index)
// ---------------------------------------------------------------------------------------------------------
inline float dist2f(const float q[3], const OPoint& p) {
    float dx = q[0] - p.x, dy = q[1] - p.y, dz = q[2] - p.z;
    return (dx * dx + dy * dy) + dz * dz;
}
struct Knn5 {
    float d[5];
    int i[5];
    int n = 0;
    void init() { n = 0; for (int k = 0; k < 5; ++k) { d[k] = std::numeric_limits<float>::infinity(); i[k] = -1; } }
    inline bool better(float dd, int ii, int k) const { return dd < d[k] || (dd == d[k] && ii < i[k]); }
    inline void push(float dd, int ii) {
        if (n == 5 && !better(dd, ii, 4)) return;
        int k = n < 5 ? n : 4;
        while (k > 0 && better(dd, ii, k - 1)) { d[k] = d[k - 1]; i[k] = i[k - 1]; --k; }
        d[k] = dd; i[k] = ii;
        if (n < 5) ++n;
    }
};

// kd-tree in the spirit of FLANN's KDTreeSingleIndex (leaf_max_size 15, exact search); used for the timed CPU
// baseline and the full-frame oracle, validated against brute force in the tests.
struct KdTree {
    struct Node { int lo, hi; int dim; float split_lo, split_hi; int left, right; };
    std::vector<Node> nodes;
    std::vector<int> order;
    const std::vector<OPoint>* pts = nullptr;
    float bb_lo[3], bb_hi[3];
    static float coord(const OPoint& p, int d) { return d == 0 ? p.x : (d == 1 ? p.y : p.z); }
    void build(const std::vector<OPoint>& P) {
        pts = &P;
        nodes.clear();
        order.resize(P.size());
        for (size_t i = 0; i < P.size(); ++i) order[i] = (int)i;
        if (P.empty()) return;
        nodes.reserve(P.size() / 4 + 16);
        for (int d = 0; d < 3; ++d) { bb_lo[d] = FLT_MAX; bb_hi[d] = -FLT_MAX; }
        for (const OPoint& p : P)
            for (int d = 0; d < 3; ++d) { bb_lo[d] = std::min(bb_lo[d], coord(p, d)); bb_hi[d] = std::max(bb_hi[d], coord(p, d)); }
        build_rec(0, (int)P.size());
    }
    int build_rec(int lo, int hi) {
        int id = (int)nodes.size();
        nodes.push_back({lo, hi, -1, 0, 0, -1, -1});
        if (hi - lo <= 15) return id;
        float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
        for (int k = lo; k < hi; ++k)
            for (int d = 0; d < 3; ++d) { float c = coord((*pts)[order[k]], d); mn[d] = std::min(mn[d], c); mx[d] = std::max(mx[d], c); }
        int dim = 0;
        for (int d = 1; d < 3; ++d) if (mx[d] - mn[d] > mx[dim] - mn[dim]) dim = d;
        if (mx[dim] == mn[dim]) return id;   // all identical: keep as a (large) leaf
        int mid = (lo + hi) / 2;
        std::nth_element(order.begin() + lo, order.begin() + mid, order.begin() + hi,
                         [&](int a, int b) { return coord((*pts)[a], dim) < coord((*pts)[b], dim); });
        float lmax = -FLT_MAX, rmin = FLT_MAX;
        for (int k = lo; k < mid; ++k) lmax = std::max(lmax, coord((*pts)[order[k]], dim));
        for (int k = mid; k < hi; ++k) rmin = std::min(rmin, coord((*pts)[order[k]], dim));
        nodes[id].dim = dim; nodes[id].split_lo = lmax; nodes[id].split_hi = rmin;
        int l = build_rec(lo, mid);
        int r = build_rec(mid, hi);
        nodes[id].left = l; nodes[id].right = r;
        return id;
    }
    void search(const float q[3], Knn5& res) const {
        res.init();
        if (nodes.empty()) return;
        search_rec(0, q, res);
    }
    void search_rec(int id, const float q[3], Knn5& res) const {
        const Node& nd = nodes[id];
        if (nd.dim < 0) {
            for (int k = nd.lo; k < nd.hi; ++k) res.push(dist2f(q, (*pts)[order[k]]), order[k]);
            return;
        }
        float v = q[nd.dim];
        // left subtree: coord <= split_lo, right subtree: coord >= split_hi (split_lo <= split_hi)
        int first = v <= 0.5f * (nd.split_lo + nd.split_hi) ? nd.left : nd.right;
        int second = first == nd.left ? nd.right : nd.left;
        search_rec(first, q, res);
        float gap = first == nd.left ? nd.split_hi - v : v - nd.split_lo;   // float slab distance to the other side
        // prune only when the slab distance strictly exceeds the current 5th distance (keeps ties by lower index exact);
        // the float slab bound is conservative: (q - c)^2 <= full float distance of any point beyond the split
        if (res.n < 5 || gap <= 0 || gap * gap <= res.d[4]) search_rec(second, q, res);
    }
};

// ---------------------------------------------------------------------------------------------------------
// cost functions + Ceres restatement
// ---------------------------------------------------------------------------------------------------------
struct Residual {
    int kind;       // 0 edge, 1 surf
    V3 p;           // curr_point (sensor frame)
    V3 a, b;        // edge: last_point_a / last_point_b ; surf: a = plane_unit_norm, b.x = negative_OA_dot_norm
    double weight;  // point_weight (0 when weightType == 0)
};

// EdgeAnalyticCostFunction::Evaluate / SurfNormAnalyticCostFunction::Evaluate (src/lidarOptimization.cpp:12-78);
// the 1x7 global Jacobian times PoseSE3Parameterization::ComputeJacobian (:97-104) = its first six columns.
inline void eval_residual(const Residual& R, const Quat& q, V3 t, double* r, double J[6]) {
    V3 lp = rotate(q, R.p) + t;
    if (R.kind == 0) {
        V3 nu = cross(lp - R.a, lp - R.b);
        V3 de = R.a - R.b;
        double de_norm = norm(de), nu_norm = norm(nu);
        double res = nu_norm / de_norm;
        if (R.weight == 1 || R.weight == 2) res = R.weight * res;      // :25-28
        else if (R.weight == 12) res = R.weight * res;
        *r = res;
        if (J) {
            // J = -nu^T/|nu| * skew(de) * [-skew(lp) | I] / |de|
            V3 w = (-1.0 / nu_norm) * nu;
            // row vector w^T * skew(de) = (de x w)^T ... w^T [de]x = -( [de]x w )^T = -(de x w)^T = (w x de)^T
            V3 m = cross(w, de);
            // m^T * (-skew(lp)) = -(m^T [lp]x) = -( (lp x ... ) ) ; m^T [lp]x = (m x lp)^T  =>  -(m x lp) = lp x m
            V3 jr = cross(lp, m);
            J[0] = jr.x / de_norm; J[1] = jr.y / de_norm; J[2] = jr.z / de_norm;
            J[3] = m.x / de_norm; J[4] = m.y / de_norm; J[5] = m.z / de_norm;
        }
    } else {
        double res = dot(R.a, lp) + R.b.x;
        if (R.weight != 0) res = R.weight * res;                         // :62-63
        *r = res;
        if (J) {
            V3 jr = cross(lp, R.a);   // n^T (-[lp]x) = (lp x n)^T
            J[0] = jr.x; J[1] = jr.y; J[2] = jr.z;
            J[3] = R.a.x; J[4] = R.a.y; J[5] = R.a.z;
        }
    }
}

// getTransformFromSe3 (src/lidarOptimization.cpp:106-143)
inline void exp_se3(const double d[6], Quat& q, V3& t) {
    V3 omega{d[0], d[1], d[2]}, upsilon{d[3], d[4], d[5]};
    double theta = norm(omega), half = 0.5 * theta;
    double real = std::cos(half), imag;
    if (theta < 1e-10) {
        double t2 = theta * theta, t4 = t2 * t2;
        imag = 0.5 - 0.0208333 * t2 + 0.000260417 * t4;
    } else {
        imag = std::sin(half) / theta;
    }
    q = {imag * omega.x, imag * omega.y, imag * omega.z, real};
    double Jm[9];
    if (theta < 1e-10) {
        quat_to_mat(q, Jm);
    } else {
        double O[9] = {0, -omega.z, omega.y, omega.z, 0, -omega.x, -omega.y, omega.x, 0};
        double O2[9];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) O2[3 * i + j] = O[3 * i] * O[j] + O[3 * i + 1] * O[3 + j] + O[3 * i + 2] * O[6 + j];
        double c1 = (1 - std::cos(theta)) / (theta * theta), c2 = (theta - std::sin(theta)) / std::pow(theta, 3);
        for (int i = 0; i < 9; ++i) Jm[i] = ((i % 4 == 0) ? 1.0 : 0.0) + c1 * O[i] + c2 * O2[i];
    }
    t = {Jm[0] * upsilon.x + Jm[1] * upsilon.y + Jm[2] * upsilon.z, Jm[3] * upsilon.x + Jm[4] * upsilon.y + Jm[5] * upsilon.z,
         Jm[6] * upsilon.x + Jm[7] * upsilon.y + Jm[8] * upsilon.z};
}

// PoseSE3Parameterization::Plus (src/lidarOptimization.cpp:80-95); x = [qx qy qz qw tx ty tz]
inline void se3_plus(const double x[7], const double d[6], double out[7]) {
    Quat dq; V3 dt;
    exp_se3(d, dq, dt);
    Quat q{x[0], x[1], x[2], x[3]};
    Quat qp = qmul(dq, q);
    V3 tp = rotate(dq, V3{x[4], x[5], x[6]}) + dt;
    out[0] = qp.x; out[1] = qp.y; out[2] = qp.z; out[3] = qp.w; out[4] = tp.x; out[5] = tp.y; out[6] = tp.z;
}

// ceres::HuberLoss(a = 0.1)::Evaluate
inline void huber(double s, double rho[3]) {
    const double a = 0.1, b = a * a;
    if (s > b) {
        double r = std::sqrt(s);
        rho[0] = 2.0 * a * r - b;
        rho[1] = std::max(std::numeric_limits<double>::min(), a / r);
        rho[2] = -rho[1] / (2.0 * s);
    } else {
        rho[0] = s; rho[1] = 1.0; rho[2] = 0.0;
    }
}

// Evaluator: cost = 1/2 sum rho(r^2); robustified residuals / Jacobian rows via ceres::Corrector
// (rho'' <= 0 for Huber, so the corrector is the plain sqrt(rho') scaling).
double evaluate(const std::vector<Residual>& res, const double x[7], std::vector<double>* r_out, std::vector<double>* J_out) {
    Quat q{x[0], x[1], x[2], x[3]};
    V3 t{x[4], x[5], x[6]};
    double cost = 0;
    if (r_out) r_out->resize(res.size());
    if (J_out) J_out->resize(res.size() * 6);
    for (size_t i = 0; i < res.size(); ++i) {
        double r, J[6];
        eval_residual(res[i], q, t, &r, J_out ? J : nullptr);
        double rho[3];
        huber(r * r, rho);
        cost += 0.5 * rho[0];
        double sc = std::sqrt(rho[1]);
        if (J_out) for (int k = 0; k < 6; ++k) (*J_out)[6 * i + k] = sc * J[k];
        if (r_out) (*r_out)[i] = sc * r;
    }
    return cost;
}

// min |A y - b| via Householder QR (ceres DENSE_QR: Eigen HouseholderQR on the (n+6) x 6 augmented system)
bool dense_qr_solve(std::vector<double>& A, std::vector<double>& b, int rows, double y[6]) {
    const int cols = 6;
    for (int k = 0; k < cols; ++k) {
        double tail = 0;
        for (int i = k + 1; i < rows; ++i) tail += A[i * cols + k] * A[i * cols + k];
        double c0 = A[k * cols + k], tau, beta;
        if (tail <= std::numeric_limits<double>::min()) { tau = 0; beta = c0; }
        else {
            beta = std::sqrt(c0 * c0 + tail);
            if (c0 >= 0) beta = -beta;
            for (int i = k + 1; i < rows; ++i) A[i * cols + k] /= (c0 - beta);
            tau = (beta - c0) / beta;
        }
        A[k * cols + k] = beta;
        for (int j = k + 1; j < cols; ++j) {
            double s = A[k * cols + j];
            for (int i = k + 1; i < rows; ++i) s += A[i * cols + k] * A[i * cols + j];
            s *= tau;
            A[k * cols + j] -= s;
            for (int i = k + 1; i < rows; ++i) A[i * cols + j] -= s * A[i * cols + k];
        }
        double s = b[k];
        for (int i = k + 1; i < rows; ++i) s += A[i * cols + k] * b[i];
        s *= tau;
        b[k] -= s;
        for (int i = k + 1; i < rows; ++i) b[i] -= s * A[i * cols + k];
    }
    for (int i = cols - 1; i >= 0; --i) {
        double s = b[i];
        for (int j = i + 1; j < cols; ++j) s -= A[i * cols + j] * y[j];
        if (A[i * cols + i] == 0) return false;
        y[i] = s / A[i * cols + i];
        if (!std::isfinite(y[i])) return false;
    }
    return true;
}

struct LmInfo { int iterations; double final_cost; int successful; };

// ceres::Solve with the options at src/odomEstimationClass.cpp:263-271 (TRUST_REGION, LEVENBERG_MARQUARDT, DENSE_QR,
// max_num_iterations 4, jacobi_scaling, function_tolerance 1e-6, gradient_tolerance 1e-10, parameter_tolerance 1e-8,
// initial radius 1e4, max radius 1e16, min_relative_decrease 1e-3, min/max_lm_diagonal 1e-6 / 1e32).
// max_iter / ftol: the reference's values are 4 and 1e-6 (:265 and the Ceres default); the parity-pinning test runs the same
// state machine to convergence (large max_iter, tiny ftol) against an independent minimiser of the same cost.
LmInfo lm_solve(const std::vector<Residual>& res, double x[7], int max_iter = 4, double ftol = 1e-6) {
    LmInfo info{0, 0.0, 0};
    const int n = (int)res.size();
    if (n == 0) return info;   // no residual blocks: Ceres returns at once, parameters untouched
    std::vector<double> r, J;
    double cost = evaluate(res, x, &r, &J);
    info.final_cost = cost;
    auto gradient_max_norm = [&](const double xx[7], const std::vector<double>& Jm, const std::vector<double>& rm, const double* scale) {
        double g[6] = {0, 0, 0, 0, 0, 0};
        for (int i = 0; i < n; ++i)
            for (int k = 0; k < 6; ++k) g[k] += (Jm[6 * i + k] / (scale ? scale[k] : 1.0)) * rm[i];
        double ng[6], xp[7];
        for (int k = 0; k < 6; ++k) ng[k] = -g[k];
        se3_plus(xx, ng, xp);
        double m = 0;
        for (int k = 0; k < 7; ++k) m = std::max(m, std::fabs(xx[k] - xp[k]));
        return m;
    };
    double gmax = gradient_max_norm(x, J, r, nullptr);
    double scale[6];
    for (int k = 0; k < 6; ++k) {
        double s = 0;
        for (int i = 0; i < n; ++i) s += J[6 * i + k] * J[6 * i + k];
        scale[k] = 1.0 / (1.0 + std::sqrt(s));
    }
    for (int i = 0; i < n; ++i) for (int k = 0; k < 6; ++k) J[6 * i + k] *= scale[k];
    if (gmax <= 1e-10) return info;
    double radius = 1e4, decrease = 2.0;
    bool reuse_diag = false;
    double diag[6];
    double x_norm = 0;
    for (int k = 0; k < 7; ++k) x_norm += x[k] * x[k];
    x_norm = std::sqrt(x_norm);
    for (int iter = 1; iter <= max_iter; ++iter) {
        info.iterations = iter;
        if (!reuse_diag) {
            for (int k = 0; k < 6; ++k) {
                double s = 0;
                for (int i = 0; i < n; ++i) s += J[6 * i + k] * J[6 * i + k];
                diag[k] = std::min(std::max(s, 1e-6), 1e32);
            }
        }
        std::vector<double> A((size_t)(n + 6) * 6, 0.0), b(n + 6, 0.0);
        std::memcpy(A.data(), J.data(), sizeof(double) * n * 6);
        for (int k = 0; k < 6; ++k) A[(size_t)(n + k) * 6 + k] = std::sqrt(diag[k] / radius);
        std::memcpy(b.data(), r.data(), sizeof(double) * n);
        double step[6];
        bool ok = dense_qr_solve(A, b, n + 6, step);
        for (int k = 0; k < 6; ++k) step[k] = -step[k];
        reuse_diag = true;
        double model_cost_change = 0;
        if (ok) {
            for (int i = 0; i < n; ++i) {
                double mr = 0;
                for (int k = 0; k < 6; ++k) mr += J[6 * i + k] * step[k];
                model_cost_change -= mr * (r[i] + mr / 2.0);
            }
        }
        if (!ok || !(model_cost_change > 0)) {   // invalid step: treated as rejected with zero quality
            radius /= decrease; decrease *= 2.0;
            if (radius < 1e-32) break;
            continue;
        }
        double delta[6], xc[7];
        for (int k = 0; k < 6; ++k) delta[k] = step[k] * scale[k];
        se3_plus(x, delta, xc);
        double cand = evaluate(res, xc, nullptr, nullptr);
        double sn = 0;
        for (int k = 0; k < 7; ++k) sn += (x[k] - xc[k]) * (x[k] - xc[k]);
        if (std::sqrt(sn) <= 1e-8 * (x_norm + 1e-8)) break;            // parameter tolerance (candidate not applied)
        double cost_change = cost - cand;
        if (std::fabs(cost_change) <= ftol * cost) break;              // function tolerance (candidate not applied)
        double rel = cost_change / model_cost_change;
        if (rel > 1e-3) {
            std::memcpy(x, xc, sizeof(double) * 7);
            x_norm = 0;
            for (int k = 0; k < 7; ++k) x_norm += x[k] * x[k];
            x_norm = std::sqrt(x_norm);
            cost = evaluate(res, x, &r, &J);
            info.final_cost = cost;
            info.successful++;
            gmax = gradient_max_norm(x, J, r, nullptr);
            for (int i = 0; i < n; ++i) for (int k = 0; k < 6; ++k) J[6 * i + k] *= scale[k];
            radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * rel - 1.0, 3));
            radius = std::min(1e16, radius);
            decrease = 2.0;
            reuse_diag = false;
            if (gmax <= 1e-10) break;
        } else {
            radius /= decrease; decrease *= 2.0;
        }
        if (radius < 1e-32) break;
    }
    return info;
}

// ---------------------------------------------------------------------------------------------------------
// association passes (addEdgeCostFactor / addSurfCostFactor), weightType 0/1/2/12
// ---------------------------------------------------------------------------------------------------------
struct AssocOut {
    std::vector<uint8_t>* flag = nullptr;   // per query: 0 none, 1 geometric fit ok but skipped, 2 residual
    std::vector<double>* geom = nullptr;    // 8 doubles per query
};

template <class Searcher>
void associate(int kind, const Searcher& knn, std::vector<OPoint>& map, std::vector<OPoint>& queries, const double pose[7],
               int k_new, float theta_p, int theta_max, double weightType, std::vector<Residual>& out, AssocOut dbg) {
    Quat q{pose[0], pose[1], pose[2], pose[3]};
    V3 t{pose[4], pose[5], pose[6]};
    std::vector<double> sparsity, observe_vec;
    std::vector<float> obs_f;
    const size_t first = out.size();
    if (dbg.flag) dbg.flag->assign(queries.size(), 0);
    if (dbg.geom) dbg.geom->assign(queries.size() * 8, 0.0);
    for (size_t i = 0; i < queries.size(); ++i) {
        OPoint& qp = queries[i];
        V3 pc{(double)qp.x, (double)qp.y, (double)qp.z};
        V3 pw = rotate(q, pc) + t;                                   // pointAssociateToMap :162-168
        float qf[3] = {(float)pw.x, (float)pw.y, (float)pw.z};
        Knn5 nn;
        knn(qf, nn);
        if (!(nn.n == 5 && nn.d[4] < 1.0f)) continue;               // :300 / :451
        V3 nb[5];
        for (int j = 0; j < 5; ++j) nb[j] = {(double)map[nn.i[j]].x, (double)map[nn.i[j]].y, (double)map[nn.i[j]].z};
        Residual R{};
        R.kind = kind; R.p = pc; R.weight = 0;
        if (kind == 0) {
            V3 c{0, 0, 0};
            for (int j = 0; j < 5; ++j) c = c + nb[j];                // :304-311
            c = {c.x / 5.0, c.y / 5.0, c.z / 5.0};                    // :312
            double C[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            for (int j = 0; j < 5; ++j) {
                V3 d = nb[j] - c;
                double v[3] = {d.x, d.y, d.z};
                for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) C[3 * a + b] += v[a] * v[b];
            }
            double w[3], V[9];
            eig3_sym(C, w, V);
            if (!(w[2] > 3 * w[1])) continue;                        // :326
            V3 dir{V[2], V[5], V[8]};
            R.a = (0.1 * dir) + c;                                   // :330-331
            R.b = (-0.1 * dir) + c;
        } else {
            double A[15], B[5] = {-1, -1, -1, -1, -1}, nrm[3];
            for (int j = 0; j < 5; ++j) { A[3 * j] = nb[j].x; A[3 * j + 1] = nb[j].y; A[3 * j + 2] = nb[j].z; }
            colpiv_qr_solve(A, B, 5, nrm);                           // :461
            V3 nv{nrm[0], nrm[1], nrm[2]};
            double nn_ = norm(nv);
            double negOA = 1 / nn_;                                  // :462
            nv = {nv.x / nn_, nv.y / nn_, nv.z / nn_};               // normalize()
            bool valid = true;
            for (int j = 0; j < 5; ++j)
                if (std::fabs(nv.x * nb[j].x + nv.y * nb[j].y + nv.z * nb[j].z + negOA) > 0.2) { valid = false; break; }   // :466-476
            if (!valid) continue;
            R.a = nv; R.b = {(double)(float)negOA, 0, 0};   // surfInfo::negative_OA_dot_norm is a float (include/odomEstimationClass.h:93-94)
        }
        // persistence bookkeeping (:332-355 / :480-504) -- sequential: later queries see the incremented g
        int sg = 0, sr = 0;
        for (int j = 0; j < 5; ++j) { sg += map[nn.i[j]].g; sr += map[nn.i[j]].r; }
        float observe = (float)(sg / 5.0 + 1);
        float round = (float)(sr / 5.0);
        for (int j = 0; j < 5; ++j) map[nn.i[j]].g = (uint8_t)std::min(255, map[nn.i[j]].g + 1);
        if (observe / round > 5) observe = 255;
        if (dbg.geom) {
            double* g8 = dbg.geom->data() + 8 * i;
            g8[0] = R.a.x; g8[1] = R.a.y; g8[2] = R.a.z; g8[3] = R.b.x; g8[4] = R.b.y; g8[5] = R.b.z;
        }
        if (observe < round * theta_p && round > k_new && observe < theta_max) {
            if (dbg.flag) (*dbg.flag)[i] = 1;
            continue;
        }
        qp.r = (uint8_t)std::min(255, (int)round);
        qp.g = (uint8_t)std::min(255, (int)observe);
        if (dbg.flag) (*dbg.flag)[i] = 2;
        // point sparsity (:367-385)
        V3 cn{0, 0, 0};
        for (int j = 0; j < 5; ++j) cn = cn + nb[j];
        cn = {cn.x / 5, cn.y / 5, cn.z / 5};
        float sum = 0;
        for (int j = 0; j < 5; ++j) sum += (float)norm(cn - nb[j]);
        sum /= 5.0;
        sparsity.push_back(sum);
        obs_f.push_back(observe);
        out.push_back(R);
    }
    if (weightType == 0) return;
    // weight normalisers (observeMean :136-160, pointSparsityMean include/odomEstimationClass.h:111-126)
    const size_t cnt = out.size() - first;
    if (cnt == 0) return;
    if (weightType == 1 || weightType == 12) {
        observe_vec.assign(obs_f.begin(), obs_f.end());
        double mn = *std::min_element(observe_vec.begin(), observe_vec.end()), mx = *std::max_element(observe_vec.begin(), observe_vec.end());
        double len = mx - mn;
        if (len != 0)
            for (double& e : observe_vec) { e = (e - mn) / len; e -= 1.0; e = std::fabs(e); e *= 2.0; e = std::max(0.1, e); }
    }
    if (weightType == 2 || weightType == 12) {
        double mn = *std::min_element(sparsity.begin(), sparsity.end()), mx = *std::max_element(sparsity.begin(), sparsity.end());
        double len = mx - mn;
        if (len != 0)
            for (double& e : sparsity) { e = (e - mn) / len; e -= 1.0; e = std::fabs(e); e *= 2.0; }
    }
    for (size_t k = 0; k < cnt; ++k) {
        double w = 0;
        if (weightType == 1) w = observe_vec[k];
        else if (weightType == 2) w = sparsity[k];
        else if (weightType == 12) w = kind == 0 ? (sparsity[k] + observe_vec[k]) / 2 : (observe_vec[k] + sparsity[k]) / 2;
        out[first + k].weight = w;
    }
}

struct BruteKnn {
    const std::vector<OPoint>* map;
    void operator()(const float q[3], Knn5& res) const {
        res.init();
        for (size_t i = 0; i < map->size(); ++i) res.push(dist2f(q, (*map)[i]), (int)i);
    }
};
struct TreeKnn {
    const KdTree* tree;
    void operator()(const float q[3], Knn5& res) const { tree->search(q, res); }
};

// ---------------------------------------------------------------------------------------------------------
// Odom_ES_EstimationClass
// ---------------------------------------------------------------------------------------------------------
struct OracleOdom {
    double map_resolution; int k_new; float theta_p; int theta_max; double weightType;
    double parameters[7] = {0, 0, 0, 1, 0, 0, 0};
    Iso odom = iso_identity(), last_odom = iso_identity();
    int optimization_count = 2;
    std::vector<OPoint> cornerMap, surfMap;
    std::vector<double> iter_poses;
    int n_edge_ds = 0, n_surf_ds = 0, n_edge_res = 0, n_surf_res = 0, lm_iterations = 0, passes = 0;
    // wall-clock split of the last update (seconds): down-sample, tree build, association, solve, map update
    double t_ds = 0, t_build = 0, t_assoc = 0, t_solve = 0, t_map = 0;

    void init_map(const std::vector<OPoint>& edge, const std::vector<OPoint>& surf) {   // :217-222
        cornerMap.insert(cornerMap.end(), edge.begin(), edge.end());
        surfMap.insert(surfMap.end(), surf.begin(), surf.end());
        optimization_count = 12;
    }

    void update(const std::vector<OPoint>& edge_in, const std::vector<OPoint>& surf_in) {   // :229-282
        if (optimization_count > 2) optimization_count--;
        Iso pred = iso_mul(odom, iso_mul(iso_inv(last_odom), odom));
        last_odom = odom;
        odom = pred;
        Quat q = mat_to_quat(odom.R);
        parameters[0] = q.x; parameters[1] = q.y; parameters[2] = q.z; parameters[3] = q.w;
        parameters[4] = odom.t[0]; parameters[5] = odom.t[1]; parameters[6] = odom.t[2];
        std::vector<OPoint> E, S;
        voxel_grid_pcl(edge_in, (float)map_resolution, E);           // setLeafSize(double->float) :189
        voxel_grid_pcl(surf_in, (float)(map_resolution * 2), S);     // :190
        n_edge_ds = (int)E.size(); n_surf_ds = (int)S.size();
        iter_poses.clear();
        passes = 0;
        if (cornerMap.size() > 10 && surfMap.size() > 50) {          // :247
            KdTree te, ts;
            te.build(cornerMap);
            ts.build(surfMap);
            for (int it = 0; it < optimization_count; ++it) {
                std::vector<Residual> res;
                associate(0, TreeKnn{&te}, cornerMap, E, parameters, k_new, theta_p, theta_max, weightType, res, AssocOut{});
                n_edge_res = (int)res.size();
                associate(1, TreeKnn{&ts}, surfMap, S, parameters, k_new, theta_p, theta_max, weightType, res, AssocOut{});
                n_surf_res = (int)res.size() - n_edge_res;
                LmInfo li = lm_solve(res, parameters);
                lm_iterations = li.iterations;
                iter_poses.insert(iter_poses.end(), parameters, parameters + 7);
                ++passes;
            }
        }
        Quat qf{parameters[0], parameters[1], parameters[2], parameters[3]};
        quat_to_mat(qf, odom.R);                                     // :278-280
        odom.t[0] = parameters[4]; odom.t[1] = parameters[5]; odom.t[2] = parameters[6];
        // addPointsToMap :589-647
        V3 t{parameters[4], parameters[5], parameters[6]};
        for (const OPoint& p : E) {
            V3 w = rotate(qf, V3{(double)p.x, (double)p.y, (double)p.z}) + t;
            OPoint o = p; o.x = (float)w.x; o.y = (float)w.y; o.z = (float)w.z; o.a = 255;
            cornerMap.push_back(o);
        }
        for (const OPoint& p : S) {
            V3 w = rotate(qf, V3{(double)p.x, (double)p.y, (double)p.z}) + t;
            OPoint o = p; o.x = (float)w.x; o.y = (float)w.y; o.z = (float)w.z; o.a = 255;
            surfMap.push_back(o);
        }
        double c[3] = {odom.t[0], odom.t[1], odom.t[2]};
        map_maintain(surfMap, c, (float)map_resolution * 2, k_new, theta_p, theta_max);      // rgbds(tmpSurf, map_resolution * 2), float member
        map_maintain(cornerMap, c, (float)map_resolution, k_new, theta_p, theta_max);
    }
};

// ---------------------------------------------------------------------------------------------------------
// Odom_BPF_EstimationClass (src/odomEstimationClass.cpp:649-1306): the same arithmetic over beam, pillar (line) and facade (plane)
// ---------------------------------------------------------------------------------------------------------
struct OracleOdomBPF {
    double map_resolution; int k_new; float theta_p; int theta_max; double weightType;
    double parameters[7] = {0, 0, 0, 1, 0, 0, 0};
    Iso odom = iso_identity(), last_odom = iso_identity();
    int optimization_count = 2;
    std::vector<OPoint> map[3];      // beam, pillar, facade
    std::vector<double> iter_poses;
    int n_ds[3] = {0, 0, 0}, n_line_res = 0, n_plane_res = 0, lm_iterations = 0, passes = 0;

    void init_map(const std::vector<OPoint> in[3]) {   // :692-698
        for (int k = 0; k < 3; ++k) map[k].insert(map[k].end(), in[k].begin(), in[k].end());
        optimization_count = 12;
    }

    void update(const std::vector<OPoint> in[3]) {     // :706-760
        if (optimization_count > 2) optimization_count--;
        Iso pred = iso_mul(odom, iso_mul(iso_inv(last_odom), odom));
        last_odom = odom;
        odom = pred;
        Quat q = mat_to_quat(odom.R);
        parameters[0] = q.x; parameters[1] = q.y; parameters[2] = q.z; parameters[3] = q.w;
        parameters[4] = odom.t[0]; parameters[5] = odom.t[1]; parameters[6] = odom.t[2];
        const int type[3] = {0, 0, 1};
        const double leaf_mul[3] = {1, 1, 2};                                   // :658-660
        std::vector<OPoint> D[3];
        for (int k = 0; k < 3; ++k) { voxel_grid_pcl(in[k], (float)(map_resolution * leaf_mul[k]), D[k]); n_ds[k] = (int)D[k].size(); }
        iter_poses.clear();
        passes = 0;
        if (map[0].size() > 10 && map[1].size() > 10 && map[2].size() > 50) {   // :720
            KdTree tree[3];
            for (int k = 0; k < 3; ++k) tree[k].build(map[k]);
            for (int it = 0; it < optimization_count; ++it) {
                std::vector<Residual> res;
                for (int k = 0; k < 3; ++k)                                     // addBeam / addPillar / addFacadeCostFactor :733-735
                    associate(type[k], TreeKnn{&tree[k]}, map[k], D[k], parameters, k_new, theta_p, theta_max, weightType, res, AssocOut{});
                n_line_res = n_plane_res = 0;
                for (const Residual& r : res) (r.kind == 0 ? n_line_res : n_plane_res) += 1;
                LmInfo li = lm_solve(res, parameters);
                lm_iterations = li.iterations;
                iter_poses.insert(iter_poses.end(), parameters, parameters + 7);
                ++passes;
            }
        }
        Quat qf{parameters[0], parameters[1], parameters[2], parameters[3]};
        quat_to_mat(qf, odom.R);
        odom.t[0] = parameters[4]; odom.t[1] = parameters[5]; odom.t[2] = parameters[6];
        V3 t{parameters[4], parameters[5], parameters[6]};
        for (int k = 0; k < 3; ++k)                                              // addPointsToMap :1217-1295
            for (const OPoint& p : D[k]) {
                V3 w = rotate(qf, V3{(double)p.x, (double)p.y, (double)p.z}) + t;
                OPoint o = p; o.x = (float)w.x; o.y = (float)w.y; o.z = (float)w.z; o.a = 255;
                map[k].push_back(o);
            }
        double c[3] = {odom.t[0], odom.t[1], odom.t[2]};
        for (int k = 0; k < 3; ++k) map_maintain(map[k], c, (float)map_resolution * (float)leaf_mul[k], k_new, theta_p, theta_max);
    }
};

std::vector<OPoint> from_xyz4(const float* p, int n) {
    std::vector<OPoint> v(n);
    for (int i = 0; i < n; ++i) v[i] = {p[4 * i], p[4 * i + 1], p[4 * i + 2], 0, 0, 0, 255};   // copyPointCloud XYZI -> XYZRGB
    return v;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------
// C entry points (ctypes)
// ---------------------------------------------------------------------------------------------------------
extern "C" {

int pforacle_voxel_downsample(const OPoint* in, int n, float leaf, OPoint* out, int* n_out) {
    std::vector<OPoint> v(in, in + n), o;
    voxel_grid_pcl(v, leaf, o);
    std::memcpy(out, o.data(), o.size() * sizeof(OPoint));
    *n_out = (int)o.size();
    return 0;
}

int pforacle_map_update(const OPoint* in, int n, const double center[3], float leaf, int k_new, float theta_p, int theta_max,
                        OPoint* out, int* n_out) {
    std::vector<OPoint> v(in, in + n);
    map_maintain(v, center, leaf, k_new, theta_p, theta_max);
    std::memcpy(out, v.data(), v.size() * sizeof(OPoint));
    *n_out = (int)v.size();
    return 0;
}

// mode 0: brute force, 1: kd-tree.  Same contract as pf_knn5: idx = -1 / d2 = inf unless d2[4] < 1.
int pforacle_knn5(const OPoint* map, int m, const float* q4, int nq, int mode, int32_t* idx, float* d2) {
    std::vector<OPoint> M(map, map + m);
    KdTree tree;
    if (mode == 1) tree.build(M);
    for (int i = 0; i < nq; ++i) {
        float qf[3] = {q4[4 * i], q4[4 * i + 1], q4[4 * i + 2]};
        Knn5 nn;
        if (mode == 1) tree.search(qf, nn); else BruteKnn{&M}(qf, nn);
        bool ok = nn.n == 5 && nn.d[4] < 1.0f;
        for (int k = 0; k < 5; ++k) {
            idx[5 * i + k] = ok ? nn.i[k] : -1;
            d2[5 * i + k] = ok ? nn.d[k] : std::numeric_limits<float>::infinity();
        }
    }
    return 0;
}

// kd-tree build and query timed separately (CPU baseline of the map-size sweep); same results as pforacle_knn5(mode = 1)
int pforacle_knn5_timed(const OPoint* map, int m, const float* q4, int nq, int32_t* idx, float* d2, double* s_build, double* s_query) {
    std::vector<OPoint> M(map, map + m);
    KdTree tree;
    auto t0 = std::chrono::steady_clock::now();
    tree.build(M);
    auto t1 = std::chrono::steady_clock::now();
    for (int i = 0; i < nq; ++i) {
        float qf[3] = {q4[4 * i], q4[4 * i + 1], q4[4 * i + 2]};
        Knn5 nn;
        tree.search(qf, nn);
        bool ok = nn.n == 5 && nn.d[4] < 1.0f;
        for (int k = 0; k < 5; ++k) {
            idx[5 * i + k] = ok ? nn.i[k] : -1;
            d2[5 * i + k] = ok ? nn.d[k] : std::numeric_limits<float>::infinity();
        }
    }
    auto t2 = std::chrono::steady_clock::now();
    *s_build = std::chrono::duration<double>(t1 - t0).count();
    *s_query = std::chrono::duration<double>(t2 - t1).count();
    return 0;
}

int pforacle_associate(int kind, OPoint* map, int m, OPoint* queries, int nq, const double pose[7], int k_new, float theta_p,
                       int theta_max, uint8_t* flag, double* geom8) {
    std::vector<OPoint> M(map, map + m), Q(queries, queries + nq);
    KdTree tree;
    tree.build(M);
    std::vector<Residual> res;
    std::vector<uint8_t> f;
    std::vector<double> g;
    AssocOut dbg; dbg.flag = &f; dbg.geom = &g;
    associate(kind, TreeKnn{&tree}, M, Q, pose, k_new, theta_p, theta_max, 0.0, res, dbg);
    std::memcpy(map, M.data(), sizeof(OPoint) * m);
    std::memcpy(queries, Q.data(), sizeof(OPoint) * nq);
    std::memcpy(flag, f.data(), nq);
    std::memcpy(geom8, g.data(), sizeof(double) * 8 * nq);
    return 0;
}

static std::vector<Residual> pack_residuals(const double* edge9, int ne, const double* surf7, int ns) {
    std::vector<Residual> res;
    for (int i = 0; i < ne; ++i) {
        const double* e = edge9 + 9 * i;
        res.push_back({0, {e[0], e[1], e[2]}, {e[3], e[4], e[5]}, {e[6], e[7], e[8]}, 0.0});
    }
    for (int i = 0; i < ns; ++i) {
        const double* s = surf7 + 7 * i;
        res.push_back({1, {s[0], s[1], s[2]}, {s[3], s[4], s[5]}, {s[6], 0, 0}, 0.0});
    }
    return res;
}

int pforacle_eval_normal_eq(const double pose[7], const double* edge9, int ne, const double* surf7, int ns, double H21[21], double g6[6],
                            double* cost) {
    std::vector<Residual> res = pack_residuals(edge9, ne, surf7, ns);
    std::vector<double> r, J;
    *cost = evaluate(res, pose, &r, &J);
    for (int i = 0; i < 21; ++i) H21[i] = 0;
    for (int i = 0; i < 6; ++i) g6[i] = 0;
    for (size_t i = 0; i < res.size(); ++i) {
        int k = 0;
        for (int a = 0; a < 6; ++a) {
            g6[a] += J[6 * i + a] * r[i];
            for (int b = a; b < 6; ++b) H21[k++] += J[6 * i + a] * J[6 * i + b];
        }
    }
    return 0;
}

int pforacle_lm_solve(double pose_io[7], const double* edge9, int ne, const double* surf7, int ns, int* iterations, double* final_cost) {
    std::vector<Residual> res = pack_residuals(edge9, ne, surf7, ns);
    LmInfo li = lm_solve(res, pose_io);
    *iterations = li.iterations;
    *final_cost = li.final_cost;
    return 0;
}

// the same state machine with the iteration cap and the function tolerance opened up (pinning tests)
int pforacle_lm_solve_ex(double pose_io[7], const double* edge9, int ne, const double* surf7, int ns, int max_iter, double ftol,
                         int* iterations, double* final_cost) {
    std::vector<Residual> res = pack_residuals(edge9, ne, surf7, ns);
    LmInfo li = lm_solve(res, pose_io, max_iter, ftol);
    *iterations = li.iterations;
    *final_cost = li.final_cost;
    return 0;
}

// 0: stable voxel sort (convention shared with the GPU), 1: the reference's literal std::sort (see sort_voxel_pairs)
int pforacle_set_sort_mode(int literal) { const int old = g_sort_literal; g_sort_literal = literal ? 1 : 0; return old; }

void pforacle_se3_plus(const double x[7], const double d[6], double out[7]) { se3_plus(x, d, out); }
void pforacle_eig3(const double A[9], double w[3], double V[9]) { eig3_sym(A, w, V); }
void pforacle_qr_solve(const double* A, const double* b, int rows, double x[3]) { colpiv_qr_solve(A, b, rows, x); }
// single residual + 6-column Jacobian (for finite-difference checks)
void pforacle_residual(int kind, const double* geom, const double pose[7], double* r, double J[6]) {
    Residual R = kind == 0 ? Residual{0, {geom[0], geom[1], geom[2]}, {geom[3], geom[4], geom[5]}, {geom[6], geom[7], geom[8]}, 0.0}
                           : Residual{1, {geom[0], geom[1], geom[2]}, {geom[3], geom[4], geom[5]}, {geom[6], 0, 0}, 0.0};
    eval_residual(R, Quat{pose[0], pose[1], pose[2], pose[3]}, V3{pose[4], pose[5], pose[6]}, r, J);
}

void* pforacle_odom_create(double map_resolution, int k_new, float theta_p, int theta_max, double weight_type) {
    OracleOdom* o = new OracleOdom();
    o->map_resolution = map_resolution; o->k_new = k_new; o->theta_p = theta_p; o->theta_max = theta_max; o->weightType = weight_type;
    return o;
}
void pforacle_odom_destroy(void* h) { delete (OracleOdom*)h; }
int pforacle_odom_init_map(void* h, const float* edge, int ne, const float* surf, int ns) {
    ((OracleOdom*)h)->init_map(from_xyz4(edge, ne), from_xyz4(surf, ns));
    return 0;
}
int pforacle_odom_update(void* h, const float* edge, int ne, const float* surf, int ns, double pose_out[7]) {
    OracleOdom* o = (OracleOdom*)h;
    o->update(from_xyz4(edge, ne), from_xyz4(surf, ns));
    Quat q = mat_to_quat(o->odom.R);   // what the node reads: Quaterniond(odom.rotation()) (src/odomEstimationNode.cpp:144)
    (void)q;
    std::memcpy(pose_out, o->parameters, sizeof(double) * 7);
    return 0;
}
int pforacle_odom_map_size(void* h, int which) { OracleOdom* o = (OracleOdom*)h; return (int)(which == 0 ? o->cornerMap.size() : o->surfMap.size()); }
int pforacle_odom_get_map(void* h, int which, OPoint* out) {
    OracleOdom* o = (OracleOdom*)h;
    const std::vector<OPoint>& m = which == 0 ? o->cornerMap : o->surfMap;
    std::memcpy(out, m.data(), m.size() * sizeof(OPoint));
    return (int)m.size();
}
int pforacle_odom_iter_poses(void* h, double* out, int cap) {
    OracleOdom* o = (OracleOdom*)h;
    int n = (int)o->iter_poses.size() / 7;
    if (n > cap) n = cap;
    std::memcpy(out, o->iter_poses.data(), sizeof(double) * 7 * n);
    return n;
}
void* pforacle_bpf_create(double map_resolution, int k_new, float theta_p, int theta_max, double weight_type) {
    OracleOdomBPF* o = new OracleOdomBPF();
    o->map_resolution = map_resolution; o->k_new = k_new; o->theta_p = theta_p; o->theta_max = theta_max; o->weightType = weight_type;
    return o;
}
void pforacle_bpf_destroy(void* h) { delete (OracleOdomBPF*)h; }
int pforacle_bpf_frame(void* h, int init, const float* beam, int nb, const float* pillar, int np, const float* facade, int nf, double pose_out[7]) {
    OracleOdomBPF* o = (OracleOdomBPF*)h;
    const std::vector<OPoint> in[3] = {from_xyz4(beam, nb), from_xyz4(pillar, np), from_xyz4(facade, nf)};
    if (init) o->init_map(in); else o->update(in);
    std::memcpy(pose_out, o->parameters, sizeof(double) * 7);
    return 0;
}
int pforacle_bpf_map_size(void* h, int which) { return (int)((OracleOdomBPF*)h)->map[which].size(); }
int pforacle_bpf_get_map(void* h, int which, OPoint* out) {
    const std::vector<OPoint>& m = ((OracleOdomBPF*)h)->map[which];
    std::memcpy(out, m.data(), m.size() * sizeof(OPoint));
    return (int)m.size();
}
void pforacle_bpf_stats(void* h, int out[8]) {
    OracleOdomBPF* o = (OracleOdomBPF*)h;
    out[0] = o->n_ds[0]; out[1] = o->n_ds[1]; out[2] = o->n_ds[2]; out[3] = o->n_line_res; out[4] = o->n_plane_res;
    out[5] = o->passes; out[6] = o->lm_iterations; out[7] = 0;
}
int pforacle_bpf_iter_poses(void* h, double* out, int cap) {
    OracleOdomBPF* o = (OracleOdomBPF*)h;
    int n = (int)o->iter_poses.size() / 7;
    if (n > cap) n = cap;
    std::memcpy(out, o->iter_poses.data(), sizeof(double) * 7 * n);
    return n;
}

void pforacle_odom_stats(void* h, int out[8]) {
    OracleOdom* o = (OracleOdom*)h;
    out[0] = o->n_edge_ds; out[1] = o->n_surf_ds; out[2] = o->n_edge_res; out[3] = o->n_surf_res;
    out[4] = (int)o->cornerMap.size(); out[5] = (int)o->surfMap.size(); out[6] = o->passes; out[7] = o->lm_iterations;
}

}  // extern "C"
