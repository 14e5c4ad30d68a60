// ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into the product.
//
// CPU restatement of LaserMappingClass (/root/reference/src/laserMappingClass.cpp, include/laserMappingClass.h).
// The reference source cannot be compiled here (it needs PCL, Eigen and pcl_ros, include/laserMappingClass.h:5-14);
// PARITY UNPINNED for the two third-party pieces restated below:
//   * pcl::transformPointCloud(float) -- PCL 1.10 pcl/common/impl/transforms.hpp, detail::Transformer<float> (SSE2 path on
//     x86-64): out = m.col(0)*x + (m.col(1)*y + (m.col(2)*z + m.col(3))), all float, no FMA;
//   * pcl::VoxelGrid<PointXYZI>::applyFilter, downsample_all_data = true: key = floor(p * (1/leaf)) - min_b, stable order
//     inside a voxel = ascending input index, CentroidPoint accumulators (float sums of x, y, z, intensity) / float(n).
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <tuple>
#include <vector>

namespace {

struct P4 { float x, y, z, i; };

constexpr double kCell = 50.0;   // LASER_CELL_WIDTH / HEIGHT / DEPTH (include/laserMappingClass.h:23-25)
constexpr int kRangeH = 2, kRangeV = 2;   // LASER_CELL_RANGE_* (:29-30)

void voxel_grid_xyzi(std::vector<P4>& cloud, float leaf) {
    if (cloud.empty()) return;
    const float inv = 1.0f / leaf;
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (const P4& p : cloud) {
        mn[0] = std::min(mn[0], p.x); mn[1] = std::min(mn[1], p.y); mn[2] = std::min(mn[2], p.z);
        mx[0] = std::max(mx[0], p.x); mx[1] = std::max(mx[1], p.y); mx[2] = std::max(mx[2], p.z);
    }
    int64_t d[3];
    for (int a = 0; a < 3; ++a) d[a] = (int64_t)((mx[a] - mn[a]) * inv) + 1;
    if (d[0] * d[1] * d[2] > (int64_t)INT32_MAX) return;   // "Leaf size is too small": output = input
    int minb[3], divb[3];
    for (int a = 0; a < 3; ++a) {
        minb[a] = (int)std::floor(mn[a] * inv);
        divb[a] = (int)std::floor(mx[a] * inv) - minb[a] + 1;
    }
    struct KI { unsigned key, idx; };
    std::vector<KI> iv(cloud.size());
    for (size_t i = 0; i < cloud.size(); ++i) {
        const int i0 = (int)(std::floor(cloud[i].x * inv) - (float)minb[0]);
        const int i1 = (int)(std::floor(cloud[i].y * inv) - (float)minb[1]);
        const int i2 = (int)(std::floor(cloud[i].z * inv) - (float)minb[2]);
        iv[i] = {(unsigned)(i0 + i1 * divb[0] + i2 * divb[0] * divb[1]), (unsigned)i};
    }
    std::stable_sort(iv.begin(), iv.end(), [](const KI& a, const KI& b) { return a.key < b.key; });
    std::vector<P4> out;
    for (size_t s = 0; s < iv.size();) {
        size_t e = s + 1;
        while (e < iv.size() && iv[e].key == iv[s].key) ++e;
        float sx = 0, sy = 0, sz = 0, si = 0;
        for (size_t k = s; k < e; ++k) {
            const P4& p = cloud[iv[k].idx];
            sx += p.x; sy += p.y; sz += p.z; si += p.i;
        }
        const float n = (float)(e - s);
        out.push_back({sx / n, sy / n, sz / n, si / n});
        s = e;
    }
    cloud.swap(out);
}

struct OracleMapping {
    float leaf;
    // the reference's growable 3-D array of cell clouds, keyed by the cell coordinates relative to the initial origin;
    // a key is present <=> the reference's pointer is non-NULL (init :7-22 creates the 5x5x5 block around cell (0,0,0))
    std::map<std::tuple<int, int, int>, std::vector<P4>> cells;
    long dropped = 0;

    explicit OracleMapping(double res) : leaf((float)res) {
        for (int i = -kRangeH; i <= kRangeH; ++i)
            for (int j = -kRangeH; j <= kRangeH; ++j)
                for (int k = -kRangeV; k <= kRangeV; ++k) cells[{i, j, k}];
    }

    // updateCurrentPointsToMap, src/laserMappingClass.cpp:152-191.  R row-major (double), t translation.
    void update(const P4* pc, int n, const double R[9], const double t[3]) {
        const int cx = (int)std::floor(t[0] / kCell + 0.5), cy = (int)std::floor(t[1] / kCell + 0.5), cz = (int)std::floor(t[2] / kCell + 0.5);
        for (int i = cx - kRangeH; i <= cx + kRangeH; ++i)            // checkPoints :110-149
            for (int j = cy - kRangeH; j <= cy + kRangeH; ++j)
                for (int k = cz - kRangeV; k <= cz + kRangeV; ++k) cells[{i, j, k}];
        float m[12];                                                   // pose_current.cast<float>()
        for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) m[4 * r + c] = (float)R[3 * r + c]; m[4 * r + 3] = (float)t[r]; }
        for (int q = 0; q < n; ++q) {
            const P4& s = pc[q];
            P4 o;
            o.x = m[0] * s.x + (m[1] * s.y + (m[2] * s.z + m[3]));
            o.y = m[4] * s.x + (m[5] * s.y + (m[6] * s.z + m[7]));
            o.z = m[8] * s.x + (m[9] * s.y + (m[10] * s.z + m[11]));
            o.i = (float)std::min(1.0, std::max((double)s.z + 2.0, 0.0) / 5);          // :169
            const int px = (int)std::floor((double)o.x / kCell + 0.5), py = (int)std::floor((double)o.y / kCell + 0.5),
                      pz = (int)std::floor((double)o.z / kCell + 0.5);                  // :170-172
            auto it = cells.find({px, py, pz});
            if (it == cells.end()) { ++dropped; continue; }   // the reference dereferences a NULL / out-of-range cell here (UB)
            it->second.push_back(o);
        }
        for (int i = cx - kRangeH; i <= cx + kRangeH; ++i)            // :180-189
            for (int j = cy - kRangeH; j <= cy + kRangeH; ++j)
                for (int k = cz - kRangeV; k <= cz + kRangeV; ++k) voxel_grid_xyzi(cells[{i, j, k}], leaf);
    }

    // getMap :196-208: every non-NULL cell in i, j, k order
    std::vector<P4> get_map() const {
        std::vector<P4> out;
        for (const auto& kv : cells) out.insert(out.end(), kv.second.begin(), kv.second.end());
        return out;
    }
};

}  // namespace

extern "C" {

void* pforacle_mapping_create(double map_resolution) { return new OracleMapping(map_resolution); }
void pforacle_mapping_destroy(void* h) { delete (OracleMapping*)h; }
// rt = row-major 3x4 [R | t] of the Eigen::Isometry3d pose_current (double)
int pforacle_mapping_update(void* h, const float* xyzi, int n, const double rt[12]) {
    const double R[9] = {rt[0], rt[1], rt[2], rt[4], rt[5], rt[6], rt[8], rt[9], rt[10]};
    const double t[3] = {rt[3], rt[7], rt[11]};
    ((OracleMapping*)h)->update((const P4*)xyzi, n, R, t);
    return 0;
}
long pforacle_mapping_size(void* h) { return (long)((OracleMapping*)h)->get_map().size(); }
long pforacle_mapping_dropped(void* h) { return ((OracleMapping*)h)->dropped; }
int pforacle_mapping_get_map(void* h, float* out) {
    const std::vector<P4> m = ((OracleMapping*)h)->get_map();
    if (!m.empty()) std::memcpy(out, m.data(), m.size() * sizeof(P4));
    return (int)m.size();
}

}  // extern "C"
