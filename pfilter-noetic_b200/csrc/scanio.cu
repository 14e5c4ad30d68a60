// Wire / disk side of the hot path, on the boundary's own side: the 16-byte float4 scan layout of the C
// ABI from the two forms the reference's callers hold a scan in.  Host code only; the destination is any host buffer -- normally one
// from pf_host_alloc, so that the packed scan is what the H2D copy of pf_frame_process / pf_frame_submit reads.
//   sensor_msgs/PointCloud2   what the nodes receive and run through pcl::fromROSMsg into PointXYZI (/root/reference/src/laserProcessingNode.cpp:52-63)
//   KITTI velodyne *.bin      little-endian float32 x, y, z, reflectance: already the ABI layout
#include <errno.h>
#include <stdio.h>

#include "common.cuh"

using namespace pf;

namespace {
// sensor_msgs/PointField datatype codes: 1 INT8, 2 UINT8, 3 INT16, 4 UINT16, 5 INT32, 6 UINT32, 7 FLOAT32, 8 FLOAT64
inline int field_size(int t) { return t == 1 || t == 2 ? 1 : t == 3 || t == 4 ? 2 : t == 5 || t == 6 || t == 7 ? 4 : t == 8 ? 8 : 0; }
inline float load_field(const uint8_t* p, int t) {
    switch (t) {
        case 1: { int8_t v; memcpy(&v, p, 1); return (float)v; }
        case 2: { uint8_t v; memcpy(&v, p, 1); return (float)v; }
        case 3: { int16_t v; memcpy(&v, p, 2); return (float)v; }
        case 4: { uint16_t v; memcpy(&v, p, 2); return (float)v; }
        case 5: { int32_t v; memcpy(&v, p, 4); return (float)v; }
        case 6: { uint32_t v; memcpy(&v, p, 4); return (float)v; }
        case 7: { float v; memcpy(&v, p, 4); return v; }
        default: { double v; memcpy(&v, p, 8); return (float)v; }
    }
}
}  // namespace

extern "C" int pf_pack_pointcloud2(const uint8_t* data, uint64_t data_bytes, const pf_pc2_layout* L, float* xyzi_out, int cap_points, int* n_points) {
    PF_REQUIRE(L && xyzi_out && n_points && (data || data_bytes == 0), "null argument");
    PF_REQUIRE(!L->is_bigendian, "big-endian PointCloud2 payloads are not supported");
    PF_REQUIRE(L->point_step > 0 && L->height > 0, "bad PointCloud2 layout (point_step %u, height %u)", L->point_step, L->height);
    const uint64_t width = L->width, height = L->height, step = L->point_step;
    const uint64_t row_step = L->row_step ? L->row_step : step * width;
    PF_REQUIRE(row_step >= step * width, "row_step %llu < point_step x width", (unsigned long long)row_step);
    const uint64_t n = width * height;
    PF_REQUIRE(n <= (uint64_t)cap_points, "PointCloud2 of %llu points exceeds the buffer of %d", (unsigned long long)n, cap_points);
    PF_REQUIRE(n == 0 || (height - 1) * row_step + step * width <= data_bytes, "PointCloud2 payload of %llu bytes is shorter than its layout", (unsigned long long)data_bytes);
    const int off[4] = {L->off_x, L->off_y, L->off_z, L->off_intensity};
    const int typ[4] = {L->type_x, L->type_y, L->type_z, L->type_intensity};
    for (int c = 0; c < 4; ++c) {
        if (c == 3 && off[c] < 0) continue;                       // no intensity field: 0, as pcl::fromROSMsg leaves it
        PF_REQUIRE(off[c] >= 0, "PointCloud2 has no %c field", "xyz"[c]);
        PF_REQUIRE(field_size(typ[c]) > 0, "unknown PointField datatype %d", typ[c]);
        PF_REQUIRE((uint64_t)off[c] + field_size(typ[c]) <= step, "field at offset %d does not fit point_step %u", off[c], L->point_step);
    }
    const bool fast = typ[0] == 7 && typ[1] == 7 && typ[2] == 7 && (off[3] < 0 || typ[3] == 7);
    for (uint64_t r = 0; r < height; ++r) {
        const uint8_t* row = data + r * row_step;
        float* o = xyzi_out + 4 * r * width;
        if (fast) {
            for (uint64_t i = 0; i < width; ++i, row += step, o += 4) {
                memcpy(o, row + off[0], 4); memcpy(o + 1, row + off[1], 4); memcpy(o + 2, row + off[2], 4);
                if (off[3] >= 0) memcpy(o + 3, row + off[3], 4); else o[3] = 0.f;
            }
        } else {
            for (uint64_t i = 0; i < width; ++i, row += step, o += 4) {
                for (int c = 0; c < 3; ++c) o[c] = load_field(row + off[c], typ[c]);
                o[3] = off[3] >= 0 ? load_field(row + off[3], typ[3]) : 0.f;
            }
        }
    }
    *n_points = (int)n;
    return PF_OK;
}

extern "C" int pf_read_kitti_bin(const char* path, float* xyzi_out, int cap_points, int* n_points) {
    PF_REQUIRE(path && xyzi_out && n_points && cap_points >= 0, "bad argument");
    FILE* f = fopen(path, "rb");
    PF_REQUIRE(f, "%s: %s", path, strerror(errno));
    long long bytes = -1;
    if (fseek(f, 0, SEEK_END) == 0) bytes = (long long)ftell(f);
    if (bytes < 0 || fseek(f, 0, SEEK_SET) != 0) { fclose(f); set_error("%s: cannot determine the file size", path); return PF_ERR_INVALID; }
    if (bytes % 16 != 0) { fclose(f); set_error("%s: %lld bytes is not a multiple of 16 (x, y, z, reflectance float32)", path, bytes); return PF_ERR_INVALID; }
    if (bytes / 16 > (long long)cap_points) { fclose(f); set_error("%s holds %lld points, the buffer %d", path, bytes / 16, cap_points); return PF_ERR_CAPACITY; }
    const size_t got = fread(xyzi_out, 16, (size_t)(bytes / 16), f);
    fclose(f);
    PF_REQUIRE(got == (size_t)(bytes / 16), "%s: short read", path);
    *n_points = (int)got;
    return PF_OK;
}

extern "C" int pf_write_kitti_bin(const char* path, const float* xyzi, int n_points) {
    PF_REQUIRE(path && (xyzi || n_points == 0) && n_points >= 0, "bad argument");
    FILE* f = fopen(path, "wb");
    PF_REQUIRE(f, "%s: %s", path, strerror(errno));
    const size_t put = fwrite(xyzi, 16, (size_t)n_points, f);
    const int rc = fclose(f);
    PF_REQUIRE(put == (size_t)n_points && rc == 0, "%s: short write", path);
    return PF_OK;
}
