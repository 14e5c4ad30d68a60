// Synthetic LiDAR sequence generator (host only, deterministic).
//
// Produces HDL-64E- / VLP-32-shaped scans of an analytic "planes + poles" street scene along a
// known trajectory, as specified in SURVEY.md section 8 row D2.  This is test/bench DATA, not part
// of the hot path and not part of the oracle.  The ring elevation tables are chosen so that every
// ray lands well inside the ring bins of the reference's ring classifier
// (/root/reference/src/laserProcessingClass.cpp:38-57), so ring ids are unambiguous.
//
// Output layout: float4 {x, y, z, intensity}, ring-major, azimuth ascending, rays with no return
// inside the range gate are dropped (like a real sensor).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "pf_synth.h"

namespace {

struct Rect {      // axis-aligned vertical rectangle: plane axis (0: x = c, 1: y = c), extent in the other axis and z
    int axis;
    double c, lo, hi, z0, z1;
};
struct Pole {      // vertical cylinder
    double x, y, r, z0, z1;
};
struct Volume {    // axis-aligned box of scattering medium (foliage): a ray ends at an exponentially distributed depth inside it
    double lo[3], hi[3], density;   // density = expected returns per metre of path
};
struct Scene {
    double ground_z;
    std::vector<Rect> rects;
    std::vector<Pole> poles;
    std::vector<Volume> volumes;
};

inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
inline double u01(uint64_t h) { return ((h >> 11) + 0.5) * (1.0 / 9007199254740992.0); }

struct Rng {
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed) {}
    uint64_t next() { s = splitmix64(s); return s; }
    double uni() { return u01(next()); }
    double uni(double a, double b) { return a + (b - a) * uni(); }
};

void add_box(Scene& sc, double x0, double x1, double y0, double y1, double z0, double z1) {
    sc.rects.push_back({0, x0, y0, y1, z0, z1});
    sc.rects.push_back({0, x1, y0, y1, z0, z1});
    sc.rects.push_back({1, y0, x0, x1, z0, z1});
    sc.rects.push_back({1, y1, x0, x1, z0, z1});
}

// "planes+poles": street corridor along +x.  Facades at y = +-12 m with 4 m deep recesses every
// 20 m, 6 m tall; poles r = 0.15 m every 10 m at y = +-8 m; a few seeded boxes (parked vehicles).
Scene build_street(uint64_t seed, double x_begin, double x_end) {
    Scene sc;
    sc.ground_z = -1.73;
    const double zt = sc.ground_z + 6.0;
    for (int side = -1; side <= 1; side += 2) {
        for (double x = x_begin; x < x_end; x += 20.0) {
            double y_front = side * 12.0, y_back = side * 16.0;
            sc.rects.push_back({1, y_front, x, x + 10.0, sc.ground_z, zt});
            sc.rects.push_back({1, y_back, x + 10.0, x + 20.0, sc.ground_z, zt});
            double ylo = side > 0 ? y_front : y_back, yhi = side > 0 ? y_back : y_front;
            sc.rects.push_back({0, x + 10.0, ylo, yhi, sc.ground_z, zt});
            sc.rects.push_back({0, x + 20.0, ylo, yhi, sc.ground_z, zt});
        }
        for (double x = x_begin + 5.0; x < x_end; x += 10.0)
            sc.poles.push_back({x, side * 8.0, 0.15, sc.ground_z, sc.ground_z + 5.0});
    }
    Rng rng(seed * 7919ull + 17);
    int nbox = (int)((x_end - x_begin) / 15.0);
    for (int i = 0; i < nbox; ++i) {
        double cx = rng.uni(x_begin, x_end);
        double side = rng.uni() < 0.5 ? -1.0 : 1.0;
        double cy = side * rng.uni(4.5, 10.0);
        double lx = rng.uni(1.5, 4.5), ly = rng.uni(1.2, 2.2), h = rng.uni(1.2, 2.8);
        add_box(sc, cx - lx / 2, cx + lx / 2, cy - ly / 2, cy + ly / 2, sc.ground_z, sc.ground_z + h);
    }
    // far end walls close the corridor so that long rays terminate
    sc.rects.push_back({0, x_begin - 1.0, -16.0, 16.0, sc.ground_z, zt});
    sc.rects.push_back({0, x_end + 1.0, -16.0, 16.0, sc.ground_z, zt});
    return sc;
}

// "campus": cluttered open area (random vertical walls + poles) for the slow loop trajectory.
Scene build_campus(uint64_t seed, double half) {
    Scene sc;
    sc.ground_z = -1.73;
    Rng rng(seed * 104729ull + 3);
    int nb = 260;
    for (int i = 0; i < nb; ++i) {
        double cx = rng.uni(-half, half), cy = rng.uni(-half, half);
        double r = std::sqrt(cx * cx + cy * cy);
        double lx = rng.uni(2.0, 14.0), ly = rng.uni(2.0, 14.0), h = rng.uni(2.0, 9.0);
        // keep the driving loop (radius 60 m about the origin) free: no part of a box within 4 m of it -- a box CENTRE outside the
        // band is not enough, the sensor would drive through the corner of a large box and lose every return inside the range gate
        const double half_diag = 0.5 * std::sqrt(lx * lx + ly * ly);
        if (r - half_diag < 64.0 && r + half_diag > 56.0) continue;
        add_box(sc, cx - lx / 2, cx + lx / 2, cy - ly / 2, cy + ly / 2, sc.ground_z, sc.ground_z + h);
    }
    for (int i = 0; i < 400; ++i) {
        double cx = rng.uni(-half, half), cy = rng.uni(-half, half);
        double r = std::sqrt(cx * cx + cy * cy);
        if (r > 57.0 && r < 63.0) continue;
        sc.poles.push_back({cx, cy, rng.uni(0.08, 0.3), sc.ground_z, sc.ground_z + rng.uni(2.5, 7.0)});
    }
    add_box(sc, -half - 1, half + 1, -half - 1, half + 1, sc.ground_z, sc.ground_z + 8.0);
    return sc;
}

// "dense campus": the campus plus volumetric scatter (tree crowns, hedges) so that almost every return opens a new voxel -- the
// scene that makes the persistent local map of BASELINE.json configs[3] grow into the millions when the filter keeps the points.
Scene build_campus_dense(uint64_t seed, double half) {
    Scene sc = build_campus(seed, half);
    Rng rng(seed * 15485863ull + 11);
    for (int i = 0; i < 900; ++i) {
        double cx = rng.uni(-half, half), cy = rng.uni(-half, half);
        double r = std::sqrt(cx * cx + cy * cy);
        if (r > 56.0 && r < 64.0) continue;
        double sx = rng.uni(2.0, 9.0), sy = rng.uni(2.0, 9.0), z0 = sc.ground_z + rng.uni(0.0, 3.0), h = rng.uni(2.0, 10.0);
        Volume v{{cx - sx / 2, cy - sy / 2, z0}, {cx + sx / 2, cy + sy / 2, z0 + h}, rng.uni(0.15, 0.6)};
        sc.volumes.push_back(v);
    }
    return sc;
}

// "dense street": the planes+poles street with volumetric scatter (hedges, tree crowns) on both sides of the lane -- the odometry keeps
// tracking here for thousands of frames (the campus loop does not: the reference algorithm's yaw diverges after ~540 frames), and
// almost every scatter return opens a new voxel, which is what makes a persistent local map grow.
Scene build_street_dense(uint64_t seed, double x_begin, double x_end) {
    Scene sc = build_street(seed, x_begin, x_end);
    Rng rng(seed * 32452843ull + 5);
    const int nvol = (int)((x_end - x_begin) / 1.5);
    for (int i = 0; i < nvol; ++i) {
        const double cx = rng.uni(x_begin, x_end), side = rng.uni() < 0.5 ? -1.0 : 1.0;
        const double sx = rng.uni(1.5, 6.0), sy = rng.uni(1.0, 4.0), z0 = sc.ground_z + rng.uni(0.0, 3.0), h = rng.uni(1.5, 6.0);
        const double cy = side * rng.uni(4.0 + sy / 2, 11.0);       // clear of the lane (the trajectory weaves within |y| < 1.5)
        Volume v{{cx - sx / 2, cy - sy / 2, z0}, {cx + sx / 2, cy + sy / 2, z0 + h}, rng.uni(0.1, 0.5)};
        sc.volumes.push_back(v);
    }
    return sc;
}

struct Pose2 { double x, y, z, yaw; };

Pose2 trajectory(const pf_synth_params& p, int k) {
    Pose2 q{0, 0, 0, 0};
    if (p.trajectory == PF_SYNTH_TRAJ_STREET) {
        // ~1 m/frame along +x with a gentle weave; yaw = 0.1 sin(2 pi k / 100)
        q.x = p.speed * k;
        q.yaw = 0.1 * std::sin(2.0 * M_PI * k / 100.0);
        q.y = 1.5 * std::sin(2.0 * M_PI * k / 100.0);
        q.z = 0.0;
    } else {
        // closed loop of radius 60 m, p.speed metres of arc per frame
        double R = 60.0, a = p.speed * k / R;
        q.x = R * std::sin(a);
        q.y = R * (1.0 - std::cos(a)) - R;
        q.yaw = a;
    }
    return q;
}

double elevation_deg(int sensor, int ring) {
    if (sensor == 64) return ring < 32 ? 1.95 - ring / 3.0 : -8.68 - (ring - 32) / 2.0;
    if (sensor == 32) return -92.0 / 3.0 + (ring + 0.5) * 4.0 / 3.0;
    return -15.0 + 2.0 * ring;   // 16 lines: bin centres of int((ang+15)/2+0.5)
}

struct Cache {
    pf_synth_params key;
    bool valid = false;
    Scene scene;
};
Cache g_cache;

const Scene& scene_for(const pf_synth_params& p) {
    if (!g_cache.valid || g_cache.key.seed != p.seed || g_cache.key.scene != p.scene) {
        g_cache.scene = p.scene == PF_SYNTH_SCENE_STREET ? build_street(p.seed, -60.0, 1200.0)
                      : p.scene == PF_SYNTH_SCENE_CAMPUS ? build_campus(p.seed, 110.0)
                      : p.scene == PF_SYNTH_SCENE_CAMPUS_DENSE ? build_campus_dense(p.seed, 110.0) : build_street_dense(p.seed, -60.0, 1200.0);
        g_cache.key = p;
        g_cache.valid = true;
    }
    return g_cache.scene;
}

}  // namespace

extern "C" void pf_synth_default_params(pf_synth_params* p) {
    std::memset(p, 0, sizeof(*p));
    p->sensor_lines = 64;
    p->azimuth_steps = 1800;
    p->seed = 2022;
    p->scene = PF_SYNTH_SCENE_STREET;
    p->trajectory = PF_SYNTH_TRAJ_STREET;
    p->speed = 1.0;
    p->range_sigma = 0.02;
    p->elev_jitter_deg = 0.02;
    p->min_range = 3.2;
    p->max_range = 88.0;
}

extern "C" void pf_synth_pose(const pf_synth_params* p, int frame, double pose[7]) {
    Pose2 q = trajectory(*p, frame);
    pose[0] = 0; pose[1] = 0; pose[2] = std::sin(q.yaw / 2); pose[3] = std::cos(q.yaw / 2);
    pose[4] = q.x; pose[5] = q.y; pose[6] = q.z;
}

extern "C" int pf_synth_scan(const pf_synth_params* p, int frame, float* out_xyzi, int cap) {
    const Scene& sc = scene_for(*p);
    Pose2 pose = trajectory(*p, frame);
    const double cy = std::cos(pose.yaw), sy = std::sin(pose.yaw);

    // cull primitives to those within reach of the sensor
    const double reach = p->max_range + 30.0;
    std::vector<Rect> rects;
    std::vector<Pole> poles;
    for (const Rect& r : sc.rects) {
        double px = r.axis == 0 ? r.c : 0.5 * (r.lo + r.hi), py = r.axis == 1 ? r.c : 0.5 * (r.lo + r.hi);
        double ext = 0.5 * (r.hi - r.lo);
        if (std::hypot(px - pose.x, py - pose.y) < reach + ext) rects.push_back(r);
    }
    for (const Pole& c : sc.poles)
        if (std::hypot(c.x - pose.x, c.y - pose.y) < reach) poles.push_back(c);
    std::vector<Volume> vols;
    for (const Volume& v : sc.volumes)
        if (std::hypot(0.5 * (v.lo[0] + v.hi[0]) - pose.x, 0.5 * (v.lo[1] + v.hi[1]) - pose.y) < reach) vols.push_back(v);

    // Azimuth bins around the sensor: a ray only meets primitives whose subtended azimuth interval (one bin of margin either side)
    // contains its own azimuth.  Pure culling: the surviving hit set, and therefore every output bit, is unchanged.
    constexpr int kBins = 720;
    std::vector<std::vector<int>> brect(kBins), bpole(kBins), bvol(kBins);
    auto add_arc = [&](std::vector<std::vector<int>>& bins, int id, double centre, double half_width) {
        if (half_width >= M_PI - 0.02) { for (int b = 0; b < kBins; ++b) bins[b].push_back(id); return; }
        const double w = 2.0 * M_PI / kBins;
        const int b0 = (int)std::floor((centre - half_width + M_PI) / w) - 1, b1 = (int)std::floor((centre + half_width + M_PI) / w) + 1;
        for (int b = b0; b <= b1; ++b) bins[((b % kBins) + kBins) % kBins].push_back(id);
    };
    auto arc_of_points = [&](const double (*pts)[2], int np, double& centre, double& half_width) {
        double mx = 0, my = 0;
        for (int i = 0; i < np; ++i) { mx += pts[i][0] - pose.x; my += pts[i][1] - pose.y; }
        centre = std::atan2(my, mx);
        half_width = 0;
        for (int i = 0; i < np; ++i) {
            double d = std::atan2(pts[i][1] - pose.y, pts[i][0] - pose.x) - centre;
            while (d > M_PI) d -= 2 * M_PI;
            while (d < -M_PI) d += 2 * M_PI;
            half_width = std::max(half_width, std::fabs(d));
        }
    };
    for (size_t i = 0; i < rects.size(); ++i) {
        const Rect& r = rects[i];
        const double pts[2][2] = {{r.axis == 0 ? r.c : r.lo, r.axis == 0 ? r.lo : r.c}, {r.axis == 0 ? r.c : r.hi, r.axis == 0 ? r.hi : r.c}};
        // distance from the sensor to the segment: very close segments subtend almost pi and the mean direction is ill-defined
        const double along = r.axis == 0 ? pose.y : pose.x, across = std::fabs((r.axis == 0 ? pose.x : pose.y) - r.c);
        if (across < 0.5 && along > r.lo - 0.5 && along < r.hi + 0.5) { add_arc(brect, (int)i, 0, M_PI); continue; }
        double c, hw;
        arc_of_points(pts, 2, c, hw);
        add_arc(brect, (int)i, c, hw);
    }
    for (size_t i = 0; i < poles.size(); ++i) {
        const Pole& c = poles[i];
        const double d = std::hypot(c.x - pose.x, c.y - pose.y);
        if (d <= c.r * 1.5) { add_arc(bpole, (int)i, 0, M_PI); continue; }
        add_arc(bpole, (int)i, std::atan2(c.y - pose.y, c.x - pose.x), std::asin(std::min(1.0, c.r / d)));
    }
    for (size_t i = 0; i < vols.size(); ++i) {
        const Volume& v = vols[i];
        if (pose.x > v.lo[0] - 0.5 && pose.x < v.hi[0] + 0.5 && pose.y > v.lo[1] - 0.5 && pose.y < v.hi[1] + 0.5) { add_arc(bvol, (int)i, 0, M_PI); continue; }
        const double pts[4][2] = {{v.lo[0], v.lo[1]}, {v.lo[0], v.hi[1]}, {v.hi[0], v.lo[1]}, {v.hi[0], v.hi[1]}};
        double c, hw;
        arc_of_points(pts, 4, c, hw);
        add_arc(bvol, (int)i, c, hw);
    }

    int n = 0;
    for (int ring = 0; ring < p->sensor_lines; ++ring) {
        const double elev0 = elevation_deg(p->sensor_lines, ring);
        for (int a = 0; a < p->azimuth_steps; ++a) {
            uint64_t h = splitmix64(p->seed * 0x100000001B3ull ^ ((uint64_t)frame << 40) ^ ((uint64_t)ring << 24) ^ (uint64_t)a);
            uint64_t h1 = splitmix64(h), h2 = splitmix64(h1), h3 = splitmix64(h2);
            double elev = (elev0 + p->elev_jitter_deg * (2.0 * u01(h1) - 1.0)) * M_PI / 180.0;
            double az = 2.0 * M_PI * (a + 0.5) / p->azimuth_steps - M_PI;
            // ray in sensor frame
            double ce = std::cos(elev), dxs = ce * std::cos(az), dys = ce * std::sin(az), dz = std::sin(elev);
            // to world
            double dx = cy * dxs - sy * dys, dy = sy * dxs + cy * dys;
            double ox = pose.x, oy = pose.y, oz = pose.z;
            double tbest = 1e30;
            int bin = (int)std::floor((std::atan2(dy, dx) + M_PI) / (2.0 * M_PI / kBins));
            bin = bin < 0 ? 0 : (bin >= kBins ? kBins - 1 : bin);
            if (dz < -1e-9) {
                double t = (sc.ground_z - oz) / dz;
                if (t > 0 && t < tbest) tbest = t;
            }
            for (int ri : brect[bin]) {
                const Rect& r = rects[ri];
                double d = r.axis == 0 ? dx : dy, o = r.axis == 0 ? ox : oy;
                if (std::fabs(d) < 1e-12) continue;
                double t = (r.c - o) / d;
                if (t <= 0 || t >= tbest) continue;
                double u = r.axis == 0 ? oy + t * dy : ox + t * dx, z = oz + t * dz;
                if (u >= r.lo && u <= r.hi && z >= r.z0 && z <= r.z1) tbest = t;
            }
            for (int ci : bpole[bin]) {
                const Pole& c = poles[ci];
                double fx = ox - c.x, fy = oy - c.y;
                double A = dx * dx + dy * dy, B = fx * dx + fy * dy, C = fx * fx + fy * fy - c.r * c.r;
                double disc = B * B - A * C;
                if (disc <= 0 || A < 1e-14) continue;
                double t = (-B - std::sqrt(disc)) / A;
                if (t <= 0 || t >= tbest) continue;
                double z = oz + t * dz;
                if (z >= c.z0 && z <= c.z1) tbest = t;
            }
            for (int vi : bvol[bin]) {      // slab test, then an exponential free path inside the medium
                const Volume& v = vols[vi];
                const double d3[3] = {dx, dy, dz}, o3[3] = {ox, oy, oz};
                double t0 = 0.0, t1 = tbest;
                for (int a = 0; a < 3 && t0 < t1; ++a) {
                    if (std::fabs(d3[a]) < 1e-12) { if (o3[a] < v.lo[a] || o3[a] > v.hi[a]) t1 = -1.0; continue; }
                    double ta = (v.lo[a] - o3[a]) / d3[a], tb = (v.hi[a] - o3[a]) / d3[a];
                    if (ta > tb) std::swap(ta, tb);
                    if (ta > t0) t0 = ta;
                    if (tb < t1) t1 = tb;
                }
                if (!(t0 < t1)) continue;
                const double depth = -std::log(u01(splitmix64(h ^ (0xA24BAED4963EE407ull * (vi + 1))))) / v.density;
                if (t0 + depth < t1) tbest = t0 + depth;
            }
            if (tbest > 1e29) continue;
            // Box-Muller range noise along the ray (also breaks exact curvature ties)
            double g = std::sqrt(-2.0 * std::log(u01(h2))) * std::cos(2.0 * M_PI * u01(h3));
            double range = tbest + p->range_sigma * g;
            double rxy = range * ce;
            if (rxy < p->min_range || rxy > p->max_range) continue;    // keep clear of the 3-90 m gate edges
            if (n >= cap) return -1;
            out_xyzi[4 * n + 0] = (float)(range * dxs);
            out_xyzi[4 * n + 1] = (float)(range * dys);
            out_xyzi[4 * n + 2] = (float)(range * dz);
            out_xyzi[4 * n + 3] = (float)(u01(splitmix64(h3)));
            ++n;
        }
    }
    return n;
}
