// K3 / K4: uniform 1 m search grid over a map and exact 5-nearest-neighbour search, one warp per query.
//
// Replaces pcl::KdTreeFLANN::setInputCloud + nearestKSearch(k = 5) (/root/reference/src/odomEstimationClass.cpp:249-250,
// :299, :447).  FLANN L2_Simple<float> distance: ((dx*dx)+(dy*dy))+(dz*dz) with d = query - point, float, no FMA;
// results ascending, ties by lower map index.
// Exactness contract: the reference only uses a result when the 5th distance is < 1.0 (:300, :451).  Every point
// closer than 1 m lies in the 27 cells around the query's 1 m cell (fl(q - p) >= 1 for any point two cells away),
// so the 27-cell search returns exactly FLANN's answer whenever d2[4] < 1; otherwise the query is reported invalid.
#pragma once
#include "primitives.cuh"

namespace pf {

constexpr int kGridCellCap = 1 << 23;   // cells per map (crop box 201^3 < 2^23); key = kind << 23 | cell

struct KnnGrid {           // device view of one map's search structure
    const float4* pts;     // [m] sorted by cell: x, y, z, original index (int bits)
    const int* cell_start; // [dims product]
    const int* cell_end;
    const int* geom;       // device: [0..2] origin cell (floor of min), [3..5] dims
};

struct GridBuild {
    const Pt* map[2];      // the two maps (edge, surf)
    const int* n_map[2];   // device counts
    float4* pts[2];        // outputs, capacity cap[k]
    int* cell_start[2];
    int* cell_end[2];
    int* geom[2];          // [6] each
    unsigned* state;       // filled by build_grids
};

// Enqueues bounds -> keys -> clear -> sort -> fill for both maps. slot: state slot (see primitives.cuh).
int build_grids(Workspace& ws, const GridBuild& G, int slot, int cap0, int cap1);

#ifdef __CUDACC__
struct Top5 {
    unsigned long long k[5];   // (float bits of d2) << 32 | index, ascending; empty = ~0
};

__device__ __forceinline__ void top5_init(Top5& t) {
#pragma unroll
    for (int i = 0; i < 5; ++i) t.k[i] = ~0ull;
}
__device__ __forceinline__ void top5_insert(Top5& t, unsigned long long key) {
    if (key >= t.k[4]) return;
    t.k[4] = key;
#pragma unroll
    for (int i = 4; i > 0; --i) {
        if (t.k[i] < t.k[i - 1]) { unsigned long long s = t.k[i]; t.k[i] = t.k[i - 1]; t.k[i - 1] = s; }
    }
}

// A group of W lanes (a whole warp, or a half warp so that a warp works on two queries at once: the search is a chain of three or
// four dependent memory round trips, so queries in flight are what sets the throughput) cooperates on one query; every lane of
// the group returns the same 5 results (ascending).  Returns true when the result is usable under the reference's contract
// (5 neighbours found and d2[4] < 1).  All W lanes of the group must call it together.
template <int W>
__device__ __forceinline__ bool knn5_group(const KnnGrid& g, float qx, float qy, float qz, int idx[5], float d2[5]) {
    static_assert(W == 32 || W == 16, "group of 16 or 32 lanes");
    const unsigned lane = lane_id() & (unsigned)(W - 1);                               // lane within the group
    const unsigned gm = W == 32 ? 0xffffffffu : (0xffffu << (lane_id() & 16u));       // the group's lanes
    const int ox = g.geom[0], oy = g.geom[1], oz = g.geom[2], dx = g.geom[3], dy = g.geom[4], dz = g.geom[5];
    // floor() of the float coordinate is the exact cell; clamp far-away queries so the int conversion cannot overflow
    const float fx = fminf(fmaxf(floorf(qx), -1.0e9f), 1.0e9f), fy = fminf(fmaxf(floorf(qy), -1.0e9f), 1.0e9f),
                fz = fminf(fmaxf(floorf(qz), -1.0e9f), 1.0e9f);
    const long long cx = (long long)fx - ox, cy = (long long)fy - oy, cz = (long long)fz - oz;
    // lanes 0..8: one row (fixed y, z offset) of three x-adjacent cells = one contiguous range of the sorted points
    int rs = 0, re = 0;
    if (lane < 9) {
        const long long yy = cy + (int)(lane % 3) - 1, zz = cz + (int)(lane / 3) - 1;
        if (yy >= 0 && yy < dy && zz >= 0 && zz < dz) {
            const long long x0 = cx - 1 < 0 ? 0 : cx - 1, x1 = cx + 1 >= dx ? dx - 1 : cx + 1;
            if (x0 <= x1) {
                // the cell sort is a counting sort: cell c holds [cell_start[c], cell_end[c]) and cell_end[c] == cell_start[c + 1], so
                // the three x-adjacent cells of a row are the one range [cell_start[first], cell_end[last]) -- two loads, not six
                const long long base = (zz * dy + yy) * dx;
                const int s = __ldg(g.cell_start + base + x0), e = __ldg(g.cell_end + base + x1);
                if (e > s) { rs = s; re = e; }
            }
        }
    }
    // Every point of a row with cell offsets (ry, rz) is at least my / mz away from the query in y / z, where my, mz are the
    // distances to the cell faces -- exact in float (|q - floor(q)| < 1), and rounding is monotone, so
    // lower = fl(fl(my my) + fl(mz mz)) bounds the float distance the search computes for any point of the row from below.
    // Pass A scans the centre row (offsets 0, 0); `upper` = the fifth smallest of the lanes' best distances so far is an upper bound
    // of the final fifth distance; pass B scans the other rows and skips those with lower > upper (strictly: a skipped point can
    // neither enter the result nor tie with it).  On dense maps this drops most corner rows; the result is unchanged.
    float lower = 0.f;
    if (lane < 9) {
        const int ry = (int)(lane % 3) - 1, rz = (int)(lane / 3) - 1;      // the row's cell offsets in y and z
        const float my = ry == 0 ? 0.f : (ry < 0 ? __fsub_rn(qy, fy) : __fsub_rn(__fadd_rn(fy, 1.0f), qy));
        const float mz = rz == 0 ? 0.f : (rz < 0 ? __fsub_rn(qz, fz) : __fsub_rn(__fadd_rn(fz, 1.0f), qz));
        lower = __fadd_rn(__fmul_rn(my, my), __fmul_rn(mz, mz));
    }
    Top5 t;
    top5_init(t);
    auto scan = [&](int len) {      // len: this lane's row length (lanes >= 9: 0); rows with len == 0 are skipped
        int incl = len;
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) {
            int y = __shfl_up_sync(gm, incl, o, W);
            if (lane >= o) incl += y;
        }
        const int total = __shfl_sync(gm, incl, 8, W);
        const int excl = incl - len;
        for (int i = (int)lane; i < ((total + W - 1) & ~(W - 1)); i += W) {
            // row r with excl[r] <= i < excl[r] + len[r]: the last of the nine rows whose offset does not exceed i (offsets ascend:
            // four probes of a binary search, each lane reading the lane that holds the row it asks about, instead of eight compares)
            int r = (i >= __shfl_sync(gm, excl, 8, W)) ? 8 : 0;
#pragma unroll
            for (int st = 4; st >= 1; st >>= 1) {
                const int probe = r + st;                                 // <= 7 unless r == 8 (then the probes read rows 9..15: empty)
                if (i >= __shfl_sync(gm, excl, probe, W) && probe < 9) r = probe;
            }
            const int rbase = __shfl_sync(gm, rs, r, W), rex = __shfl_sync(gm, excl, r, W);
            if (i < total) {
                const float4 p = __ldg(g.pts + rbase + (i - rex));
                const float ddx = __fsub_rn(qx, p.x), ddy = __fsub_rn(qy, p.y), ddz = __fsub_rn(qz, p.z);
                const float d = __fadd_rn(__fadd_rn(__fmul_rn(ddx, ddx), __fmul_rn(ddy, ddy)), __fmul_rn(ddz, ddz));
                top5_insert(t, ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)__float_as_int(p.w));
            }
        }
    };
    const int len = (lane < 9) ? re - rs : 0;
    scan(lane == 4 ? len : 0);                                  // pass A: the centre row
    unsigned upper = 0x7f800000u;                               // +inf: fewer than five lanes hold a candidate, nothing is skipped
    {
        unsigned best = (unsigned)(t.k[0] >> 32);               // float bits of non-negative distances order like unsigned; empty = ~0
        unsigned u = 0u;
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            u = __reduce_min_sync(gm, best);
            if (best == u) best = 0xffffffffu;                  // ties drop several lanes at once: the bound only gets weaker
        }
        if (u < 0x7f800000u) upper = u;
    }
    scan((lane != 4 && __float_as_uint(lower) <= upper) ? len : 0);      // pass B
    // merge the 32 sorted lists: five rounds of warp arg-min on (d2, index)
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const unsigned hi = (unsigned)(t.k[0] >> 32), lo = (unsigned)t.k[0];
        const unsigned hmin = __reduce_min_sync(gm, hi);
        const unsigned lmin = __reduce_min_sync(gm, hi == hmin ? lo : 0xffffffffu);
        if (hi == hmin && lo == lmin && t.k[0] != ~0ull) {   // unique winner pops its head
            t.k[0] = t.k[1]; t.k[1] = t.k[2]; t.k[2] = t.k[3]; t.k[3] = t.k[4]; t.k[4] = ~0ull;
        }
        d2[k] = __uint_as_float(hmin);
        idx[k] = (int)lmin;
    }
    const bool ok = (idx[4] != -1) && (d2[4] < 1.0f);
    return ok;
}
// the whole warp on one query
__device__ __forceinline__ bool knn5_warp(const KnnGrid& g, float qx, float qy, float qz, int idx[5], float d2[5]) {
    return knn5_group<32>(g, qx, qy, qz, idx, d2);
}
#endif

}  // namespace pf
