// K9 streaming map update: see merge.cuh.  Kernels (both clouds ride in the same launches, blockIdx.y = cloud):
//   k_mm_keys   : voxel key (10 + 10 + 10 bits relative to the crop-box corner, cloud in bit 31) of every B point;
//                 cropped points -> 0xffffffff (sorted last, dropped)
//   radix_sort  : stable, so equal keys keep their buffer order = the canonical summation order
//   k_mm_heads  : one thread per run of equal keys in sorted B: lower_bound in the sorted map part by recomputed keys
//                 (no key array for the map is ever stored).  A run that meets a live map point is finished right here
//                 (map point + run, delete rule, r update) and recorded as "matched" = a replacement for that map point;
//                 the others are reduced to finished voxels ("inserts"); both lists are compacted in key order
//   k_mm_tiles  : per tile of 1024 sorted map points, the first matched voxel / insert that falls into it
//   k_mm_count  : streaming pass 1: points every tile will emit (no shared memory, no barrier)
//   k_mm_write  : streaming pass 2: CropBox, delete rule, r update, replacements, inserts; output offsets from pass 1
//   k_mm_finish : appends the exceptions (centroids that left their voxel) behind the sorted part and publishes the counts
// Why two passes instead of one pass with a decoupled look-back: with ~700 tiles in flight the look-back chains every tile
// to the slowest of its predecessors; all CTAs fall into lockstep (load burst, idle, store burst) and the single-pass
// version measured 1.2-1.4 TB/s on an 8 M-point map.  Two independent streaming passes (16 B + 16 B read, 16 B written per
// point) have no inter-CTA dependency at all and measure 3.0 TB/s of algorithmic bytes (4.5 TB/s of DRAM traffic).
#include "merge.cuh"
#include <stdlib.h>

namespace pf {

namespace {

constexpr int kMergeTile = 1024;
constexpr unsigned kInvalidKey = 0xffffffffu;

struct CropBox { float lo[3], hi[3]; };

__device__ __forceinline__ CropBox crop_of(const double* c) {   // CropBox min/max as floats (:607-613)
    CropBox b;
#pragma unroll
    for (int a = 0; a < 3; ++a) { b.lo[a] = (float)(c[a] - 100); b.hi[a] = (float)(c[a] + 100); }
    return b;
}
__device__ __forceinline__ bool in_box(const CropBox& b, const Pt& p) {
    return !((p.x < b.lo[0] || p.y < b.lo[1] || p.z < b.lo[2]) || (p.x > b.hi[0] || p.y > b.hi[1] || p.z > b.hi[2]));
}
__device__ __forceinline__ int ifloor_div(float x, float leaf) {   // floor(x / leaf) as rgbds computes it (:63-65), clamped
    const float f = floorf(__fdiv_rn(x, leaf));
    return (int)fminf(fmaxf(f, -1.0e6f), 1.0e6f);
}
struct Origin { int o[3]; };
__device__ __forceinline__ Origin origin_of(const CropBox& b, float leaf) {
    Origin g;
#pragma unroll
    for (int a = 0; a < 3; ++a) g.o[a] = ifloor_div(b.lo[a], leaf);
    return g;
}
// comparable 63-bit key of any point (inside or outside the crop box): lexicographic (z, y, x) voxel coordinates
__device__ __forceinline__ long long key64_of(const Pt& p, float leaf, const Origin& g) {
    const long long dx = ifloor_div(p.x, leaf) - g.o[0] + (1 << 20), dy = ifloor_div(p.y, leaf) - g.o[1] + (1 << 20),
                    dz = ifloor_div(p.z, leaf) - g.o[2] + (1 << 20);
    return (dz << 42) | (dy << 21) | dx;
}
__device__ __forceinline__ long long key64_of_key30(unsigned k) {
    const long long dx = (k & 1023u) + (1 << 20), dy = ((k >> 10) & 1023u) + (1 << 20), dz = ((k >> 20) & 1023u) + (1 << 20);
    return (dz << 42) | (dy << 21) | dx;
}

struct VoxAcc {
    float sx = 0.f, sy = 0.f, sz = 0.f;
    int rmax = -1, gmax = -1, n = 0;
};
__device__ __forceinline__ void acc_add(VoxAcc& a, const Pt& v) {   // ordered float sums (:108-126)
    a.sx = __fadd_rn(a.sx, v.x); a.sy = __fadd_rn(a.sy, v.y); a.sz = __fadd_rn(a.sz, v.z);
    a.rmax = max(a.rmax, (int)pt_r(v.rgba));
    a.gmax = max(a.gmax, (int)pt_g(v.rgba));
    a.n += 1;
}
// centroid, delete rule (extractstablepoint :12-14) and r update (:634-646); returns keep
__device__ __forceinline__ bool acc_finish(const VoxAcc& a, const MapMergeParams& P, Pt* o) {
    const float fn = (float)a.n;
    o->x = __fdiv_rn(a.sx, fn); o->y = __fdiv_rn(a.sy, fn); o->z = __fdiv_rn(a.sz, fn);
    const bool drop = ((float)a.gmax < __fmul_rn((float)a.rmax, P.theta_p)) && (a.rmax > P.k_new) && (a.gmax < P.theta_max + 1);
    const int r2 = a.rmax > 250 ? 255 : a.rmax + 2;
    o->rgba = pack_rgba((unsigned)r2, (unsigned)a.gmax, 0u, 255u);
    return !drop;
}
// delete rule for a voxel that keeps its single map point (extractstablepoint :12-14 on the point's own counters)
__device__ __forceinline__ bool single_point_kept(const MapMergeParams& P, uint32_t rgba) {
    const int r = (int)pt_r(rgba), g = (int)pt_g(rgba);
    return !(((float)g < __fmul_rn((float)r, P.theta_p)) && (r > P.k_new) && (g < P.theta_max + 1));
}
// a centroid that left the voxel it was averaged in cannot stay in the sorted part
__device__ __forceinline__ bool left_voxel(const Pt& o, long long k64, float leaf, const Origin& g) { return key64_of(o, leaf, g) != k64; }

__device__ __forceinline__ void put_exception(const MapMergeParams& P, int cloud, const Pt& o) {
    const unsigned slot = atomicAdd(&P.state[10 + cloud], 1u);
    if ((int)slot < P.s.exc_cap) P.s.exc[(size_t)cloud * P.s.exc_cap + slot] = o;
}

__global__ void __launch_bounds__(256) k_mm_keys(MapMergeParams P, uint32_t* __restrict__ keys) {
    PF_PDL_ENTRY();
    const int cloud = blockIdx.y;
    const MapMergeCloud& c = P.c[cloud];
    const int mA = *c.n_sorted;
    const int nB = max(0, *c.n_app - mA);
    const int nB0 = max(0, *P.c[0].n_app - *P.c[0].n_sorted);
    const int base = cloud == 0 ? 0 : nB0;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        P.state[cloud] = (unsigned)nB;
        if (cloud == 0) {
            const int nB1 = max(0, *P.c[1].n_app - *P.c[1].n_sorted);
            reinterpret_cast<int*>(P.state)[8] = nB0 + nB1;
        }
    }
    const CropBox box = crop_of(P.center);
    const Origin g = origin_of(box, c.leaf);
    int valid = 0;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < nB; q += gridDim.x * blockDim.x) {
        const Pt p = c.buf[mA + q];
        unsigned key = kInvalidKey;
        if (in_box(box, p)) {
            const int dx = ifloor_div(p.x, c.leaf) - g.o[0], dy = ifloor_div(p.y, c.leaf) - g.o[1], dz = ifloor_div(p.z, c.leaf) - g.o[2];
            if ((unsigned)dx < 1024u && (unsigned)dy < 1024u && (unsigned)dz < 1024u) {
                key = (unsigned)dx | ((unsigned)dy << 10) | ((unsigned)dz << 20) | ((unsigned)cloud << 31);
                ++valid;
            } else {
                atomicOr(&P.state[15], 2u);   // leaf too small for 10-bit voxel coordinates inside the 200 m box
            }
        }
        keys[base + q] = key;
    }
    valid = __reduce_add_sync(0xffffffffu, valid);
    __shared__ int red[8];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = valid;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int k = 0; k < 8; ++k) t += red[k];
        if (t) atomicAdd(&P.state[2 + cloud], (unsigned)t);
    }
}

__global__ void __launch_bounds__(256) k_mm_heads(MapMergeParams P, const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                                                  unsigned long long* status, int status_stride, unsigned* ctrl, int ticket_word) {
    PF_PDL_ENTRY();
    const int cloud = blockIdx.y;
    const MapMergeCloud& c = P.c[cloud];
    __shared__ int s_tile;
    __shared__ int s_tmp[9];
    __shared__ unsigned s_look[kScanSmemWords];
    const int nv = (int)P.state[2 + cloud];
    const int s0 = cloud == 0 ? 0 : (int)P.state[2];
    const int end = s0 + nv;
    const int bbase = cloud == 0 ? 0 : (int)P.state[0];
    const int mA = *c.n_sorted;
    const CropBox box = crop_of(P.center);
    const Origin g = origin_of(box, c.leaf);
    const Pt* bpts = c.buf + mA;
    const int lbase = cloud * P.s.cap;
    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) s_tile = (int)atomicAdd(&ctrl[ticket_word + cloud], 1u);
        __syncthreads();
        const int tile = s_tile;
        if (nv == 0) return;     // list sizes stay 0 (state zeroed by k_begin_step)
        if (tile * 256 >= nv) return;
        const int e = s0 + tile * 256 + threadIdx.x;
        bool matched = false, ins = false;
        int rA = 0, delta = 0;
        Pt o{0.f, 0.f, 0.f, 0u};
        if (e < end) {
            const unsigned key = keys[e];
            if (e == s0 || keys[e - 1] != key) {
                int e2 = e + 1;
                while (e2 < end && keys[e2] == key) ++e2;
                const long long k64 = key64_of_key30(key);
                int lo = 0, hi = mA;
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (key64_of(c.buf[mid], c.leaf, g) < k64) lo = mid + 1; else hi = mid;
                }
                rA = lo;
                VoxAcc acc;
                bool a_kept = false;
                if (rA < mA) {
                    const Pt a = c.buf[rA];
                    matched = key64_of(a, c.leaf, g) == k64 && in_box(box, a);
                    if (matched) {      // the map point leads the voxel's sum; remember what it would have done on its own
                        acc_add(acc, a);
                        a_kept = single_point_kept(P, a.rgba);
                    }
                }
                {
                    for (int q = e; q < e2; q += 8) {
                        Pt v[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u)
                            if (q + u < e2) v[u] = bpts[(int)vals[q + u] - bbase];
#pragma unroll
                        for (int u = 0; u < 8; ++u)
                            if (q + u < e2) acc_add(acc, v[u]);
                    }
                    bool keep = acc_finish(acc, P, &o);
                    if (keep && acc.n > 1 && left_voxel(o, k64, c.leaf, g)) { put_exception(P, cloud, o); keep = false; }
                    if (matched) {      // finished voxel replaces the map point in the streaming pass; alpha = 0 marks "not kept"
                        if (!keep) o.rgba &= 0x00ffffffu;
                        delta = (keep ? 1 : 0) - (a_kept ? 1 : 0);
                    } else {
                        ins = keep;
                    }
                }
            }
        }
        const unsigned tag = (ctrl[0] << 3);
        int total_m, total_i;
        const int lm = block_scan_excl_256(matched ? 1 : 0, s_tmp, &total_m);
        const unsigned excl_m = chained_scan_exclusive(status + (size_t)cloud * status_stride, tag | 2u, tile, (unsigned)total_m, s_look);
        if (matched) {
            const int idx = lbase + (int)excl_m + lm;
            P.s.m_ra[idx] = rA; P.s.m_pt[idx] = o; P.s.m_delta[idx] = delta;
        }
        const int li = block_scan_excl_256(ins ? 1 : 0, s_tmp, &total_i);
        const unsigned excl_i = chained_scan_exclusive(status + (size_t)(2 + cloud) * status_stride, tag | 3u, tile, (unsigned)total_i, s_look);
        if (ins) {
            const int idx = lbase + (int)excl_i + li;
            P.s.i_ra[idx] = rA; P.s.i_pt[idx] = o;
        }
        if (tile == (nv - 1) / 256 && threadIdx.x == 0) {
            P.state[4 + cloud] = excl_m + (unsigned)total_m;
            P.state[6 + cloud] = excl_i + (unsigned)total_i;
        }
    }
}

// Static partition of the map tiles over the CTAs of the two streaming kernels: CTA b owns tiles [b * per, (b + 1) * per).
struct TileRange { int lo, hi, ntiles; };
__device__ __forceinline__ TileRange tile_range(int mA) {
    TileRange r;
    r.ntiles = mA / kMergeTile + 1;     // the last tile also takes the inserts behind the last map point
    // proportional split: every CTA gets floor or ceil of ntiles / grid tiles (with ceil-sized shares 8020 tiles on 2368 CTAs left
    // 363 CTAs without work and the second wave of resident CTAs a third empty)
    r.lo = (int)((long long)blockIdx.x * r.ntiles / (int)gridDim.x);
    r.hi = (int)((long long)(blockIdx.x + 1) * r.ntiles / (int)gridDim.x);
    return r;
}

// per tile of the sorted map part: first matched head / first insert at or behind the tile's first map index; zeroes the
// per-tile and per-CTA counters of the count pass
__global__ void __launch_bounds__(256) k_mm_tiles(MapMergeParams P) {
    PF_PDL_ENTRY();
    const int cloud = blockIdx.y;
    const int mA = *P.c[cloud].n_sorted;
    const int ntiles = mA / kMergeTile + 1;
    const int nm = (int)P.state[4 + cloud], ni = (int)P.state[6 + cloud];
    const int* m_ra = P.s.m_ra + cloud * P.s.cap;
    const int* i_ra = P.s.i_ra + cloud * P.s.cap;
    int* tm = P.s.tile_m + (size_t)cloud * P.s.tile_cap;
    int* ti = P.s.tile_i + (size_t)cloud * P.s.tile_cap;
    int* agg = P.s.tile_agg + (size_t)cloud * P.s.tile_cap;
    if (ntiles + 1 > P.s.tile_cap) {
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(&P.state[15], 8u);
        return;
    }
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < kMergeMaxGrid; t += gridDim.x * blockDim.x) P.s.cta_sum[cloud * kMergeMaxGrid + t] = 0;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t <= ntiles; t += gridDim.x * blockDim.x) {
        int a = nm, b = ni;
        if (t < ntiles) {
            const int key = t * kMergeTile;
            int lo = 0, hi = nm;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (m_ra[mid] < key) lo = mid + 1; else hi = mid; }
            a = lo;
            lo = 0; hi = ni;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (i_ra[mid] < key) lo = mid + 1; else hi = mid; }
            b = lo;
        }
        tm[t] = a; ti[t] = b; agg[t] = 0;
    }
}

// Streaming pass 1 (count): how many points every tile of the sorted map part will emit -- its own points that survive the
// crop box and the delete rule, corrected by the matched voxels finished in k_mm_heads, plus the inserts that fall into the
// tile.  No shared memory, no barrier: 16 B read per map point, one warp-aggregated atomic per warp and tile.
template <bool PREFETCH>
__global__ void __launch_bounds__(256, PREFETCH ? 5 : 8) k_mm_count(MapMergeParams P) {
    PF_PDL_ENTRY();
    const int cloud = blockIdx.y;
    const MapMergeCloud& c = P.c[cloud];
    const int mA = *c.n_sorted;
    const TileRange R = tile_range(mA);
    const int lbase = cloud * P.s.cap;
    const int* tm = P.s.tile_m + (size_t)cloud * P.s.tile_cap;
    const int* ti = P.s.tile_i + (size_t)cloud * P.s.tile_cap;
    int* agg = P.s.tile_agg + (size_t)cloud * P.s.tile_cap;
    const CropBox box = crop_of(P.center);
    const int tid = threadIdx.x;
    int cta_part = 0;
    const unsigned long long keep_in_l2 = l2_policy_by_code(P.hint % 10);
    auto load_tile = [&](int t, Pt (&q)[4]) {
        const int base = t * kMergeTile;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int i = base + k * 256 + tid;        // coalesced: the count does not care which thread sees which point
            q[k] = Pt{3.0e38f, 0.f, 0.f, 0u};         // far outside any crop box
            if (i < mA) *reinterpret_cast<float4*>(&q[k]) = ld_f4_l2hint(c.buf + i, keep_in_l2);
        }
    };
    Pt p[4];
    if (PREFETCH && R.lo < R.hi) load_tile(R.lo, p);
    for (int t = R.lo; t < R.hi; ++t) {
        Pt nxt[4];
        if (PREFETCH) { if (t + 1 < R.hi) load_tile(t + 1, nxt); }      // the next tile's loads are in flight while this one is counted
        else load_tile(t, p);
        int v = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) v += (in_box(box, p[k]) && single_point_kept(P, p[k].rgba)) ? 1 : 0;
        if (PREFETCH) {
#pragma unroll
            for (int k = 0; k < 4; ++k) p[k] = nxt[k];
        }
        const int mLo = tm[t], mHi = tm[t + 1];
        for (int h = mLo + tid; h < mHi; h += 256) v += P.s.m_delta[lbase + h];
        if (tid == 0) v += ti[t + 1] - ti[t];
        v = __reduce_add_sync(0xffffffffu, v);
        if ((tid & 31) == 0 && v) atomicAdd(&agg[t], v);
        cta_part += v;
    }
    if ((tid & 31) == 0 && cta_part) atomicAdd(&P.s.cta_sum[cloud * kMergeMaxGrid + blockIdx.x], cta_part);
}

// first index in [lo, hi) with a[idx] >= key (few entries: the heads / inserts that fall into one tile)
__device__ __forceinline__ int lower_bound_i(const int* __restrict__ a, int lo, int hi, int key) {
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (a[mid] < key) lo = mid + 1; else hi = mid; }
    return lo;
}

// Streaming pass 2 (write): same static partition, tiles walked in DESCENDING order so that the part of the map the count
// pass read last is read again first (what is still in the 126 MB L2).  The output offset of a tile needs no communication:
// exclusive prefix of the CTA sums + the per-tile counts of pass 1.  Inside a tile a thread owns the slots tid, tid + 256,
// ... (coalesced 16-byte loads; kept points of a warp row land on consecutive output positions = coalesced stores); the
// kept-prefix is 32 ballot words + one 32-lane scan, the few matched voxels / inserts of the tile are found by binary
// search in their sorted lists, so a tile costs two barriers and 264 bytes of shared memory.
__global__ void __launch_bounds__(256, 8) k_mm_write(MapMergeParams P) {
    const int cloud = blockIdx.y;
    const MapMergeCloud& c = P.c[cloud];
    __shared__ unsigned s_mask[2][32];          // kept ballots of the 32 (row, warp) segments, double buffered by tile parity
    __shared__ int s_segpre[2][33];             // exclusive kept-prefix of the segments, [32] = kept in the tile
    __shared__ int s_tmp[9];
    const int mA = *c.n_sorted;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int lbase = cloud * P.s.cap;
    const int ni_all = (int)P.state[6 + cloud];
    if (mA == 0) {       // first update after initMapWithPoints: everything is an insert, already in key order
        for (int e = blockIdx.x * 256 + tid; e < ni_all; e += gridDim.x * 256) c.out[e] = P.s.i_pt[lbase + e];
        if (blockIdx.x == 0 && tid == 0) *c.n_sorted_out = ni_all;
        return;
    }
    const TileRange R = tile_range(mA);
    if (R.lo >= R.hi) return;
    const int* m_ra = P.s.m_ra + lbase;
    const int* i_ra = P.s.i_ra + lbase;
    const int* tm = P.s.tile_m + (size_t)cloud * P.s.tile_cap;
    const int* ti = P.s.tile_i + (size_t)cloud * P.s.tile_cap;
    const int* agg = P.s.tile_agg + (size_t)cloud * P.s.tile_cap;
    const CropBox box = crop_of(P.center);
    PF_PDL_ENTRY();                       // results of the count pass (everything above only reads what earlier kernels left)
    // exclusive prefix of the CTA sums in front of this CTA
    int running;
    {
        int v = 0;
        for (int b = tid; b <= (int)blockIdx.x; b += 256) v += P.s.cta_sum[cloud * kMergeMaxGrid + b];
        int total;
        block_scan_excl_256(v, s_tmp, &total);
        running = total;                        // inclusive of this CTA: output end of its last tile
    }
    if (R.hi == R.ntiles && tid == 0) *c.n_sorted_out = running;
    const unsigned long long use_once = l2_policy_by_code(P.hint / 10 % 10);
    const unsigned long long out_pol = l2_policy_by_code(P.hint / 100 % 10);
    for (int tile = R.hi - 1, par = 0; tile >= R.lo; --tile, par ^= 1) {
        const int emit = agg[tile];
        const int gbase = running - emit;
        running = gbase;
        const int base = tile * kMergeTile;
        Pt p[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int i = base + k * 256 + tid;
            p[k] = Pt{0.f, 0.f, 0.f, 0u};
            if (i < mA) *reinterpret_cast<float4*>(&p[k]) = ld_f4_l2hint(c.buf + i, use_once);
        }
        const int mLo = tm[tile], mHi = tm[tile + 1], iLo = ti[tile], iHi = ti[tile + 1];
        unsigned mask[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int i = base + k * 256 + tid;
            bool keep = false;
            if (i < mA && in_box(box, p[k])) {
                int h = mHi;
                if (mHi > mLo) { h = lower_bound_i(m_ra, mLo, mHi, i); if (h < mHi && m_ra[h] != i) h = mHi; }
                if (h == mHi) {     // the voxel keeps its single point: x / 1 = x, counters from the point itself
                    keep = single_point_kept(P, p[k].rgba);
                    const unsigned r = pt_r(p[k].rgba);
                    p[k].rgba = pack_rgba(r > 250u ? 255u : r + 2u, pt_g(p[k].rgba), 0u, 255u);
                } else {            // finished in k_mm_heads
                    p[k] = P.s.m_pt[lbase + h];
                    keep = (p[k].rgba >> 24) != 0u;
                }
            }
            mask[k] = __ballot_sync(0xffffffffu, keep);
            if (lane == 0) s_mask[par][k * 8 + w] = mask[k];
        }
        __syncthreads();
        if (w == 0) {
            const int cnt = __popc(s_mask[par][lane]);
            int x = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, x, o);
                if (lane >= o) x += y;
            }
            s_segpre[par][lane] = x - cnt;
            if (lane == 31) s_segpre[par][32] = x;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (mask[k] >> lane & 1u) {
                const int i = base + k * 256 + tid;
                int pos = gbase + s_segpre[par][k * 8 + w] + __popc(mask[k] & lanemask_lt());
                if (iHi > iLo) pos += lower_bound_i(i_ra, iLo, iHi, i + 1) - iLo;      // inserts in front of this point
                st_f4_l2hint(c.out + pos, *reinterpret_cast<const float4*>(&p[k]), out_pol);
            }
        }
        const int kept = s_segpre[par][32];
        for (int e = iLo + tid; e < iHi; e += 256) {
            const int s = i_ra[e] - base;      // the insert goes in front of slot s; s >= 1024: behind the last map point
            const int before = s >= kMergeTile ? kept : s_segpre[par][s >> 5] + __popc(s_mask[par][s >> 5] & ((1u << (s & 31)) - 1u));
            c.out[gbase + before + (e - iLo)] = P.s.i_pt[lbase + e];
        }
        if (tid == 0 && kept + (iHi - iLo) != emit) atomicOr(&P.state[15], 16u);     // the two passes must agree
    }
}

__global__ void __launch_bounds__(256) k_mm_finish(MapMergeParams P) {
    PF_PDL_ENTRY();
    const int cloud = blockIdx.x;
    const MapMergeCloud& c = P.c[cloud];
    const int ns = *c.n_sorted_out;
    int ne = (int)P.state[10 + cloud];
    if (ne > P.s.exc_cap) {
        if (threadIdx.x == 0) atomicOr(&P.state[15], 4u);
        ne = P.s.exc_cap;
    }
    for (int j = threadIdx.x; j < ne; j += blockDim.x) c.out[ns + j] = P.s.exc[(size_t)cloud * P.s.exc_cap + j];
    if (threadIdx.x == 0) *c.n_out = ns + ne;
}

}  // namespace

int map_merge_scratch_create(MapMergeScratch& s, int cap_b, int exc_cap, int cap_a) {
    s.cap = cap_b;
    s.exc_cap = exc_cap;
    s.tile_cap = cap_a / kMergeTile + 3;
    PF_CUDA(cudaMalloc(&s.tile_m, sizeof(int) * 2 * (size_t)s.tile_cap));
    PF_CUDA(cudaMalloc(&s.tile_i, sizeof(int) * 2 * (size_t)s.tile_cap));
    PF_CUDA(cudaMalloc(&s.tile_agg, sizeof(int) * 2 * (size_t)s.tile_cap));
    PF_CUDA(cudaMalloc(&s.cta_sum, sizeof(int) * 2 * kMergeMaxGrid));
    const size_t n = (size_t)2 * cap_b;
    PF_CUDA(cudaMalloc(&s.m_ra, sizeof(int) * n));
    PF_CUDA(cudaMalloc(&s.m_pt, sizeof(Pt) * n));
    PF_CUDA(cudaMalloc(&s.m_delta, sizeof(int) * n));
    PF_CUDA(cudaMalloc(&s.i_ra, sizeof(int) * n));
    PF_CUDA(cudaMalloc(&s.i_pt, sizeof(Pt) * n));
    PF_CUDA(cudaMalloc(&s.exc, sizeof(Pt) * 2 * (size_t)exc_cap));
    return PF_OK;
}

void map_merge_scratch_destroy(MapMergeScratch& s) {
    cudaFree(s.m_ra); cudaFree(s.m_pt); cudaFree(s.m_delta); cudaFree(s.i_ra); cudaFree(s.i_pt); cudaFree(s.exc);
    cudaFree(s.tile_m); cudaFree(s.tile_i); cudaFree(s.tile_agg); cudaFree(s.cta_sum);
    s = MapMergeScratch();
}

int map_merge(Workspace& ws, const MapMergeParams& P_in, int capB0, int capB1, int capA0, int capA1) {
    MapMergeParams P = P_in;
    P.state = ws.ctrl + kSlotBase + 3 * kSlotWords;
    // L2 eviction priorities of the two streaming passes as decimal digits (count loads, write loads, write stores; 0 = evict_last,
    // 1 = normal, 2 = evict_first, 3 = unchanged).  Measured on an 8.2 M-point map (profiles/k9_l2_hint_sweep_r1.txt): marking
    // the count pass's lines evict_first beats keeping them for the write pass (evict_last) by 5 %
    static const int hint_env = [] { const char* e = getenv("PF_MM_HINT"); return e ? atoi(e) : 212; }();
    P.hint = hint_env;
    const int capB = capB0 + capB1;
    PF_REQUIRE(capB <= ws.cap, "map_merge: %d unsorted points exceed workspace capacity %d", capB, ws.cap);
    PF_REQUIRE(capB0 <= P.s.cap && capB1 <= P.s.cap, "map_merge: %d / %d unsorted points exceed the scratch capacity %d", capB0, capB1, P.s.cap);
    const int capBmax = capB0 > capB1 ? capB0 : capB1, capAmax = capA0 > capA1 ? capA0 : capA1;
    int nblk = div_up(capBmax > 0 ? capBmax : 1, 256 * 4);
    if (nblk > 4 * kSMs) nblk = 4 * kSMs;
    PF_CUDA(launch_pdl(k_mm_keys, dim3(nblk, 2), dim3(256), 0, ws.stream, P, ws.keys[0]));
    ws.launches += 1;
    int rb = 0;
    PF_CHECK(radix_sort(ws, reinterpret_cast<const int*>(P.state) + 8, capB > 0 ? capB : 1, 4, true, &rb));
    int tiles = div_up(capBmax > 0 ? capBmax : 1, 256);
    if (tiles > 6 * kSMs) tiles = 6 * kSMs;
    PF_CUDA(launch_pdl(k_mm_heads, dim3(tiles, 2), dim3(256), 0, ws.stream, P, ws.keys[rb], ws.vals[rb], ws.scan_status, ws.status_stride, ws.ctrl, 5));
    const int mtiles_all = capAmax / kMergeTile + 1;
    int tblk = div_up(mtiles_all + 1, 256);
    if (tblk > 2 * kSMs) tblk = 2 * kSMs;
    PF_CUDA(launch_pdl(k_mm_tiles, dim3(tblk, 2), dim3(256), 0, ws.stream, P));
    static const int grid_mult = [] { const char* e = getenv("PF_MM_GRID"); const int v = e ? atoi(e) : 16; return v < 1 ? 1 : (v > 16 ? 16 : v); }();
    int grid = mtiles_all;
    if (grid > grid_mult * kSMs) grid = grid_mult * kSMs;

    if (ws.ev_a) cudaEventRecord(ws.ev_a, ws.stream);
    {
        // count pass: 8 CTAs per SM at 32 registers without a register prefetch measured 3 % faster than 5 CTAs with one
        // (PF_MM_COUNT=0 selects the prefetching form)
        static const int count_variant = [] { const char* e = getenv("PF_MM_COUNT"); return e ? atoi(e) : 1; }();
        if (count_variant == 1) PF_CUDA(launch_pdl(k_mm_count<false>, dim3(grid, 2), dim3(256), 0, ws.stream, P));
        else PF_CUDA(launch_pdl(k_mm_count<true>, dim3(grid, 2), dim3(256), 0, ws.stream, P));
        // same grid: same tile partition.  The write pass is launched with a programmatic edge (its CTAs are scheduled while the
        // count pass drains and wait in cudaGridDependencySynchronize): saves the launch gap of the pair
        PF_CUDA(launch_pdl(k_mm_write, dim3(grid, 2), dim3(256), 0, ws.stream, P));
        ws.launches += 5;
    }
    if (ws.ev_b) cudaEventRecord(ws.ev_b, ws.stream);
    PF_CUDA(launch_pdl(k_mm_finish, dim3(2), dim3(256), 0, ws.stream, P));
    PF_CUDA(cudaGetLastError());
    return PF_OK;
}

}  // namespace pf

// ------------------------------------------------------------------------------------------------------------
// stage tap (host buffers, synchronous)
// ------------------------------------------------------------------------------------------------------------
using namespace pf;

static int map_merge_tap(int device, const pf_point* sorted_map, int m_sorted, const pf_point* extra, int n_extra, const double center[3],
                         float leaf, int k_new, float theta_p, int theta_max, pf_point* out, int cap_out, int* n_out, int* n_sorted_out,
                         int reps, float* ms_total, float* ms_stream) {
    PF_REQUIRE(m_sorted >= 0 && n_extra >= 0 && (sorted_map || m_sorted == 0) && (extra || n_extra == 0) && center && n_out && n_sorted_out,
               "bad argument");
    PF_REQUIRE(leaf >= 0.2f, "leaf %g: the streaming map update needs leaf >= 0.2 m (10-bit voxel coordinates in the 200 m crop box)", leaf);
    PF_REQUIRE(!out || cap_out >= m_sorted + n_extra, "output buffer holds %d points, need up to %d", cap_out, m_sorted + n_extra);
    PF_REQUIRE(reps >= 1, "reps must be >= 1");
    PF_CUDA(cudaSetDevice(device));
    cudaStream_t stream;
    PF_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    Workspace ws;
    MapMergeScratch sc{};
    const int tot = m_sorted + n_extra > 0 ? m_sorted + n_extra : 1;
    const int capb = n_extra > 0 ? n_extra : 1;
    Pt *d_buf = nullptr, *d_out = nullptr;
    int* d_counts = nullptr;      // [0] n_sorted, [1] n_app, [2] n_out, [3] n_sorted_out, [4..7] the empty second cloud
    double* d_center = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    int rc = PF_OK;
    auto body = [&]() -> int {
        // workspace: sort capacity for the unsorted part, look-back words for the map's merge tiles (1024 points each)
        PF_CHECK(workspace_create(ws, capb > tot / 4 ? capb : tot / 4, stream));
        PF_CHECK(map_merge_scratch_create(sc, capb, 4096, tot));
        PF_CUDA(cudaMalloc(&d_buf, sizeof(Pt) * tot));
        PF_CUDA(cudaMalloc(&d_out, sizeof(Pt) * tot));
        PF_CUDA(cudaMalloc(&d_counts, sizeof(int) * 8));
        PF_CUDA(cudaMalloc(&d_center, sizeof(double) * 3));
        for (int i = 0; i < 4; ++i) PF_CUDA(cudaEventCreate(&ev[i]));
        const int counts[8] = {m_sorted, m_sorted + n_extra, 0, 0, 0, 0, 0, 0};
        PF_CUDA(cudaMemcpyAsync(d_counts, counts, sizeof(counts), cudaMemcpyHostToDevice, stream));
        if (m_sorted) PF_CUDA(cudaMemcpyAsync(d_buf, sorted_map, sizeof(Pt) * m_sorted, cudaMemcpyHostToDevice, stream));
        if (n_extra) PF_CUDA(cudaMemcpyAsync(d_buf + m_sorted, extra, sizeof(Pt) * n_extra, cudaMemcpyHostToDevice, stream));
        PF_CUDA(cudaMemcpyAsync(d_center, center, sizeof(double) * 3, cudaMemcpyHostToDevice, stream));
        MapMergeParams P{};
        P.c[0] = MapMergeCloud{d_buf, d_counts + 0, d_counts + 1, d_out, d_counts + 2, d_counts + 3, leaf};
        P.c[1] = MapMergeCloud{d_buf, d_counts + 4, d_counts + 5, d_out, d_counts + 6, d_counts + 7, leaf};
        P.center = d_center;
        P.k_new = k_new; P.theta_p = theta_p; P.theta_max = theta_max;
        P.s = sc;
        ws.ev_a = ev[2]; ws.ev_b = ev[3];
        float tot_ms = 0.f, str_ms = 0.f;
        for (int r = 0; r < reps; ++r) {      // every repetition merges the same input into d_out
            PF_CUDA(cudaEventRecord(ev[0], stream));
            PF_CHECK(workspace_begin_step(ws));
            PF_CHECK(map_merge(ws, P, capb, 0, m_sorted, 0));
            PF_CUDA(cudaEventRecord(ev[1], stream));
            PF_CUDA(cudaStreamSynchronize(stream));
            float a = 0.f, b = 0.f;
            PF_CUDA(cudaEventElapsedTime(&a, ev[0], ev[1]));
            PF_CUDA(cudaEventElapsedTime(&b, ev[2], ev[3]));
            if (r > 0 || reps == 1) { tot_ms += a; str_ms += b; }     // repetition 0 is the warm-up when reps > 1
        }
        const int timed = reps > 1 ? reps - 1 : 1;
        if (ms_total) *ms_total = tot_ms / timed;
        if (ms_stream) *ms_stream = str_ms / timed;
        int res[4];
        unsigned err = 0;
        PF_CUDA(cudaMemcpyAsync(res, d_counts, sizeof(res), cudaMemcpyDeviceToHost, stream));
        PF_CUDA(cudaMemcpyAsync(&err, map_merge_error_word(ws), sizeof(unsigned), cudaMemcpyDeviceToHost, stream));
        PF_CUDA(cudaStreamSynchronize(stream));
        if (err) { set_error("map merge failed (error bits %u: 2 = voxel coordinates out of range, 4 = too many exceptions, 8 / 16 = internal)", err); return PF_ERR_CAPACITY; }
        *n_out = res[2];
        *n_sorted_out = res[3];
        if (out && res[2]) PF_CUDA(cudaMemcpy(out, d_out, sizeof(Pt) * res[2], cudaMemcpyDeviceToHost));
        return PF_OK;
    };
    rc = body();
    cudaStreamSynchronize(stream);
    workspace_destroy(ws);
    map_merge_scratch_destroy(sc);
    cudaFree(d_buf); cudaFree(d_out); cudaFree(d_counts); cudaFree(d_center);
    for (int i = 0; i < 4; ++i) if (ev[i]) cudaEventDestroy(ev[i]);
    cudaStreamDestroy(stream);
    return rc;
}

extern "C" int pf_map_merge(int device, const pf_point* sorted_map, int m_sorted, const pf_point* extra, int n_extra, const double center[3],
                            float leaf, int k_new, float theta_p, int theta_max, pf_point* out, int cap_out, int* n_out, int* n_sorted_out) {
    PF_REQUIRE(out, "null output");
    return map_merge_tap(device, sorted_map, m_sorted, extra, n_extra, center, leaf, k_new, theta_p, theta_max, out, cap_out, n_out, n_sorted_out,
                         1, nullptr, nullptr);
}

extern "C" int pf_map_merge_timed(int device, const pf_point* sorted_map, int m_sorted, const pf_point* extra, int n_extra,
                                  const double center[3], float leaf, int k_new, float theta_p, int theta_max, int reps, int* n_out,
                                  int* n_sorted_out, float* ms_total, float* ms_stream) {
    return map_merge_tap(device, sorted_map, m_sorted, extra, n_extra, center, leaf, k_new, theta_p, theta_max, nullptr, 0, n_out, n_sorted_out,
                         reps, ms_total, ms_stream);
}
