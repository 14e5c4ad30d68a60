// K10: global map -- device-resident replacement of LaserMappingClass
// (/root/reference/src/laserMappingClass.cpp:7-32, :110-149, :152-191, :196-208; include/laserMappingClass.h:23-30).
//
// The reference tiles space into 50 m cells (cell = floor(x / 50 + 0.5)), pushes every transformed point of the frame into
// its cell's cloud and then re-runs pcl::VoxelGrid on all 5 x 5 x 5 cells around the pose, every frame; getMap concatenates
// every cell.  VoxelGrid emits a cell's voxels in ascending (vz, vy, vx) order and getMap walks the cells in (cx, cy, cz)
// order, so the whole map is one array sorted by (cell, voxel) -- and it stays sorted from one update to the next.  One
// update is therefore the same streaming merge as the local-map update (merge.cu):
//   k_mp_append  transform (pcl::transformPointCloud, float) + intensity (:169) + cell id (:170-172) of the frame's points,
//                written behind the map in input order
//   k_mp_keys    31-bit key (cell inside the 5x5x5 block : 7 | voxel inside the cell : 3 x 8) of every unsorted point
//   radix_sort   stable -> canonical summation order (ascending input index)
//   k_mp_heads   one thread per voxel run of the sorted new points: binary search in the sorted map; a run that meets a map
//                point is finished right there (VoxelGrid centroid of map point + run: x, y, z, intensity) and becomes that
//                point's replacement, the others are reduced to finished voxels ("inserts")
//   k_mp_tiles   per tile of 1024 map points: its replacements / inserts and how many points it will emit (no pass over the map
//                is needed for that: every map point survives unless its replacement left the voxel) + per-CTA sums
//   k_mp_write   ONE streaming pass over the map, no inter-CTA dependency (output offsets = prefix of the tile counts): cells
//                outside the block pass through, replacements are substituted, inserts placed in key order
//   k_mp_finish  appends "exceptions" (centroids that float rounding pushed out of their voxel) behind the sorted part
// A point keeps the cell it was first binned into (the reference never re-bins: the cloud a centroid lives in is its cell),
// hence the 4-byte cell id beside every point: 20 B read + 20 B written per map point per update, O(map) streaming instead
// of the reference's 125 sorts + O(map) concatenation.
#include "primitives.cuh"

namespace pf {

namespace {

constexpr int kMpTile = 1024;
constexpr unsigned kBadKey = 0xffffffffu;
constexpr unsigned kNoCell = 0xffffffffu;
constexpr int kBlock = 5;            // 2 * LASER_CELL_RANGE + 1 cells per axis (include/laserMappingClass.h:29-30)
constexpr int kExcCap = 4096;
constexpr int kMpMaxGrid = 8 * kSMs;   // CTAs of the write pass

struct MapperParams {
    float4* buf;  unsigned* cbuf;     // current map: points + cell ids, [0, n_sorted) sorted | exceptions | this frame's points
    float4* out;  unsigned* cell_out; // next map
    int* counts;                      // device: [0] n_sorted [1] n_map [2] n_app [3] n_sorted_out [4] err bits [5] buffer capacity [6] map capacity
    unsigned long long* dropped;      // device: points that fell outside the 5x5x5 block (the reference's out-of-range access)
    float inv_leaf;
    int bx, by, bz;                   // lowest cell of the block around the pose
    int *m_ra, *m_keep, *i_ra;        // replacements: map index, kept (0 = the centroid left its voxel -> exception)
    float4* m_pt;                     // replacement points
    float4* i_pt;  unsigned* i_cell;
    int *tile_m, *tile_i, *tile_agg, *cta_sum;   // per 1024-point tile: first replacement / insert, points emitted; per CTA of the write pass
    int tile_cap, write_grid;
    float4* exc;   unsigned* exc_cell;
    unsigned* state;                  // [0] nB [2] nvalid [4] n matched [6] n inserts [10] n exceptions
};

__host__ __device__ __forceinline__ int cell_of(double v) { return (int)floor(v / 50.0 + 0.5); }   // :154-156, :170-172

__device__ __forceinline__ unsigned pack_cell(int cx, int cy, int cz) {
    return ((unsigned)(cx + 512) << 20) | ((unsigned)(cy + 512) << 10) | (unsigned)(cz + 512);
}
__device__ __forceinline__ void unpack_cell(unsigned c, int& cx, int& cy, int& cz) {
    cx = (int)(c >> 20) - 512; cy = (int)((c >> 10) & 1023u) - 512; cz = (int)(c & 1023u) - 512;
}
__device__ __forceinline__ bool in_block(const MapperParams& P, unsigned cell) {
    int cx, cy, cz;
    unpack_cell(cell, cx, cy, cz);
    return (unsigned)(cx - P.bx) < (unsigned)kBlock && (unsigned)(cy - P.by) < (unsigned)kBlock && (unsigned)(cz - P.bz) < (unsigned)kBlock;
}
// voxel of a point inside its cell: VoxelGrid's floor(x * inverse_leaf) relative to the first voxel that touches the cell (+1
// of slack for a centroid rounded onto the cell face), 8 bits per axis, (vz, vy, vx) = VoxelGrid's output order
__device__ __forceinline__ unsigned vox24(const float4& p, unsigned cell, float inv) {
    int c[3];
    unpack_cell(cell, c[0], c[1], c[2]);
    const float v[3] = {p.x, p.y, p.z};
    unsigned r[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const int base = (int)floor(((double)c[a] - 0.5) * 50.0 * (double)inv);
        const int q = (int)floorf(__fmul_rn(v[a], inv)) - base + 1;
        r[a] = (unsigned)min(max(q, 0), 255);
    }
    return (r[2] << 16) | (r[1] << 8) | r[0];
}
__device__ __forceinline__ unsigned long long key64(unsigned cell, unsigned v24) { return ((unsigned long long)cell << 24) | v24; }
__device__ __forceinline__ unsigned key31(const MapperParams& P, unsigned cell, unsigned v24) {
    int cx, cy, cz;
    unpack_cell(cell, cx, cy, cz);
    return ((unsigned)(((cx - P.bx) * kBlock + (cy - P.by)) * kBlock + (cz - P.bz)) << 24) | v24;
}
__device__ __forceinline__ unsigned cell_of_key31(const MapperParams& P, unsigned k) {
    const int b = (int)(k >> 24);
    return pack_cell(P.bx + b / (kBlock * kBlock), P.by + (b / kBlock) % kBlock, P.bz + b % kBlock);
}

struct Acc4 { float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f; int n = 0; };
__device__ __forceinline__ void acc_add(Acc4& a, const float4& v) {   // CentroidPoint accumulators, in order
    a.sx = __fadd_rn(a.sx, v.x); a.sy = __fadd_rn(a.sy, v.y); a.sz = __fadd_rn(a.sz, v.z); a.si = __fadd_rn(a.si, v.w);
    a.n += 1;
}
__device__ __forceinline__ float4 acc_mean(const Acc4& a) {
    const float fn = (float)a.n;
    return make_float4(__fdiv_rn(a.sx, fn), __fdiv_rn(a.sy, fn), __fdiv_rn(a.sz, fn), __fdiv_rn(a.si, fn));
}
__device__ __forceinline__ void put_exception(const MapperParams& P, const float4& o, unsigned cell) {
    const unsigned slot = atomicAdd(&P.state[10], 1u);
    if ((int)slot < kExcCap) { P.exc[slot] = o; P.exc_cell[slot] = cell; }
}

// transform + intensity + cell of the frame's points (:162-176); m = row-major 3x4 float pose (pose_current.cast<float>())
struct Mat34 { float m[12]; };
__global__ void __launch_bounds__(256) k_mp_append(MapperParams P, const float4* __restrict__ in, int n_in, Mat34 T) {
    const int n_map = P.counts[1], cap = P.counts[5];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        int tot = n_map + n_in;
        if (tot > cap) { atomicOr(&P.counts[4], 1); tot = cap; }
        P.counts[2] = tot;
    }
    int drop = 0;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n_in; q += gridDim.x * blockDim.x) {
        if (n_map + q >= cap) break;
        const float4 s = ld_stream_f4(in + q);
        float4 o;
        // detail::Transformer<float>::se3 (SSE): col0 * x + (col1 * y + (col2 * z + col3))
        o.x = __fadd_rn(__fmul_rn(T.m[0], s.x), __fadd_rn(__fmul_rn(T.m[1], s.y), __fadd_rn(__fmul_rn(T.m[2], s.z), T.m[3])));
        o.y = __fadd_rn(__fmul_rn(T.m[4], s.x), __fadd_rn(__fmul_rn(T.m[5], s.y), __fadd_rn(__fmul_rn(T.m[6], s.z), T.m[7])));
        o.z = __fadd_rn(__fmul_rn(T.m[8], s.x), __fadd_rn(__fmul_rn(T.m[9], s.y), __fadd_rn(__fmul_rn(T.m[10], s.z), T.m[11])));
        o.w = (float)fmin(1.0, fmax((double)s.z + 2.0, 0.0) / 5);                                  // :169
        const int cx = cell_of((double)o.x), cy = cell_of((double)o.y), cz = cell_of((double)o.z);
        unsigned cell = kNoCell;
        if ((unsigned)(cx - P.bx) < (unsigned)kBlock && (unsigned)(cy - P.by) < (unsigned)kBlock && (unsigned)(cz - P.bz) < (unsigned)kBlock &&
            abs(cx) < 512 && abs(cy) < 512 && abs(cz) < 512 && isfinite(o.x) && isfinite(o.y) && isfinite(o.z))
            cell = pack_cell(cx, cy, cz);
        else
            ++drop;
        P.buf[n_map + q] = o;
        P.cbuf[n_map + q] = cell;
    }
    drop = __reduce_add_sync(0xffffffffu, drop);
    if ((threadIdx.x & 31) == 0 && drop) atomicAdd(P.dropped, (unsigned long long)drop);
}

__global__ void __launch_bounds__(256) k_mp_keys(MapperParams P, uint32_t* __restrict__ keys) {
    const int mA = P.counts[0];
    const int nB = max(0, P.counts[2] - mA);
    if (blockIdx.x == 0 && threadIdx.x == 0) P.state[0] = (unsigned)nB;
    int valid = 0;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < nB; q += gridDim.x * blockDim.x) {
        const unsigned cell = P.cbuf[mA + q];
        unsigned key = kBadKey;
        if (cell != kNoCell) {
            const float4 p = P.buf[mA + q];
            if (in_block(P, cell)) { key = key31(P, cell, vox24(p, cell, P.inv_leaf)); ++valid; }
            else put_exception(P, p, cell);     // an exception of a cell that is not filtered this frame waits for its turn
        }
        keys[q] = key;
    }
    valid = __reduce_add_sync(0xffffffffu, valid);
    if ((threadIdx.x & 31) == 0 && valid) atomicAdd(&P.state[2], (unsigned)valid);
}

__global__ void __launch_bounds__(256) k_mp_heads(MapperParams P, const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                                                  unsigned long long* status, int status_stride, unsigned* ctrl) {
    __shared__ int s_tile;
    __shared__ int s_tmp[9];
    __shared__ unsigned s_look[kScanSmemWords];
    const int nv = (int)P.state[2];
    const int mA = P.counts[0];
    const float4* bpts = P.buf + mA;
    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) s_tile = (int)atomicAdd(&ctrl[1], 1u);
        __syncthreads();
        const int tile = s_tile;
        if (tile * 256 >= nv) return;
        const int e = tile * 256 + threadIdx.x;
        bool matched = false, ins = false;
        int rA = 0, len = 0;
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        unsigned cell = 0;
        if (e < nv) {
            const unsigned key = keys[e];
            if (e == 0 || keys[e - 1] != key) {
                int e2 = e + 1;
                while (e2 < nv && keys[e2] == key) ++e2;
                len = e2 - e;
                cell = cell_of_key31(P, key);
                const unsigned v24 = key & 0xffffffu;
                const unsigned long long k64 = key64(cell, v24);
                int lo = 0, hi = mA;
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    const unsigned cm = P.cbuf[mid];
                    // cells order the map; the voxel only matters inside the run's own cell
                    const bool less = cm != cell ? cm < cell : key64(cm, vox24(P.buf[mid], cm, P.inv_leaf)) < k64;
                    if (less) lo = mid + 1; else hi = mid;
                }
                rA = lo;
                if (rA < mA) {
                    const unsigned ca = P.cbuf[rA];
                    matched = ca == cell && vox24(P.buf[rA], ca, P.inv_leaf) == v24;
                }
                Acc4 acc;
                if (matched) acc_add(acc, P.buf[rA]);        // the map point leads the voxel's sum
                for (int q = e; q < e2; ++q) acc_add(acc, bpts[vals[q]]);
                o = acc_mean(acc);
                bool keep = true;
                if (acc.n > 1 && vox24(o, cell, P.inv_leaf) != v24) { put_exception(P, o, cell); keep = false; }
                if (matched) len = keep ? 1 : 0; else ins = keep;
            }
        }
        const unsigned tag = (ctrl[0] << 3);
        int total_m, total_i;
        const int lm = block_scan_excl_256(matched ? 1 : 0, s_tmp, &total_m);
        const unsigned excl_m = chained_scan_exclusive(status, tag | 2u, tile, (unsigned)total_m, s_look);
        if (matched) { const int idx = (int)excl_m + lm; P.m_ra[idx] = rA; P.m_pt[idx] = o; P.m_keep[idx] = len; }
        const int li = block_scan_excl_256(ins ? 1 : 0, s_tmp, &total_i);
        const unsigned excl_i = chained_scan_exclusive(status + status_stride, tag | 3u, tile, (unsigned)total_i, s_look);
        if (ins) { const int idx = (int)excl_i + li; P.i_ra[idx] = rA; P.i_pt[idx] = o; P.i_cell[idx] = cell; }
        if (tile == (nv - 1) / 256 && threadIdx.x == 0) {
            P.state[4] = excl_m + (unsigned)total_m;
            P.state[6] = excl_i + (unsigned)total_i;
        }
    }
}

__device__ __forceinline__ int lower_bound_i(const int* __restrict__ a, int lo, int hi, int key) {
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (a[mid] < key) lo = mid + 1; else hi = mid; }
    return lo;
}

// static partition of the map tiles over the CTAs of the write pass
__device__ __forceinline__ void tile_range(int mA, int grid, int b, int& lo, int& hi, int& ntiles) {
    ntiles = mA / kMpTile + 1;      // the last tile also takes the inserts behind the last map point
    const int per = (ntiles + grid - 1) / grid;
    lo = min(ntiles, b * per);
    hi = min(ntiles, lo + per);
}

__global__ void __launch_bounds__(256) k_mp_tiles(MapperParams P) {
    const int mA = P.counts[0];
    const int ntiles = mA / kMpTile + 1;
    const int nm = (int)P.state[4], ni = (int)P.state[6];
    if (ntiles + 1 > P.tile_cap) { if (blockIdx.x == 0 && threadIdx.x == 0) atomicOr(&P.counts[4], 4); return; }
    const int per = (ntiles + P.write_grid - 1) / P.write_grid;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t <= ntiles; t += gridDim.x * blockDim.x) {
        int a = nm, b = ni;
        if (t < ntiles) { a = lower_bound_i(P.m_ra, 0, nm, t * kMpTile); b = lower_bound_i(P.i_ra, 0, ni, t * kMpTile); }
        P.tile_m[t] = a; P.tile_i[t] = b;
    }
    // emitted points per tile: its own map points, minus the replacements that turned into exceptions, plus its inserts
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < ntiles; t += gridDim.x * blockDim.x) {
        const int base = t * kMpTile;
        const int a0 = lower_bound_i(P.m_ra, 0, nm, base), a1 = t + 1 < ntiles ? lower_bound_i(P.m_ra, 0, nm, base + kMpTile) : nm;
        const int b0 = lower_bound_i(P.i_ra, 0, ni, base), b1 = t + 1 < ntiles ? lower_bound_i(P.i_ra, 0, ni, base + kMpTile) : ni;
        int emit = min(kMpTile, max(0, mA - base)) + (b1 - b0);
        for (int h = a0; h < a1; ++h) emit -= P.m_keep[h] ? 0 : 1;
        P.tile_agg[t] = emit;
        atomicAdd(&P.cta_sum[t / per], emit);
    }
}

// The streaming pass: same structure as k_mm_write (merge.cu): a thread owns the slots tid, tid + 256, ... of a tile (coalesced
// loads and stores), kept-prefix from 32 ballot words, replacements / inserts found by binary search in their sorted lists.
__global__ void __launch_bounds__(256, 6) k_mp_write(MapperParams P) {
    __shared__ unsigned s_mask[2][32];
    __shared__ int s_segpre[2][33];
    __shared__ int s_tmp[9];
    const int mA = P.counts[0];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int ni_all = (int)P.state[6];
    if (mA == 0) {       // empty map: everything is an insert, already in key order
        for (int e = blockIdx.x * 256 + tid; e < ni_all; e += gridDim.x * 256) { P.out[e] = P.i_pt[e]; P.cell_out[e] = P.i_cell[e]; }
        if (blockIdx.x == 0 && tid == 0) P.counts[3] = ni_all;
        return;
    }
    int lo, hi, ntiles;
    tile_range(mA, (int)gridDim.x, (int)blockIdx.x, lo, hi, ntiles);
    if (lo >= hi) return;
    int running;
    {
        int v = 0;
        for (int b = tid; b <= (int)blockIdx.x; b += 256) v += P.cta_sum[b];
        int total;
        block_scan_excl_256(v, s_tmp, &total);
        running = total;
    }
    if (hi == ntiles && tid == 0) P.counts[3] = running;
    for (int tile = hi - 1, par = 0; tile >= lo; --tile, par ^= 1) {
        const int emit = P.tile_agg[tile];
        const int gbase = running - emit;
        running = gbase;
        const int base = tile * kMpTile;
        float4 p[4];
        unsigned pc[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int i = base + k * 256 + tid;
            p[k] = make_float4(0.f, 0.f, 0.f, 0.f); pc[k] = 0u;
            if (i < mA) { p[k] = ld_stream_f4(P.buf + i); pc[k] = P.cbuf[i]; }
        }
        const int mLo = P.tile_m[tile], mHi = P.tile_m[tile + 1], iLo = P.tile_i[tile], iHi = P.tile_i[tile + 1];
        unsigned mask[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int i = base + k * 256 + tid;
            bool keep = i < mA;
            if (keep && mHi > mLo) {
                const int h = lower_bound_i(P.m_ra, mLo, mHi, i);
                if (h < mHi && P.m_ra[h] == i) { p[k] = P.m_pt[h]; keep = P.m_keep[h] != 0; }    // finished in k_mp_heads
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
                if (iHi > iLo) pos += lower_bound_i(P.i_ra, iLo, iHi, i + 1) - iLo;      // inserts in front of this point
                st_stream_f4(P.out + pos, p[k]);
                P.cell_out[pos] = pc[k];
            }
        }
        const int kept = s_segpre[par][32];
        for (int e = iLo + tid; e < iHi; e += 256) {
            const int s = P.i_ra[e] - base;      // the insert goes in front of slot s; s >= 1024: behind the last map point
            const int before = s >= kMpTile ? kept : s_segpre[par][s >> 5] + __popc(s_mask[par][s >> 5] & ((1u << (s & 31)) - 1u));
            P.out[gbase + before + (e - iLo)] = P.i_pt[e];
            P.cell_out[gbase + before + (e - iLo)] = P.i_cell[e];
        }
        if (tid == 0 && kept + (iHi - iLo) != emit) atomicOr(&P.counts[4], 8);     // count and write must agree
    }
}

__global__ void __launch_bounds__(256) k_mp_finish(MapperParams P) {
    const int ns = P.counts[3];
    int ne = (int)P.state[10];
    if (ne > kExcCap) { if (threadIdx.x == 0) atomicOr(&P.counts[4], 2); ne = kExcCap; }
    for (int j = threadIdx.x; j < ne; j += blockDim.x) { P.out[ns + j] = P.exc[j]; P.cell_out[ns + j] = P.exc_cell[j]; }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (ns + ne > P.counts[6]) atomicOr(&P.counts[4], 1);
        P.counts[0] = ns; P.counts[1] = ns + ne; P.counts[2] = ns + ne;
    }
}

}  // namespace
}  // namespace pf

using namespace pf;

struct pf_mapping {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev = nullptr;
    float leaf = 0.4f;
    int mcap = 0, pcap = 0, bufcap = 0;
    Workspace ws;
    float4* d_map[2] = {nullptr, nullptr};
    unsigned* d_cell[2] = {nullptr, nullptr};
    int cur = 0;
    float4* d_in = nullptr;
    int* d_counts = nullptr;
    unsigned long long* d_dropped = nullptr;
    int *d_mra = nullptr, *d_mkeep = nullptr, *d_ira = nullptr;
    float4 *d_mpt = nullptr, *d_ipt = nullptr, *d_exc = nullptr;
    int *d_tile = nullptr, *d_cta = nullptr;     // [3][tile_cap] tile_m, tile_i, tile_agg; [kMpMaxGrid] per-CTA sums
    int tile_cap = 0;
    unsigned *d_icell = nullptr, *d_exccell = nullptr;
    int* h_counts = nullptr;              // pinned read-back of the device counts after the last update
    unsigned long long* h_dropped = nullptr;
    bool pending = false;                 // a read-back is in flight
    int map_ub = 0, sorted_ub = 0;        // upper bounds for launch geometry
    uint64_t launches = 0;
};

namespace {

int mapping_settle(pf_mapping* h) {   // wait for the last update's counts
    if (h->pending) {
        PF_CUDA(cudaEventSynchronize(h->ev));
        h->pending = false;
        if (h->h_counts[4] & 1) { set_error("global map exceeded max_map_points = %d", h->mcap); return PF_ERR_CAPACITY; }
        if (h->h_counts[4] & 2) { set_error("more than %d centroids left their voxel in one update", kExcCap); return PF_ERR_CAPACITY; }
        if (h->h_counts[4] & 12) { set_error("global map update: internal error (bits %d)", h->h_counts[4] & 12); return PF_ERR_CUDA; }
        h->map_ub = h->h_counts[1];
        h->sorted_ub = h->h_counts[0];
    }
    return PF_OK;
}

}  // namespace

// LaserMappingClass::init, src/laserMappingClass.cpp:7-32 (the 5x5x5 block of empty cells needs no storage here)
extern "C" int pf_mapping_create(double map_resolution, int max_map_points, int max_points, int device, pf_mapping** out) {
    PF_REQUIRE(out, "null argument");
    PF_REQUIRE(map_resolution >= 0.2, "map_resolution %g: voxel coordinates inside a 50 m cell are kept in 8 bits (needs >= 0.2 m)", map_resolution);
    int ndev = 0;
    PF_CUDA(cudaGetDeviceCount(&ndev));
    PF_REQUIRE(device >= 0 && device < ndev, "device %d not available (%d devices)", device, ndev);
    PF_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    PF_CUDA(cudaGetDeviceProperties(&prop, device));
    PF_REQUIRE(prop.major == 10, "pfilter_b200 needs an sm_100a device, found sm_%d%d", prop.major, prop.minor);
    pf_mapping* h = new pf_mapping();
    h->device = device;
    h->leaf = (float)map_resolution;       // setLeafSize takes floats (:31)
    h->mcap = max_map_points > 0 ? max_map_points : (16 << 20);
    h->pcap = max_points > 0 ? max_points : 262144;
    h->bufcap = h->mcap + h->pcap;
    auto body = [&]() -> int {
        PF_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        PF_CUDA(cudaEventCreateWithFlags(&h->ev, cudaEventDisableTiming));
        const int capb = h->pcap + kExcCap;
        PF_CHECK(workspace_create(h->ws, capb > h->mcap / 4 ? capb : h->mcap / 4, h->stream));   // status words cover mcap / 1024 merge tiles
        for (int b = 0; b < 2; ++b) {
            PF_CUDA(cudaMalloc(&h->d_map[b], sizeof(float4) * (size_t)h->bufcap));
            PF_CUDA(cudaMalloc(&h->d_cell[b], sizeof(unsigned) * (size_t)h->bufcap));
        }
        PF_CUDA(cudaMalloc(&h->d_in, sizeof(float4) * (size_t)h->pcap));
        PF_CUDA(cudaMalloc(&h->d_counts, sizeof(int) * 8));
        PF_CUDA(cudaMalloc(&h->d_dropped, sizeof(unsigned long long)));
        PF_CUDA(cudaMemset(h->d_dropped, 0, sizeof(unsigned long long)));
        const int counts[8] = {0, 0, 0, 0, 0, h->bufcap, h->mcap, 0};
        PF_CUDA(cudaMemcpy(h->d_counts, counts, sizeof(counts), cudaMemcpyHostToDevice));
        PF_CUDA(cudaMalloc(&h->d_mra, sizeof(int) * capb));
        PF_CUDA(cudaMalloc(&h->d_mkeep, sizeof(int) * capb));
        PF_CUDA(cudaMalloc(&h->d_mpt, sizeof(float4) * capb));
        h->tile_cap = h->bufcap / kMpTile + 3;
        PF_CUDA(cudaMalloc(&h->d_tile, sizeof(int) * 3 * (size_t)h->tile_cap));
        PF_CUDA(cudaMalloc(&h->d_cta, sizeof(int) * kMpMaxGrid));
        PF_CUDA(cudaMalloc(&h->d_ira, sizeof(int) * capb));
        PF_CUDA(cudaMalloc(&h->d_ipt, sizeof(float4) * capb));
        PF_CUDA(cudaMalloc(&h->d_icell, sizeof(unsigned) * capb));
        PF_CUDA(cudaMalloc(&h->d_exc, sizeof(float4) * kExcCap));
        PF_CUDA(cudaMalloc(&h->d_exccell, sizeof(unsigned) * kExcCap));
        PF_CUDA(cudaMallocHost(&h->h_counts, sizeof(int) * 8));
        PF_CUDA(cudaMallocHost(&h->h_dropped, sizeof(unsigned long long)));
        memset(h->h_counts, 0, sizeof(int) * 8);
        *h->h_dropped = 0;
        return PF_OK;
    };
    const int rc = body();
    if (rc != PF_OK) { pf_mapping_destroy(h); return rc; }
    *out = h;
    return PF_OK;
}

extern "C" int pf_mapping_destroy(pf_mapping* h) {
    if (!h) return PF_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    workspace_destroy(h->ws);
    for (int b = 0; b < 2; ++b) { cudaFree(h->d_map[b]); cudaFree(h->d_cell[b]); }
    cudaFree(h->d_in); cudaFree(h->d_counts); cudaFree(h->d_dropped);
    cudaFree(h->d_mra); cudaFree(h->d_mkeep); cudaFree(h->d_mpt); cudaFree(h->d_ira); cudaFree(h->d_ipt); cudaFree(h->d_icell);
    cudaFree(h->d_tile); cudaFree(h->d_cta);
    cudaFree(h->d_exc); cudaFree(h->d_exccell);
    cudaFreeHost(h->h_counts); cudaFreeHost(h->h_dropped);
    if (h->ev) cudaEventDestroy(h->ev);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return PF_OK;
}

// LaserMappingClass::updateCurrentPointsToMap, src/laserMappingClass.cpp:152-191.  rt = row-major 3x4 [R | t] (double) of
// the Eigen::Isometry3d pose_current.  Asynchronous: returns once the work is enqueued; the next call (or get_map) waits.
static int mapping_update(pf_mapping* h, const float* xyzi, int n, const double rt[12], int device_input) {
    PF_REQUIRE(h && rt && (xyzi || n == 0), "null argument");
    PF_REQUIRE(n >= 0 && n <= h->pcap, "cloud of %d points exceeds max_points %d", n, h->pcap);
    PF_CUDA(cudaSetDevice(h->device));
    PF_CHECK(mapping_settle(h));
    const float4* src = reinterpret_cast<const float4*>(xyzi);
    if (!device_input && n > 0) {
        PF_CUDA(cudaMemcpyAsync(h->d_in, xyzi, sizeof(float4) * (size_t)n, cudaMemcpyHostToDevice, h->stream));
        src = h->d_in;
    }
    Workspace& ws = h->ws;
    MapperParams P{};
    P.buf = h->d_map[h->cur]; P.cbuf = h->d_cell[h->cur];
    P.out = h->d_map[h->cur ^ 1]; P.cell_out = h->d_cell[h->cur ^ 1];
    P.counts = h->d_counts; P.dropped = h->d_dropped;
    P.inv_leaf = 1.0f / h->leaf;       // inverse_leaf_size_ = 1 / leaf_size_ (float)
    P.bx = cell_of(rt[3]) - 2; P.by = cell_of(rt[7]) - 2; P.bz = cell_of(rt[11]) - 2;   // :154-156, block of checkPoints :110-149
    P.m_ra = h->d_mra; P.m_keep = h->d_mkeep; P.m_pt = h->d_mpt; P.i_ra = h->d_ira; P.i_pt = h->d_ipt; P.i_cell = h->d_icell;
    P.tile_m = h->d_tile; P.tile_i = h->d_tile + h->tile_cap; P.tile_agg = h->d_tile + 2 * h->tile_cap; P.cta_sum = h->d_cta;
    P.tile_cap = h->tile_cap;
    P.exc = h->d_exc; P.exc_cell = h->d_exccell;
    P.state = ws.ctrl + kSlotBase;
    Mat34 T;
    for (int i = 0; i < 12; ++i) T.m[i] = (float)rt[i];     // pose_current.cast<float>() (:162)
    PF_CHECK(workspace_begin_step(ws));
    int nblk = div_up(n > 0 ? n : 1, 256 * 2);
    if (nblk > 4 * kSMs) nblk = 4 * kSMs;
    k_mp_append<<<nblk, 256, 0, h->stream>>>(P, src, n, T);
    const int capB = (h->map_ub - h->sorted_ub) + n > 0 ? (h->map_ub - h->sorted_ub) + n : 1;
    PF_REQUIRE(capB <= h->pcap + kExcCap, "internal: %d unsorted points exceed the scratch capacity", capB);
    int kblk = div_up(capB, 256 * 4);
    if (kblk > 4 * kSMs) kblk = 4 * kSMs;
    k_mp_keys<<<kblk, 256, 0, h->stream>>>(P, ws.keys[0]);
    ws.launches += 2;
    int rb = 0;
    PF_CHECK(radix_sort(ws, reinterpret_cast<const int*>(P.state), capB, 4, true, &rb));
    int tiles = div_up(capB, 256);
    if (tiles > 6 * kSMs) tiles = 6 * kSMs;
    k_mp_heads<<<tiles, 256, 0, h->stream>>>(P, ws.keys[rb], ws.vals[rb], ws.scan_status, ws.status_stride, ws.ctrl);
    int wgrid = h->map_ub / kMpTile + 1;
    if (wgrid > kMpMaxGrid) wgrid = kMpMaxGrid;
    P.write_grid = wgrid;
    PF_CUDA(cudaMemsetAsync(h->d_cta, 0, sizeof(int) * kMpMaxGrid, h->stream));
    int tblk = div_up(h->map_ub / kMpTile + 2, 256);
    if (tblk > 2 * kSMs) tblk = 2 * kSMs;
    k_mp_tiles<<<tblk, 256, 0, h->stream>>>(P);
    k_mp_write<<<wgrid, 256, 0, h->stream>>>(P);
    k_mp_finish<<<1, 256, 0, h->stream>>>(P);
    ws.launches += 4;
    PF_CUDA(cudaGetLastError());
    h->cur ^= 1;
    PF_CUDA(cudaMemcpyAsync(h->h_counts, h->d_counts, sizeof(int) * 8, cudaMemcpyDeviceToHost, h->stream));
    PF_CUDA(cudaMemcpyAsync(h->h_dropped, h->d_dropped, sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
    PF_CUDA(cudaEventRecord(h->ev, h->stream));
    h->pending = true;
    h->launches = ws.launches;
    return PF_OK;
}

extern "C" int pf_mapping_update(pf_mapping* h, const float* xyzi, int n, const double rt[12]) { return mapping_update(h, xyzi, n, rt, 0); }
extern "C" int pf_mapping_update_device(pf_mapping* h, const void* d_xyzi, int n, const double rt[12]) {
    return mapping_update(h, (const float*)d_xyzi, n, rt, 1);
}

extern "C" int pf_mapping_map_size(pf_mapping* h, int* n) {
    PF_REQUIRE(h && n, "null argument");
    PF_CUDA(cudaSetDevice(h->device));
    PF_CHECK(mapping_settle(h));
    *n = h->map_ub;
    return PF_OK;
}

// LaserMappingClass::getMap, src/laserMappingClass.cpp:196-208: all cells in (x, y, z) cell order, each in VoxelGrid order
extern "C" int pf_mapping_get_map(pf_mapping* h, float* xyzi_out, int cap, int* n) {
    PF_REQUIRE(h && xyzi_out && n, "null argument");
    PF_CUDA(cudaSetDevice(h->device));
    PF_CHECK(mapping_settle(h));
    PF_REQUIRE(h->map_ub <= cap, "map has %d points, buffer holds %d", h->map_ub, cap);
    if (h->map_ub) PF_CUDA(cudaMemcpyAsync(xyzi_out, h->d_map[h->cur], sizeof(float4) * (size_t)h->map_ub, cudaMemcpyDeviceToHost, h->stream));
    PF_CUDA(cudaStreamSynchronize(h->stream));
    *n = h->map_ub;
    return PF_OK;
}

extern "C" int pf_mapping_stats(pf_mapping* h, int* n_sorted, long long* dropped, uint64_t* launches) {
    PF_REQUIRE(h, "null argument");
    PF_CUDA(cudaSetDevice(h->device));
    PF_CHECK(mapping_settle(h));
    if (n_sorted) *n_sorted = h->sorted_ub;
    if (dropped) *dropped = (long long)*h->h_dropped;
    if (launches) *launches = h->launches;
    return PF_OK;
}

extern "C" void* pf_mapping_stream(pf_mapping* h) { return h ? (void*)h->stream : nullptr; }
