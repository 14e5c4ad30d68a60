// K9 streaming map update: see merge.cuh.  Kernels (both clouds ride in the same launches, blockIdx.y = cloud):
//   k_mm_keys   : voxel key (10 + 10 + 10 bits relative to the crop-box corner, cloud in bit 31) of every B point;
//                 cropped points -> 0xffffffff (sorted last, dropped)
//   radix_sort  : stable, so equal keys keep their buffer order = the canonical summation order
//   k_mm_heads  : one thread per run of equal keys in sorted B: lower_bound in the sorted map part by recomputed keys
//                 (no key array for the map is ever stored); runs that meet a live map point are recorded as "matched",
//                 the others are reduced to finished voxels ("inserts"); both lists are compacted in key order
//   k_mm_merge  : tiles of 1024 sorted map points: CropBox, merge with the matched run, delete rule, r update, positions
//                 by a block scan over (kept map points + inserts before each slot) and a chained scan across tiles
//   k_mm_finish : appends the exceptions (centroids that left their voxel) behind the sorted part and publishes the counts
#include "merge.cuh"

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
// a centroid that left the voxel it was averaged in cannot stay in the sorted part
__device__ __forceinline__ bool left_voxel(const Pt& o, long long k64, float leaf, const Origin& g) { return key64_of(o, leaf, g) != k64; }

__device__ __forceinline__ void put_exception(const MapMergeParams& P, int cloud, const Pt& o) {
    const unsigned slot = atomicAdd(&P.state[10 + cloud], 1u);
    if ((int)slot < P.s.exc_cap) P.s.exc[(size_t)cloud * P.s.exc_cap + slot] = o;
}

// first index i in [0, n) with a[i] >= key; all 32 lanes of the warp call it with the same arguments (32-ary search)
__device__ __forceinline__ int warp_lower_bound(const int* a, int n, int key) {
    const int lane = (int)lane_id();
    int lo = 0, hi = n;
    while (hi - lo > 32) {
        const int step = (hi - lo + 31) / 32;
        const int idx = lo + (lane + 1) * step - 1;
        const bool less = idx < hi && a[idx] < key;
        const int c = __popc(__ballot_sync(0xffffffffu, less));
        const int first_ge = lo + (c + 1) * step - 1;      // probe c is the first one that is not < key (if it exists)
        const int nlo = lo + c * step;
        hi = (c < 32 && first_ge < hi) ? first_ge : hi;
        lo = nlo < hi ? nlo : hi;
    }
    const int idx = lo + lane;
    const bool less = idx < hi && a[idx] < key;
    return lo + __popc(__ballot_sync(0xffffffffu, less));
}

__global__ void __launch_bounds__(256) k_mm_keys(MapMergeParams P, uint32_t* __restrict__ keys) {
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
    const int cloud = blockIdx.y;
    const MapMergeCloud& c = P.c[cloud];
    __shared__ int s_tile;
    __shared__ int s_tmp[9];
    __shared__ unsigned s_bcast;
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
        int rA = 0, len = 0;
        Pt o{0.f, 0.f, 0.f, 0u};
        if (e < end) {
            const unsigned key = keys[e];
            if (e == s0 || keys[e - 1] != key) {
                int e2 = e + 1;
                while (e2 < end && keys[e2] == key) ++e2;
                len = e2 - e;
                const long long k64 = key64_of_key30(key);
                int lo = 0, hi = mA;
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (key64_of(c.buf[mid], c.leaf, g) < k64) lo = mid + 1; else hi = mid;
                }
                rA = lo;
                if (rA < mA) {
                    const Pt a = c.buf[rA];
                    matched = key64_of(a, c.leaf, g) == k64 && in_box(box, a);
                }
                if (!matched) {
                    VoxAcc acc;
                    for (int q = e; q < e2; q += 8) {
                        Pt v[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u)
                            if (q + u < e2) v[u] = bpts[(int)vals[q + u] - bbase];
#pragma unroll
                        for (int u = 0; u < 8; ++u)
                            if (q + u < e2) acc_add(acc, v[u]);
                    }
                    const bool keep = acc_finish(acc, P, &o);
                    if (keep) {
                        if (acc.n > 1 && left_voxel(o, k64, c.leaf, g)) put_exception(P, cloud, o);
                        else ins = true;
                    }
                }
            }
        }
        const unsigned tag = (ctrl[0] << 3);
        int total_m, total_i;
        const int lm = block_scan_excl_256(matched ? 1 : 0, s_tmp, &total_m);
        const unsigned excl_m = chained_scan_exclusive(status + (size_t)cloud * status_stride, tag | 2u, tile, (unsigned)total_m, &s_bcast);
        if (matched) {
            const int idx = lbase + (int)excl_m + lm;
            P.s.m_ra[idx] = rA; P.s.m_start[idx] = e; P.s.m_len[idx] = len;
        }
        const int li = block_scan_excl_256(ins ? 1 : 0, s_tmp, &total_i);
        const unsigned excl_i = chained_scan_exclusive(status + (size_t)(2 + cloud) * status_stride, tag | 3u, tile, (unsigned)total_i, &s_bcast);
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

__global__ void __launch_bounds__(256) k_mm_merge(MapMergeParams P, const uint32_t* __restrict__ vals, unsigned long long* status,
                                                  int status_stride, unsigned* ctrl, int ticket_word) {
    const int cloud = blockIdx.y;
    const MapMergeCloud& c = P.c[cloud];
    __shared__ int s_link[kMergeTile];          // matched head + 1 of a slot
    __shared__ int s_cnt[kMergeTile + 1];       // inserts in front of a slot
    __shared__ int s_first[kMergeTile + 1];     // first insert entry of a slot
    __shared__ int s_pre[kMergeTile + 1];       // output offset of the first insert of a slot
    __shared__ int s_rng[4];
    __shared__ int s_tile;
    __shared__ int s_tmp[9];
    __shared__ unsigned s_bcast;
    const int mA = *c.n_sorted;
    const int ntiles = mA / kMergeTile + 1;     // the last tile also takes the inserts behind the last map point
    const int nm = (int)P.state[4 + cloud], ni = (int)P.state[6 + cloud];
    const int bbase = cloud == 0 ? 0 : (int)P.state[0];
    const int lbase = cloud * P.s.cap;
    const int* m_ra = P.s.m_ra + lbase;
    const int* i_ra = P.s.i_ra + lbase;
    const CropBox box = crop_of(P.center);
    const Origin g = origin_of(box, c.leaf);
    const Pt* bpts = c.buf + mA;
    const int tid = threadIdx.x;
    while (true) {
        __syncthreads();
        if (tid == 0) s_tile = (int)atomicAdd(&ctrl[ticket_word + cloud], 1u);
        __syncthreads();
        const int tile = s_tile;
        if (tile >= ntiles) return;
        const int base = tile * kMergeTile;
        const bool last = tile == ntiles - 1;
#pragma unroll
        for (int k = 0; k < 4; ++k) { s_link[tid + 256 * k] = 0; s_cnt[tid + 256 * k] = 0; s_first[tid + 256 * k] = 0x7fffffff; }
        if (tid == 0) { s_cnt[kMergeTile] = 0; s_first[kMergeTile] = 0x7fffffff; }
        if (tid < 32) {
            const int a = warp_lower_bound(m_ra, nm, base);
            const int b = last ? nm : warp_lower_bound(m_ra, nm, base + kMergeTile);
            if (tid == 0) { s_rng[0] = a; s_rng[1] = b; }
        } else if (tid < 64) {
            const int a = warp_lower_bound(i_ra, ni, base);
            const int b = last ? ni : warp_lower_bound(i_ra, ni, base + kMergeTile);
            if (tid == 32) { s_rng[2] = a; s_rng[3] = b; }
        }
        __syncthreads();
        const int mLo = s_rng[0], mHi = s_rng[1], iLo = s_rng[2], iHi = s_rng[3];
        for (int h = mLo + tid; h < mHi; h += 256) s_link[m_ra[h] - base] = h + 1;
        for (int e = iLo + tid; e < iHi; e += 256) {
            const int s = i_ra[e] - base;
            atomicAdd(&s_cnt[s], 1);
            atomicMin(&s_first[s], e);
        }
        __syncthreads();
        // the map points of the tile: thread owns 4 consecutive slots
        Pt o[4];
        bool keep[4];
        int v = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int slot = 4 * tid + k, i = base + slot;
            keep[k] = false;
            if (i < mA) {
                const Pt p = c.buf[i];
                if (in_box(box, p)) {
                    VoxAcc acc;
                    acc_add(acc, p);
                    const int l = s_link[slot];
                    if (l) {
                        const int st = P.s.m_start[lbase + l - 1], ln = P.s.m_len[lbase + l - 1];
                        for (int q = 0; q < ln; ++q) acc_add(acc, bpts[(int)vals[st + q] - bbase]);
                    }
                    Pt r;
                    bool kp = acc_finish(acc, P, &r);
                    if (kp && acc.n > 1 && left_voxel(r, key64_of(p, c.leaf, g), c.leaf, g)) { put_exception(P, cloud, r); kp = false; }
                    keep[k] = kp;
                    o[k] = r;
                }
            }
            v += (keep[k] ? 1 : 0) + s_cnt[slot];
        }
        if (tid == 255) v += s_cnt[kMergeTile];
        int total;
        const int t_excl = block_scan_excl_256(v, s_tmp, &total);
        {
            int run = t_excl;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                s_pre[4 * tid + k] = run;
                run += s_cnt[4 * tid + k] + (keep[k] ? 1 : 0);
            }
            if (tid == 255) s_pre[kMergeTile] = run;
        }
        const unsigned tag = (ctrl[0] << 3) | 4u;
        const unsigned gbase = chained_scan_exclusive(status + (size_t)cloud * status_stride, tag, tile, (unsigned)total, &s_bcast);
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (keep[k]) c.out[gbase + s_pre[4 * tid + k] + s_cnt[4 * tid + k]] = o[k];
        for (int e = iLo + tid; e < iHi; e += 256) {
            const int s = i_ra[e] - base;
            c.out[gbase + s_pre[s] + (e - s_first[s])] = P.s.i_pt[lbase + e];
        }
        if (last && tid == 0) *c.n_sorted_out = (int)gbase + total;
    }
}

__global__ void __launch_bounds__(256) k_mm_finish(MapMergeParams P) {
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

int map_merge_scratch_create(MapMergeScratch& s, int cap_b, int exc_cap) {
    s.cap = cap_b;
    s.exc_cap = exc_cap;
    const size_t n = (size_t)2 * cap_b;
    PF_CUDA(cudaMalloc(&s.m_ra, sizeof(int) * n));
    PF_CUDA(cudaMalloc(&s.m_start, sizeof(int) * n));
    PF_CUDA(cudaMalloc(&s.m_len, sizeof(int) * n));
    PF_CUDA(cudaMalloc(&s.i_ra, sizeof(int) * n));
    PF_CUDA(cudaMalloc(&s.i_pt, sizeof(Pt) * n));
    PF_CUDA(cudaMalloc(&s.exc, sizeof(Pt) * 2 * (size_t)exc_cap));
    return PF_OK;
}

void map_merge_scratch_destroy(MapMergeScratch& s) {
    cudaFree(s.m_ra); cudaFree(s.m_start); cudaFree(s.m_len); cudaFree(s.i_ra); cudaFree(s.i_pt); cudaFree(s.exc);
    s = MapMergeScratch();
}

int map_merge(Workspace& ws, const MapMergeParams& P_in, int capB0, int capB1, int capA0, int capA1) {
    MapMergeParams P = P_in;
    P.state = ws.ctrl + kSlotBase + 3 * kSlotWords;
    const int capB = capB0 + capB1;
    PF_REQUIRE(capB <= ws.cap, "map_merge: %d unsorted points exceed workspace capacity %d", capB, ws.cap);
    PF_REQUIRE(capB0 <= P.s.cap && capB1 <= P.s.cap, "map_merge: %d / %d unsorted points exceed the scratch capacity %d", capB0, capB1, P.s.cap);
    const int capBmax = capB0 > capB1 ? capB0 : capB1, capAmax = capA0 > capA1 ? capA0 : capA1;
    int nblk = div_up(capBmax > 0 ? capBmax : 1, 256 * 4);
    if (nblk > 4 * kSMs) nblk = 4 * kSMs;
    k_mm_keys<<<dim3(nblk, 2), 256, 0, ws.stream>>>(P, ws.keys[0]);
    ws.launches += 1;
    int rb = 0;
    PF_CHECK(radix_sort(ws, reinterpret_cast<const int*>(P.state) + 8, capB > 0 ? capB : 1, 4, true, &rb));
    int tiles = div_up(capBmax > 0 ? capBmax : 1, 256);
    if (tiles > 6 * kSMs) tiles = 6 * kSMs;
    k_mm_heads<<<dim3(tiles, 2), 256, 0, ws.stream>>>(P, ws.keys[rb], ws.vals[rb], ws.scan_status, ws.status_stride, ws.ctrl, 5);
    int mtiles = capAmax / kMergeTile + 1;
    if (mtiles > 4 * kSMs) mtiles = 4 * kSMs;
    k_mm_merge<<<dim3(mtiles, 2), 256, 0, ws.stream>>>(P, ws.vals[rb], ws.scan_status, ws.status_stride, ws.ctrl, 7);
    k_mm_finish<<<2, 256, 0, ws.stream>>>(P);
    ws.launches += 3;
    PF_CUDA(cudaGetLastError());
    return PF_OK;
}

}  // namespace pf

// ------------------------------------------------------------------------------------------------------------
// stage tap (host buffers, synchronous)
// ------------------------------------------------------------------------------------------------------------
using namespace pf;

extern "C" int pf_map_merge(int device, const pf_point* sorted_map, int m_sorted, const pf_point* extra, int n_extra, const double center[3],
                            float leaf, int k_new, float theta_p, int theta_max, pf_point* out, int cap_out, int* n_out, int* n_sorted_out) {
    PF_REQUIRE(m_sorted >= 0 && n_extra >= 0 && (sorted_map || m_sorted == 0) && (extra || n_extra == 0) && center && out && n_out && n_sorted_out,
               "bad argument");
    PF_REQUIRE(leaf >= 0.2f, "leaf %g: the streaming map update needs leaf >= 0.2 m (10-bit voxel coordinates in the 200 m crop box)", leaf);
    PF_REQUIRE(cap_out >= m_sorted + n_extra, "output buffer holds %d points, need up to %d", cap_out, m_sorted + n_extra);
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
    int rc = PF_OK;
    auto body = [&]() -> int {
        PF_CHECK(workspace_create(ws, tot, stream));
        PF_CHECK(map_merge_scratch_create(sc, capb, 4096));
        PF_CUDA(cudaMalloc(&d_buf, sizeof(Pt) * tot));
        PF_CUDA(cudaMalloc(&d_out, sizeof(Pt) * tot));
        PF_CUDA(cudaMalloc(&d_counts, sizeof(int) * 8));
        PF_CUDA(cudaMalloc(&d_center, sizeof(double) * 3));
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
        PF_CHECK(workspace_begin_step(ws));
        PF_CHECK(map_merge(ws, P, capb, 0, m_sorted, 0));
        int res[4];
        unsigned err = 0;
        PF_CUDA(cudaMemcpyAsync(res, d_counts, sizeof(res), cudaMemcpyDeviceToHost, stream));
        PF_CUDA(cudaMemcpyAsync(&err, ws.ctrl + kSlotBase + 3 * kSlotWords + 15, sizeof(unsigned), cudaMemcpyDeviceToHost, stream));
        PF_CUDA(cudaStreamSynchronize(stream));
        if (err) { set_error("map merge failed (error bits %u: 2 = voxel coordinates out of range, 4 = too many exceptions)", err); return PF_ERR_CAPACITY; }
        *n_out = res[2];
        *n_sorted_out = res[3];
        if (res[2]) PF_CUDA(cudaMemcpy(out, d_out, sizeof(Pt) * res[2], cudaMemcpyDeviceToHost));
        return PF_OK;
    };
    rc = body();
    cudaStreamSynchronize(stream);
    workspace_destroy(ws);
    map_merge_scratch_destroy(sc);
    cudaFree(d_buf); cudaFree(d_out); cudaFree(d_counts); cudaFree(d_center);
    cudaStreamDestroy(stream);
    return rc;
}
