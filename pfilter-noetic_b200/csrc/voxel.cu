// K2 / K9: voxel-grid kernels.
//   mode VOX_PCL : pcl::VoxelGrid<PointXYZRGB>::filter as called by downSamplingToMap
//                  (/root/reference/src/odomEstimationClass.cpp:176-180, :244-245); PCL 1.10 semantics restated in
//                  key = floor(x * (1/leaf)) - min_b, centroid = float sum / n, rgba = float mean.
//   mode VOX_MAP : the map maintenance of addPointsToMap (:606-647): CropBox [t-100, t+100] (inclusive, float compares),
//                  rgbds (:34-134: key = floor(x / leaf) - min_b, xyz mean, r = max r, g = max g, b = 0, a = 255),
//                  extractstablepoint (:7-25) and the saturating r += 2 (:634-646), fused into the segment reduction.
// Pipeline per call (both clouds -- edge and surf -- ride in the same launches, blockIdx.y = cloud):
//   k_vox_bounds  min/max of the kept points            (reads 16 B/point)
//   k_vox_keys    voxel key per point, cloud bit 31, cropped points -> 0xffffffff
//   radix_sort    stable: equal keys stay in ascending input index = the canonical summation order
//   k_vox_reduce  one thread per segment head sums its voxel in order; keep flags are compacted in key order by a
//                 chained scan; output written once (16 B/voxel).
#include "primitives.cuh"
#include "voxel.cuh"

namespace pf {

// order-preserving float <-> uint key
__device__ __forceinline__ unsigned f2ord(float f) {
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned k) {
    unsigned u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(u);
}

__device__ __forceinline__ Pt load_pt(const VoxCloud& c, int i) {
    Pt p = c.in[i];
    if (c.in_is_xyzi) p.rgba = pack_rgba(0, 0, 0, 255);   // pcl::copyPointCloud XYZI -> XYZRGB (src/odomEstimationNode copy.cpp:74-80)
    return p;
}

__device__ __forceinline__ bool in_crop(const VoxParams& P, const Pt& p) {
    if (P.mode != VOX_MAP) return true;
    const double* c = P.center;
    const float lo0 = (float)(c[0] - 100), lo1 = (float)(c[1] - 100), lo2 = (float)(c[2] - 100);
    const float hi0 = (float)(c[0] + 100), hi1 = (float)(c[1] + 100), hi2 = (float)(c[2] + 100);
    return !((p.x < lo0 || p.y < lo1 || p.z < lo2) || (p.x > hi0 || p.y > hi1 || p.z > hi2));
}

// state slot layout (unsigned words): [0..5] cloud0 {~min xyz, max xyz}, [6..11] cloud1, [12..13] nvalid, [14] n_total, [15] err
__global__ void __launch_bounds__(256) k_vox_bounds(VoxParams P) {
    PF_PDL_ENTRY();
    const VoxCloud& c = P.c[blockIdx.y];
    const int n = c.n_in ? *c.n_in : 0;
    unsigned mn[3] = {0u, 0u, 0u}, mx[3] = {0u, 0u, 0u};   // mn holds ~ord(min): both reduce with max
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        Pt p = load_pt(c, i);
        if (!in_crop(P, p)) continue;
        unsigned kx = f2ord(p.x), ky = f2ord(p.y), kz = f2ord(p.z);
        mn[0] = max(mn[0], ~kx); mn[1] = max(mn[1], ~ky); mn[2] = max(mn[2], ~kz);
        mx[0] = max(mx[0], kx); mx[1] = max(mx[1], ky); mx[2] = max(mx[2], kz);
    }
    __shared__ unsigned red[8][6];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        mn[a] = __reduce_max_sync(0xffffffffu, mn[a]);
        mx[a] = __reduce_max_sync(0xffffffffu, mx[a]);
    }
    if (lane == 0) for (int a = 0; a < 3; ++a) { red[w][a] = mn[a]; red[w][3 + a] = mx[a]; }
    __syncthreads();
    if (threadIdx.x < 6) {
        unsigned v = 0;
        for (int k = 0; k < 8; ++k) v = max(v, red[k][threadIdx.x]);
        if (v) atomicMax(&P.state[blockIdx.y * 6 + threadIdx.x], v);
    }
}

struct VoxGrid { int minb[3]; int div[3]; bool empty; };

__device__ __forceinline__ VoxGrid vox_grid(const VoxParams& P, int cloud, float leaf, float inv, bool* overflow) {
    VoxGrid g;
    const unsigned* s = P.state + cloud * 6;
    g.empty = (s[3] == 0u);
    *overflow = false;
    if (g.empty) { for (int a = 0; a < 3; ++a) { g.minb[a] = 0; g.div[a] = 1; } return g; }
    float mn[3], mx[3];
    for (int a = 0; a < 3; ++a) { mn[a] = ord2f(~s[a]); mx[a] = ord2f(s[3 + a]); }
    if (P.mode == VOX_PCL) {
        long long d[3];
        for (int a = 0; a < 3; ++a) {
            d[a] = (long long)(__fmul_rn(__fsub_rn(mx[a], mn[a]), inv)) + 1;
            g.minb[a] = (int)floorf(__fmul_rn(mn[a], inv));
            g.div[a] = (int)floorf(__fmul_rn(mx[a], inv)) - g.minb[a] + 1;
        }
        if (d[0] * d[1] * d[2] > 2147483647ll) *overflow = true;   // PCL: "Leaf size is too small for the input dataset"
    } else {
        for (int a = 0; a < 3; ++a) {
            g.minb[a] = (int)floorf(__fdiv_rn(mn[a], leaf));
            g.div[a] = (int)floorf(__fdiv_rn(mx[a], leaf)) - g.minb[a] + 1;
        }
    }
    return g;
}

__global__ void __launch_bounds__(256) k_vox_keys(VoxParams P, uint32_t* __restrict__ keys) {
    PF_PDL_ENTRY();
    const int cloud = blockIdx.y;
    const VoxCloud& c = P.c[cloud];
    const int n = c.n_in ? *c.n_in : 0;
    const int n0 = P.c[0].n_in ? *P.c[0].n_in : 0;
    const int base = cloud == 0 ? 0 : n0;
    if (blockIdx.x == 0 && cloud == 0 && threadIdx.x == 0) {
        const int n1 = P.c[1].n_in ? *P.c[1].n_in : 0;
        reinterpret_cast<int*>(P.state)[14] = n0 + n1;
    }
    const float leaf = c.leaf, inv = __fdiv_rn(1.0f, leaf);
    bool overflow;
    const VoxGrid g = vox_grid(P, cloud, leaf, inv, &overflow);
    if (overflow && blockIdx.x == 0 && threadIdx.x == 0) atomicOr(&P.state[15], 1u);
    int valid = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        Pt p = load_pt(c, i);
        unsigned key = 0xffffffffu;
        if (in_crop(P, p) && !overflow) {
            int i0, i1, i2;
            if (P.mode == VOX_PCL) {
                i0 = (int)(floorf(__fmul_rn(p.x, inv)) - (float)g.minb[0]);
                i1 = (int)(floorf(__fmul_rn(p.y, inv)) - (float)g.minb[1]);
                i2 = (int)(floorf(__fmul_rn(p.z, inv)) - (float)g.minb[2]);
            } else {
                i0 = (int)(floorf(__fdiv_rn(p.x, leaf)) - (float)g.minb[0]);
                i1 = (int)(floorf(__fdiv_rn(p.y, leaf)) - (float)g.minb[1]);
                i2 = (int)(floorf(__fdiv_rn(p.z, leaf)) - (float)g.minb[2]);
            }
            int idx = i0 + i1 * g.div[0] + i2 * g.div[0] * g.div[1];
            key = ((unsigned)idx & 0x7fffffffu) | ((unsigned)cloud << 31);
            ++valid;
        }
        keys[base + i] = key;
    }
    valid = __reduce_add_sync(0xffffffffu, valid);
    __shared__ int red[8];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = valid;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int k = 0; k < 8; ++k) t += red[k];
        if (t) atomicAdd(reinterpret_cast<int*>(P.state) + 12 + cloud, t);
    }
}

struct VoxSum {
    float sx = 0.f, sy = 0.f, sz = 0.f, sr = 0.f, sg = 0.f, sb = 0.f, sa = 0.f;
    int rmax = -1, gmax = -1;
};
__device__ __forceinline__ void vox_add(VoxSum& a, bool pcl, float x, float y, float z, unsigned rgba) {
    a.sx = __fadd_rn(a.sx, x); a.sy = __fadd_rn(a.sy, y); a.sz = __fadd_rn(a.sz, z);
    if (pcl) {
        a.sr += (float)(rgba & 0xff); a.sg += (float)((rgba >> 8) & 0xff);
        a.sb += (float)((rgba >> 16) & 0xff); a.sa += (float)(rgba >> 24);
    } else {
        a.rmax = max(a.rmax, (int)pt_r(rgba));
        a.gmax = max(a.gmax, (int)pt_g(rgba));
    }
}
__device__ __forceinline__ bool vox_finish(const VoxSum& a, const VoxParams& P, bool pcl, int n, Pt* o) {
    const float fn = (float)n;
    o->x = __fdiv_rn(a.sx, fn); o->y = __fdiv_rn(a.sy, fn); o->z = __fdiv_rn(a.sz, fn);
    if (pcl) {
        o->rgba = pack_rgba((unsigned)__fdiv_rn(a.sr, fn) & 0xff, (unsigned)__fdiv_rn(a.sg, fn) & 0xff,
                            (unsigned)__fdiv_rn(a.sb, fn) & 0xff, (unsigned)__fdiv_rn(a.sa, fn) & 0xff);
        return true;
    }
    // extractstablepoint (:12-14): drop if g < r*theta_p && r > k_new && g < theta_max + 1
    const bool drop = ((float)a.gmax < __fmul_rn((float)a.rmax, P.theta_p)) && (a.rmax > P.k_new) && (a.gmax < P.theta_max + 1);
    const int r2 = a.rmax > 250 ? 255 : a.rmax + 2;   // :634-646
    o->rgba = pack_rgba((unsigned)r2, (unsigned)a.gmax, 0u, 255u);
    return !drop;
}

// One CTA per tile of 256 sorted entries (persistent, tiles by ticket).  The runs of equal keys that START in the tile are
// its voxels.  The ordered float sum of a voxel (ascending input index) is inherently sequential; what matters is how many
// dependent gather latencies it costs.  Short runs (<= 16 points: the bulk of the voxels) are summed by one thread each, all
// in parallel; long runs (near-range ground voxels of a 64-ring scan hold > 250 points) are taken by a warp, which gathers 32
// points per step with the next step's gather already in flight.  Finished voxels are compacted in key order by a chained
// scan; the output is written once, coalesced.
constexpr int kVoxShortRun = 16;
__global__ void __launch_bounds__(256) k_vox_reduce(VoxParams P, const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                                                    unsigned long long* status, int status_stride, unsigned* ctrl, int ticket_word, int site) {
    PF_PDL_ENTRY();
    const int cloud = blockIdx.y;
    const VoxCloud& c = P.c[cloud];
    __shared__ int s_tile, s_nlong;
    __shared__ int s_tmp[9];
    __shared__ unsigned s_look[kScanSmemWords];
    __shared__ int s_head[256];          // sorted position of the tile's k-th voxel
    __shared__ int s_long[256];          // voxels left to the warps
    __shared__ Pt s_out[256];            // finished points
    __shared__ uint8_t s_keep[256];
    const int* st = reinterpret_cast<const int*>(P.state);
    const int start = cloud == 0 ? 0 : st[12];
    const int len = st[12 + cloud];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const bool pcl = P.mode == VOX_PCL;
    while (true) {   // persistent CTAs pull tiles by ticket until the cloud is exhausted
        __syncthreads();
        if (tid == 0) { s_tile = (int)atomicAdd(&ctrl[ticket_word + cloud], 1u); s_nlong = 0; }
        __syncthreads();
        const int tile = s_tile;
        if (len == 0) {
            if (tile == 0 && tid == 0) *c.n_out = 0;
            return;
        }
        if (tile * 256 >= len) return;
        const int n0 = P.c[0].n_in ? *P.c[0].n_in : 0;
        const int ibase = cloud == 0 ? 0 : n0;
        const int end = start + len;
        const int tile_end = min(end, start + (tile + 1) * 256);
        const int p = start + tile * 256 + tid;
        const bool head = p < end && (p == start || keys[p - 1] != keys[p]);
        int nheads;
        const int hidx = block_scan_excl_256(head ? 1 : 0, s_tmp, &nheads);
        if (head) s_head[hidx] = p;
        __syncthreads();
        if (tid < nheads) {
            const int ps = s_head[tid];
            const int e = tid + 1 < nheads ? s_head[tid + 1] : -1;      // the last voxel may run on into the following tiles
            if (e >= 0 && e - ps <= kVoxShortRun) {
                VoxSum a;
                for (int q = ps; q < e; q += 8) {
                    Pt v[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u)
                        if (q + u < e) v[u] = load_pt(c, (int)vals[q + u] - ibase);
#pragma unroll
                    for (int u = 0; u < 8; ++u)
                        if (q + u < e) vox_add(a, pcl, v[u].x, v[u].y, v[u].z, v[u].rgba);
                }
                Pt o;
                s_keep[tid] = vox_finish(a, P, pcl, e - ps, &o) ? 1 : 0;
                s_out[tid] = o;
            } else {
                s_long[atomicAdd(&s_nlong, 1)] = tid;
            }
        }
        __syncthreads();
        const int nlong = s_nlong;
        for (int j = w; j < nlong; j += 8) {
            const int hh = s_long[j];
            const int ps = s_head[hh];
            int e;
            if (hh + 1 < nheads) {
                e = s_head[hh + 1];
            } else {
                const unsigned key = keys[ps];
                e = tile_end;
                while (e < end) {
                    const bool same = e + lane < end && keys[e + lane] == key;
                    const unsigned m = __ballot_sync(0xffffffffu, same);
                    if (m != 0xffffffffu) { e += __ffs(~m) - 1; break; }
                    e += 32;
                }
            }
            VoxSum a;
            Pt cur{0.f, 0.f, 0.f, 0u};
            if (ps + lane < e) cur = load_pt(c, (int)vals[ps + lane] - ibase);
            for (int q = ps; q < e; q += 32) {
                Pt nxt{0.f, 0.f, 0.f, 0u};
                if (q + 32 + lane < e) nxt = load_pt(c, (int)vals[q + 32 + lane] - ibase);
                const int cnt = min(32, e - q);
                for (int u = 0; u < cnt; ++u)
                    vox_add(a, pcl, __shfl_sync(0xffffffffu, cur.x, u), __shfl_sync(0xffffffffu, cur.y, u), __shfl_sync(0xffffffffu, cur.z, u),
                            __shfl_sync(0xffffffffu, cur.rgba, u));
                cur = nxt;
            }
            if (lane == 0) {
                Pt o;
                s_keep[hh] = vox_finish(a, P, pcl, e - ps, &o) ? 1 : 0;
                s_out[hh] = o;
            }
        }
        __syncthreads();
        const bool keep = tid < nheads && s_keep[tid];
        int total;
        const int local = block_scan_excl_256(keep ? 1 : 0, s_tmp, &total);
        const unsigned tag = (ctrl[0] << 3) | (unsigned)site;
        const unsigned excl = chained_scan_exclusive(status + (size_t)cloud * status_stride, tag, tile, (unsigned)total, s_look);
        if (keep) c.out[excl + local] = s_out[tid];
        if (tile == (len - 1) / 256 && tid == 0) *c.n_out = (int)excl + total;
    }
}

int voxelize(Workspace& ws, const VoxParams& P_in, int slot, int cap0, int cap1) {
    VoxParams P = P_in;
    P.state = ws.ctrl + kSlotBase + slot * kSlotWords;
    const int cap = cap0 + cap1;
    PF_REQUIRE(cap <= ws.cap, "voxelize: %d points exceed workspace capacity %d", cap, ws.cap);
    const int capmax = cap0 > cap1 ? cap0 : cap1;
    int nblk = div_up(capmax, 256 * 4);
    if (nblk > 4 * kSMs) nblk = 4 * kSMs;
    if (nblk < 1) nblk = 1;
    PF_CUDA(launch_pdl(k_vox_bounds, dim3(nblk, 2), dim3(256), 0, ws.stream, P));
    PF_CUDA(launch_pdl(k_vox_keys, dim3(nblk, 2), dim3(256), 0, ws.stream, P, ws.keys[0]));
    ws.launches += 2;
    const int* n_total = reinterpret_cast<const int*>(P.state) + 14;
    int rb = 0;
    PF_CHECK(radix_sort(ws, n_total, cap, 4, true, &rb));
    int tiles = div_up(capmax, 256);
    if (tiles > 6 * kSMs) tiles = 6 * kSMs;
    PF_CUDA(launch_pdl(k_vox_reduce, dim3(tiles < 1 ? 1 : tiles, 2), dim3(256), 0, ws.stream, P, ws.keys[rb], ws.vals[rb], ws.scan_status,
                       ws.status_stride, ws.ctrl, 1 + 2 * slot, slot));
    ws.launches += 1;
    PF_CUDA(cudaGetLastError());
    return PF_OK;
}

}  // namespace pf

// ------------------------------------------------------------------------------------------------------------
// stage taps (host buffers, synchronous)
// ------------------------------------------------------------------------------------------------------------
using namespace pf;

namespace {
struct TapCtx {
    cudaStream_t stream = nullptr;
    Workspace ws;
    Pt *d_in = nullptr, *d_out = nullptr;
    int* d_counts = nullptr;   // [0] n_in, [1] n_out, [2] zero, [3] n_out of the empty cloud
    double* d_center = nullptr;
    ~TapCtx() {
        workspace_destroy(ws);
        cudaFree(d_in); cudaFree(d_out); cudaFree(d_counts); cudaFree(d_center);
        if (stream) cudaStreamDestroy(stream);
    }
};

int tap_voxel(int device, int mode, const pf_point* in, int n, const double* center, float leaf, int k_new, float theta_p, int theta_max,
              pf_point* out, int* n_out) {
    PF_REQUIRE(n >= 0 && (in || n == 0) && out && n_out, "bad argument");
    PF_REQUIRE(leaf > 0.f, "leaf must be positive");
    PF_CUDA(cudaSetDevice(device));
    TapCtx t;
    PF_CUDA(cudaStreamCreateWithFlags(&t.stream, cudaStreamNonBlocking));
    const int cap = n > 0 ? n : 1;
    PF_CHECK(workspace_create(t.ws, cap, t.stream));
    PF_CUDA(cudaMalloc(&t.d_in, sizeof(Pt) * cap));
    PF_CUDA(cudaMalloc(&t.d_out, sizeof(Pt) * cap));
    PF_CUDA(cudaMalloc(&t.d_counts, sizeof(int) * 4));
    PF_CUDA(cudaMalloc(&t.d_center, sizeof(double) * 3));
    int counts[4] = {n, 0, 0, 0};
    PF_CUDA(cudaMemcpyAsync(t.d_counts, counts, sizeof(counts), cudaMemcpyHostToDevice, t.stream));
    if (n) PF_CUDA(cudaMemcpyAsync(t.d_in, in, sizeof(Pt) * n, cudaMemcpyHostToDevice, t.stream));
    if (center) PF_CUDA(cudaMemcpyAsync(t.d_center, center, sizeof(double) * 3, cudaMemcpyHostToDevice, t.stream));
    VoxParams P{};
    P.mode = mode;
    P.c[0] = VoxCloud{t.d_in, t.d_counts + 0, t.d_out, t.d_counts + 1, leaf, 0};
    P.c[1] = VoxCloud{t.d_in, t.d_counts + 2, t.d_out, t.d_counts + 3, leaf, 0};
    P.center = t.d_center;
    P.k_new = k_new; P.theta_p = theta_p; P.theta_max = theta_max;
    PF_CHECK(workspace_begin_step(t.ws));
    PF_CHECK(voxelize(t.ws, P, 0, cap, 0));
    unsigned err = 0;
    PF_CUDA(cudaMemcpyAsync(counts, t.d_counts, sizeof(counts), cudaMemcpyDeviceToHost, t.stream));
    PF_CUDA(cudaMemcpyAsync(&err, t.ws.ctrl + kSlotBase + 15, sizeof(unsigned), cudaMemcpyDeviceToHost, t.stream));
    PF_CUDA(cudaStreamSynchronize(t.stream));
    if (err) { set_error("voxel grid: leaf size %g too small for the extent of the input (index would overflow)", leaf); return PF_ERR_INVALID; }
    *n_out = counts[1];
    if (counts[1]) PF_CUDA(cudaMemcpy(out, t.d_out, sizeof(Pt) * counts[1], cudaMemcpyDeviceToHost));
    return PF_OK;
}
}  // namespace

extern "C" int pf_voxel_downsample(int device, const pf_point* in, int n, float leaf, pf_point* out, int* n_out) {
    return tap_voxel(device, VOX_PCL, in, n, nullptr, leaf, 0, 0.f, 0, out, n_out);
}

extern "C" int pf_map_update(int device, const pf_point* in, int n, const double center[3], float leaf, int k_new, float theta_p,
                             int theta_max, pf_point* out, int* n_out) {
    PF_REQUIRE(center, "null center");
    return tap_voxel(device, VOX_MAP, in, n, center, leaf, k_new, theta_p, theta_max, out, n_out);
}
