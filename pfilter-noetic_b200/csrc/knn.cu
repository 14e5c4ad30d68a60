// K3: search-grid build (cell sort) for the two local maps, and the pf_knn5 stage tap.  See knn.cuh.
#include "knn.cuh"

namespace pf {

__device__ __forceinline__ unsigned f2ord_k(float f) {
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f_k(unsigned k) {
    unsigned u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(u);
}

// state slot: [0..5] map0 {~min, max}, [6..11] map1, [12..13] counts n0, n1, [14] n_total, [15] error
__global__ void __launch_bounds__(256) k_grid_bounds(GridBuild G) {
    PF_PDL_ENTRY();
    const int kind = blockIdx.y;
    const int n = *G.n_map[kind];
    const Pt* m = G.map[kind];
    unsigned mn[3] = {0u, 0u, 0u}, mx[3] = {0u, 0u, 0u};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        Pt p = m[i];
        unsigned kx = f2ord_k(p.x), ky = f2ord_k(p.y), kz = f2ord_k(p.z);
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
        if (v) atomicMax(&G.state[kind * 6 + threadIdx.x], v);
    }
}

// one block per map: origin / dims from the bounds
__global__ void k_grid_geom(GridBuild G) {
    PF_PDL_ENTRY();
    const int kind = blockIdx.x;
    if (threadIdx.x != 0) return;
    const unsigned* s = G.state + kind * 6;
    int* geom = G.geom[kind];
    const int n = *G.n_map[kind];
    if (kind == 0) reinterpret_cast<int*>(G.state)[14] = *G.n_map[0] + *G.n_map[1];
    if (n == 0 || s[3] == 0u) {
        for (int a = 0; a < 3; ++a) { geom[a] = 0; geom[3 + a] = 0; }
        return;
    }
    long long prod = 1;
    for (int a = 0; a < 3; ++a) {
        const float lo = floorf(ord2f_k(~s[a])), hi = floorf(ord2f_k(s[3 + a]));
        const long long o = (long long)fminf(fmaxf(lo, -1.0e9f), 1.0e9f), e = (long long)fminf(fmaxf(hi, -1.0e9f), 1.0e9f);
        const long long d = e - o + 1;
        geom[a] = (int)o;
        geom[3 + a] = (int)(d > 0x7fffffff ? 0x7fffffff : d);
        prod = (d > (1ll << 40) || prod > (1ll << 40)) ? (1ll << 41) : prod * d;
    }
    // a non-positive extent can only come from non-finite coordinates (a diverged pose turns the appended points into NaN / inf):
    // treated like an extent beyond the capacity -- the grid is void and the frame reports it
    if (geom[3] <= 0 || geom[4] <= 0 || geom[5] <= 0) prod = (1ll << 41);
    if (prod > kGridCellCap) {   // map extent larger than the search grid capacity
        atomicOr(&G.state[15], 1u);
        for (int a = 0; a < 3; ++a) geom[3 + a] = 0;
    }
}

// The cell sort is a counting sort: zero the per-cell counters, count (one atomic per point), exclusive scan over the cells
// (single pass, chained look-back), scatter (one atomic per point on the cell's cursor).  The order of the points inside a
// cell is whatever the atomics produce; the search does not depend on it (exact distances, ties broken by the carried index).
__global__ void __launch_bounds__(256) k_grid_clear(GridBuild G) {
    PF_PDL_ENTRY();
    const int kind = blockIdx.y;
    const int* geom = G.geom[kind];
    const long long cells = (long long)geom[3] * geom[4] * geom[5];
    int* cs = G.cell_start[kind];
    const int stride = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
    for (long long c = t0; c < cells; c += stride) cs[c] = 0;
}

__global__ void __launch_bounds__(256) k_grid_count(GridBuild G, uint32_t* __restrict__ keys) {
    PF_PDL_ENTRY();
    const int kind = blockIdx.y;
    const int n = *G.n_map[kind];
    const int base = kind == 0 ? 0 : *G.n_map[0];
    const int* geom = G.geom[kind];
    const int ox = geom[0], oy = geom[1], oz = geom[2], dx = geom[3], dy = geom[4], dz = geom[5];
    if ((long long)dx * dy * dz <= 0) return;
    const Pt* m = G.map[kind];
    int* cs = G.cell_start[kind];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const Pt p = m[i];
        const int cx = (int)floorf(p.x) - ox, cy = (int)floorf(p.y) - oy, cz = (int)floorf(p.z) - oz;
        if ((unsigned)cx >= (unsigned)dx || (unsigned)cy >= (unsigned)dy || (unsigned)cz >= (unsigned)dz) {     // NaN coordinate: not in any cell
            keys[base + i] = 0xffffffffu;
            atomicOr(&G.state[15], 2u);
            continue;
        }
        const unsigned cell = (unsigned)(cx + (cy + cz * dy) * dx);
        keys[base + i] = cell;
        atomicAdd(&cs[cell], 1);
    }
}

constexpr int kGridScanTile = 2048;     // cells per tile of the scan (8 per thread)
__global__ void __launch_bounds__(256) k_grid_scan(GridBuild G, unsigned long long* status, int status_stride, unsigned* ctrl) {
    PF_PDL_ENTRY();
    const int kind = blockIdx.y;
    const int* geom = G.geom[kind];
    const long long cells = (long long)geom[3] * geom[4] * geom[5];
    const int ntiles = (int)((cells + kGridScanTile - 1) / kGridScanTile);
    int* cs = G.cell_start[kind];
    int* ce = G.cell_end[kind];
    __shared__ int s_tile;
    __shared__ int s_tmp[9];
    __shared__ unsigned s_look[kScanSmemWords];
    const int tid = threadIdx.x;
    while (true) {
        __syncthreads();
        if (tid == 0) s_tile = (int)atomicAdd(&ctrl[9 + kind], 1u);
        __syncthreads();
        const int tile = s_tile;
        if (tile >= ntiles) return;
        const long long c0 = (long long)tile * kGridScanTile + 8 * tid;
        int v[8];
        if (c0 + 8 <= cells) {
            const int4 a = *reinterpret_cast<const int4*>(cs + c0), b = *reinterpret_cast<const int4*>(cs + c0 + 4);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = c0 + k < cells ? cs[c0 + k] : 0;
        }
        int sum = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) { const int t = v[k]; v[k] = sum; sum += t; }
        int total;
        const int excl = block_scan_excl_256(sum, s_tmp, &total);
        const unsigned tag = (ctrl[0] << 3) | 5u;
        const unsigned gbase = chained_scan_exclusive(status + (size_t)kind * status_stride, tag, tile, (unsigned)total, s_look);
        const int off = (int)gbase + excl;
        if (c0 + 8 <= cells) {
            const int4 a = make_int4(off + v[0], off + v[1], off + v[2], off + v[3]), b = make_int4(off + v[4], off + v[5], off + v[6], off + v[7]);
            *reinterpret_cast<int4*>(cs + c0) = a; *reinterpret_cast<int4*>(cs + c0 + 4) = b;
            *reinterpret_cast<int4*>(ce + c0) = a; *reinterpret_cast<int4*>(ce + c0 + 4) = b;     // cursor of the scatter, ends up as the cell's end
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (c0 + k < cells) { cs[c0 + k] = off + v[k]; ce[c0 + k] = off + v[k]; }
        }
    }
}

__global__ void __launch_bounds__(256) k_grid_scatter(GridBuild G, const uint32_t* __restrict__ keys) {
    PF_PDL_ENTRY();
    const int kind = blockIdx.y;
    const int n = *G.n_map[kind];
    const int base = kind == 0 ? 0 : *G.n_map[0];
    if (G.geom[kind][3] == 0) return;
    const Pt* m = G.map[kind];
    int* ce = G.cell_end[kind];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const Pt v = m[i];
        if (keys[base + i] == 0xffffffffu) continue;
        const int pos = atomicAdd(&ce[keys[base + i]], 1);
        G.pts[kind][pos] = make_float4(v.x, v.y, v.z, __int_as_float(i));
    }
}

int build_grids(Workspace& ws, const GridBuild& G_in, int slot, int cap0, int cap1) {
    GridBuild G = G_in;
    G.state = ws.ctrl + kSlotBase + slot * kSlotWords;
    const int cap = cap0 + cap1;
    PF_REQUIRE(cap <= ws.cap, "build_grids: %d points exceed workspace capacity %d", cap, ws.cap);
    PF_REQUIRE(ws.status_stride >= kGridCellCap / kGridScanTile + 8, "build_grids: workspace look-back array too small");
    const int capmax = cap0 > cap1 ? cap0 : cap1;
    int nblk = div_up(capmax, 256 * 4);
    if (nblk > 4 * kSMs) nblk = 4 * kSMs;
    if (nblk < 1) nblk = 1;
    PF_CUDA(launch_pdl(k_grid_bounds, dim3(nblk, 2), dim3(256), 0, ws.stream, G));
    PF_CUDA(launch_pdl(k_grid_geom, dim3(2), dim3(32), 0, ws.stream, G));
    PF_CUDA(launch_pdl(k_grid_clear, dim3(4 * kSMs, 2), dim3(256), 0, ws.stream, G));
    PF_CUDA(launch_pdl(k_grid_count, dim3(nblk, 2), dim3(256), 0, ws.stream, G, ws.keys[0]));
    PF_CUDA(launch_pdl(k_grid_scan, dim3(4 * kSMs, 2), dim3(256), 0, ws.stream, G, ws.scan_status, ws.status_stride, ws.ctrl));
    PF_CUDA(launch_pdl(k_grid_scatter, dim3(nblk, 2), dim3(256), 0, ws.stream, G, ws.keys[0]));
    ws.launches += 6;
    PF_CUDA(cudaGetLastError());
    return PF_OK;
}

// Two queries per warp (a half warp each): a query is a chain of dependent memory round trips (cell table, centre row, other rows),
// so the number of queries in flight sets the throughput; the ~10-70 candidates of a query keep 16 lanes busy.
#ifndef PF_KNN_TAP_GROUP
#define PF_KNN_TAP_GROUP 16
#endif
constexpr int kTapGroup = PF_KNN_TAP_GROUP;
__global__ void __launch_bounds__(256, 8) k_knn5_tap(KnnGrid g, const float4* __restrict__ q, int nq, int* __restrict__ idx_out,
                                                  float* __restrict__ d2_out) {
    const int qi = (blockIdx.x * blockDim.x + threadIdx.x) / kTapGroup;
    if (qi >= nq) return;          // uniform within a group
    const float4 p = q[qi];
    int idx[5];
    float d2[5];
    const bool ok = knn5_group<kTapGroup>(g, p.x, p.y, p.z, idx, d2);
    const unsigned lane = lane_id() & (kTapGroup - 1);
    if (lane < 5) {
        idx_out[5 * qi + lane] = ok ? idx[lane] : -1;
        d2_out[5 * qi + lane] = ok ? d2[lane] : __int_as_float(0x7f800000);
    }
}

}  // namespace pf

using namespace pf;

namespace {
struct KnnTap {
    cudaStream_t stream = nullptr;
    Workspace ws;
    Pt* d_map = nullptr;
    float4 *d_pts = nullptr, *d_q = nullptr;
    int *d_cs = nullptr, *d_ce = nullptr, *d_geom = nullptr, *d_counts = nullptr, *d_idx = nullptr;
    float* d_d2 = nullptr;
    ~KnnTap() {
        workspace_destroy(ws);
        cudaFree(d_map); cudaFree(d_pts); cudaFree(d_q); cudaFree(d_cs); cudaFree(d_ce); cudaFree(d_geom); cudaFree(d_counts);
        cudaFree(d_idx); cudaFree(d_d2);
        if (stream) cudaStreamDestroy(stream);
    }
};
}  // namespace

static int knn5_tap(int device, const pf_point* map, int m, const float* queries_xyz4, int q, int32_t* idx, float* d2, int reps,
                    float* ms_build, float* ms_query) {
    PF_REQUIRE(m >= 0 && q >= 0 && (map || m == 0) && (queries_xyz4 || q == 0) && idx && d2 && reps >= 1, "bad argument");
    PF_CUDA(cudaSetDevice(device));
    KnnTap t;
    PF_CUDA(cudaStreamCreateWithFlags(&t.stream, cudaStreamNonBlocking));
    const int mc = m > 0 ? m : 1, qc = q > 0 ? q : 1;
    PF_CHECK(workspace_create(t.ws, mc, t.stream));
    PF_CUDA(cudaMalloc(&t.d_map, sizeof(Pt) * mc));
    PF_CUDA(cudaMalloc(&t.d_pts, sizeof(float4) * mc));
    PF_CUDA(cudaMalloc(&t.d_q, sizeof(float4) * qc));
    PF_CUDA(cudaMalloc(&t.d_cs, sizeof(int) * (size_t)kGridCellCap));
    PF_CUDA(cudaMalloc(&t.d_ce, sizeof(int) * (size_t)kGridCellCap));
    PF_CUDA(cudaMalloc(&t.d_geom, sizeof(int) * 12));
    PF_CUDA(cudaMalloc(&t.d_counts, sizeof(int) * 2));
    PF_CUDA(cudaMalloc(&t.d_idx, sizeof(int) * 5 * qc));
    PF_CUDA(cudaMalloc(&t.d_d2, sizeof(float) * 5 * qc));
    int counts[2] = {m, 0};
    PF_CUDA(cudaMemcpyAsync(t.d_counts, counts, sizeof(counts), cudaMemcpyHostToDevice, t.stream));
    if (m) PF_CUDA(cudaMemcpyAsync(t.d_map, map, sizeof(Pt) * m, cudaMemcpyHostToDevice, t.stream));
    if (q) PF_CUDA(cudaMemcpyAsync(t.d_q, queries_xyz4, sizeof(float4) * q, cudaMemcpyHostToDevice, t.stream));
    GridBuild G{};
    G.map[0] = t.d_map; G.map[1] = t.d_map;
    G.n_map[0] = t.d_counts; G.n_map[1] = t.d_counts + 1;
    G.pts[0] = t.d_pts; G.pts[1] = t.d_pts;
    G.cell_start[0] = t.d_cs; G.cell_start[1] = t.d_cs;
    G.cell_end[0] = t.d_ce; G.cell_end[1] = t.d_ce;
    G.geom[0] = t.d_geom; G.geom[1] = t.d_geom + 6;
    cudaEvent_t ev[3];
    for (int i = 0; i < 3; ++i) PF_CUDA(cudaEventCreate(&ev[i]));
    KnnGrid g{t.d_pts, t.d_cs, t.d_ce, t.d_geom};
    float sum_b = 0.f, sum_q = 0.f;
    for (int r = 0; r < reps; ++r) {          // repetition 0 is the warm-up when reps > 1
        PF_CUDA(cudaEventRecord(ev[0], t.stream));
        PF_CHECK(workspace_begin_step(t.ws));
        PF_CHECK(build_grids(t.ws, G, 0, mc, 0));
        PF_CUDA(cudaEventRecord(ev[1], t.stream));
        if (q) k_knn5_tap<<<div_up(q, 256 / kTapGroup), 256, 0, t.stream>>>(g, t.d_q, q, t.d_idx, t.d_d2);
        PF_CUDA(cudaEventRecord(ev[2], t.stream));
        PF_CUDA(cudaStreamSynchronize(t.stream));
        float a = 0.f, b = 0.f;
        PF_CUDA(cudaEventElapsedTime(&a, ev[0], ev[1]));
        PF_CUDA(cudaEventElapsedTime(&b, ev[1], ev[2]));
        if (r > 0 || reps == 1) { sum_b += a; sum_q += b; }
    }
    for (int i = 0; i < 3; ++i) cudaEventDestroy(ev[i]);
    if (ms_build) *ms_build = sum_b / (reps > 1 ? reps - 1 : 1);
    if (ms_query) *ms_query = sum_q / (reps > 1 ? reps - 1 : 1);
    unsigned err = 0;
    PF_CUDA(cudaMemcpyAsync(&err, t.ws.ctrl + kSlotBase + 15, sizeof(unsigned), cudaMemcpyDeviceToHost, t.stream));
    if (q) {
        PF_CUDA(cudaMemcpyAsync(idx, t.d_idx, sizeof(int) * 5 * q, cudaMemcpyDeviceToHost, t.stream));
        PF_CUDA(cudaMemcpyAsync(d2, t.d_d2, sizeof(float) * 5 * q, cudaMemcpyDeviceToHost, t.stream));
    }
    PF_CUDA(cudaStreamSynchronize(t.stream));
    if (err) { set_error("map extent exceeds the search grid capacity (%d cells of 1 m)", kGridCellCap); return PF_ERR_CAPACITY; }
    return PF_OK;
}

extern "C" int pf_knn5(int device, const pf_point* map, int m, const float* queries_xyz4, int q, int32_t* idx, float* d2) {
    return knn5_tap(device, map, m, queries_xyz4, q, idx, d2, 1, nullptr, nullptr);
}

extern "C" int pf_knn5_timed(int device, const pf_point* map, int m, const float* queries_xyz4, int q, int32_t* idx, float* d2, int reps,
                             float* ms_build, float* ms_query) {
    return knn5_tap(device, map, m, queries_xyz4, q, idx, d2, reps, ms_build, ms_query);
}
