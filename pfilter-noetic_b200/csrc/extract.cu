// K1: LiDAR feature extraction on sm_100a.
//
// Replaces LaserProcessingClass::featureExtraction / featureExtractionFromSector
// (/root/reference/src/laserProcessingClass.cpp:10-96, :99-209).  Arithmetic spec: SURVEY.md appendix A.1.
//
// Two kernels per group of scans (a group is sized to stay L2-resident between them):
//   k_ring_classify : one thread per point, coalesced float4 loads; ring id (elevation angle, :25-61) -> 1 byte,
//                     per-256-point tile the [min,max] ring id; clears the per-point label.
//   k_ring_extract  : one CTA per (scan, ring).  Gathers the ring's points (stable, :62) from the candidate
//                     tiles into shared memory, 11-tap float curvature (:73-80), 6 sector warps doing the greedy
//                     edge pick by repeated warp arg-max (equivalent to the sorted descending walk of :110-148,
//                     no sort needed), neighbour suppression, surf = unpicked (:198-205); output offsets of the
//                     64 rings of a scan are chained through a decoupled look-back on (epoch, counts) words, so
//                     the compacted edge/surf clouds are written once, straight from shared memory.
// HBM traffic per point: 16 B read + 16 B written (+ ring id / label bytes); the second read of the points by
// k_ring_extract hits L2.
#include "common.cuh"

namespace pf {

constexpr int kTile = 256;          // points per classify tile
constexpr int kExtractThreads = 256;
constexpr int kMaxLines = 64;
constexpr int kEdgePerSector = 20;  // src/laserProcessingClass.cpp:121
constexpr int kSectors = 6;         // :81

struct ExtractParams {
    const float4* pts;        // [batch][stride]
    const int* n;             // [batch]
    uint8_t* ringid;          // [batch][stride]
    uint16_t* tile_rng;       // [batch][tiles]  min | max << 8
    uint8_t* label;           // [batch][stride] or null
    float4* edge;             // [batch][edge_stride]
    float4* surf;             // [batch][stride]
    int* n_edge;              // [batch]
    int* n_surf;              // [batch]
    unsigned long long* done; // [batch][64] look-back words
    unsigned int* ctrl;       // [0] ticket, [1] epoch, [2] error bits
    int stride, tiles, edge_stride, batch;
    int num_lines, rcap, maxtl;
    double min_d, max_d;
};

// ring id of one point or 255 (dropped); src/laserProcessingClass.cpp:25-61
__device__ __forceinline__ int ring_id_dev(float x, float y, float z, int num_lines, double min_d, double max_d) {
    float s = __fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y));
    double distance = (double)__fsqrt_rn(s);   // sqrt(float) overload, see oracle/shim/pcl/point_types.h
    if (distance < min_d || distance > max_d) return 255;
    double angle = atan((double)z / distance) * 180 / 3.14159265358979323846;
    int id;
    if (num_lines == 64) {
        if (angle >= -8.83) id = (int)((2 - angle) * 3.0 + 0.5);
        else id = 32 + (int)((-8.83 - angle) * 2.0 + 0.5);
        if (angle > 2 || angle < -24.33 || id > 63 || id < 0) return 255;
    } else if (num_lines == 32) {
        id = (int)((angle + 92.0 / 3.0) * 3.0 / 4.0);
        if (id > 31 || id < 0) return 255;
    } else {
        id = (int)((angle + 15) / 2 + 0.5);
        if (id > 15 || id < 0) return 255;
    }
    return id;
}

__global__ void __launch_bounds__(kTile) k_ring_classify(ExtractParams P) {
    const int s = blockIdx.y, tile = blockIdx.x, tid = threadIdx.x;
    if (blockIdx.x == 0 && blockIdx.y == 0 && tid == 0) {
        P.ctrl[0] = 0;                // ticket for the extract kernel that follows in stream order
        atomicAdd(&P.ctrl[1], 1u);    // new epoch invalidates all look-back words of earlier launches
    }
    const int n = P.n[s];
    if (tile * kTile >= n) return;
    const int i = tile * kTile + tid;
    const size_t g = (size_t)s * P.stride + i;
    int ring = 255;
    if (i < n) {
        float4 p = ld_stream_f4(P.pts + g);
        ring = ring_id_dev(p.x, p.y, p.z, P.num_lines, P.min_d, P.max_d);
        if (P.label) P.label[g] = 0;
    }
    P.ringid[g] = (uint8_t)ring;
    int lo = ring == 255 ? 255 : ring, hi = ring == 255 ? 0 : ring;
    lo = __reduce_min_sync(0xffffffffu, lo);
    hi = __reduce_max_sync(0xffffffffu, hi);
    __shared__ int wlo[kTile / 32], whi[kTile / 32];
    if ((tid & 31) == 0) { wlo[tid >> 5] = lo; whi[tid >> 5] = hi; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < kTile / 32; ++w) { lo = min(lo, wlo[w]); hi = max(hi, whi[w]); }
        P.tile_rng[(size_t)s * P.tiles + tile] = (uint16_t)(lo | (hi << 8));
    }
}

__global__ void k_set_int(int* p, int v) {
    if (threadIdx.x == 0) *p = v;
}

// exclusive scan of a[0..n) in shared memory by the whole CTA (kExtractThreads threads); returns the total
__device__ int block_excl_scan(int* a, int n, int* warp_tmp /*[9]*/) {
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    int carry = 0;
    for (int base = 0; base < n; base += kExtractThreads) {
        int i = base + tid;
        int v = i < n ? a[i] : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_tmp[w] = x;
        __syncthreads();
        int woff = 0, tot = 0;
#pragma unroll
        for (int k = 0; k < kExtractThreads / 32; ++k) {
            int t = warp_tmp[k];
            if (k < w) woff += t;
            tot += t;
        }
        if (i < n) a[i] = carry + woff + x - v;
        carry += tot;
        __syncthreads();
    }
    return carry;
}

__device__ __forceinline__ double step_d2(const float4* sp, int a, int b) {   // :129-132 float diffs, double squares
    float4 p = sp[a], q = sp[b];
    double dx = (double)__fsub_rn(p.x, q.x), dy = (double)__fsub_rn(p.y, q.y), dz = (double)__fsub_rn(p.z, q.z);
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

__global__ void __launch_bounds__(kExtractThreads) k_ring_extract(ExtractParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* sp = reinterpret_cast<float4*>(smem_raw);
    unsigned long long* scurv = reinterpret_cast<unsigned long long*>(sp + P.rcap);
    int* ssrc = reinterpret_cast<int*>(scurv + P.rcap);
    int* tl_tile = ssrc + P.rcap;
    int* tl_off = tl_tile + P.maxtl;
    uint8_t* sflag = reinterpret_cast<uint8_t*>(tl_off + P.maxtl + 1);

    __shared__ int s_work, s_ncand;
    __shared__ int s_warp[kExtractThreads / 32 + 1];
    __shared__ int s_edge_ids[kSectors][kEdgePerSector];
    __shared__ int s_ecnt[kSectors], s_scnt[kSectors];
    __shared__ int s_eoff, s_soff;

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    constexpr int NW = kExtractThreads / 32;

    // work item by ticket: CTAs that hold ticket t only ever wait on tickets < t, which are already running
    if (tid == 0) { s_work = (int)atomicAdd(&P.ctrl[0], 1u); s_ncand = 0; }
    __syncthreads();
    const int s = s_work / P.num_lines, r = s_work % P.num_lines;
    const unsigned epoch = *reinterpret_cast<volatile unsigned int*>(&P.ctrl[1]);
    const int n = P.n[s];
    const int ntiles = (n + kTile - 1) / kTile;
    const uint8_t* rid = P.ringid + (size_t)s * P.stride;
    const float4* pts = P.pts + (size_t)s * P.stride;

    // A. candidate tiles (ascending): tiles whose [min,max] ring range contains r
    for (int base = 0; base < ntiles; base += kExtractThreads) {
        int t = base + tid;
        bool cand = false;
        if (t < ntiles) {
            unsigned rng = P.tile_rng[(size_t)s * P.tiles + t];
            cand = (int)(rng & 0xff) <= r && r <= (int)(rng >> 8);
        }
        unsigned m = __ballot_sync(0xffffffffu, cand);
        if (lane == 0) s_warp[w] = __popc(m);
        __syncthreads();
        int off = s_ncand;
        for (int k = 0; k < w; ++k) off += s_warp[k];
        if (cand) tl_tile[off + __popc(m & lanemask_lt())] = t;
        __syncthreads();
        if (tid == 0) { int tot = 0; for (int k = 0; k < NW; ++k) tot += s_warp[k]; s_ncand += tot; }
        __syncthreads();
    }
    const int ncand = s_ncand;

    // B. matches per candidate tile
    for (int c = w; c < ncand; c += NW) {
        const int tbase = tl_tile[c] * kTile;
        int cnt = 0;
#pragma unroll
        for (int k = 0; k < kTile / 32; ++k) {
            int i = tbase + k * 32 + lane;
            int b = i < n ? rid[i] : 255;
            cnt += __popc(__ballot_sync(0xffffffffu, b == r));
        }
        if (lane == 0) tl_off[c] = cnt;
    }
    __syncthreads();
    const int nr = block_excl_scan(tl_off, ncand, s_warp);
    bool active = nr >= 131;                                   // :67
    if (nr > P.rcap) {
        active = false;
        if (tid == 0) atomicOr(&P.ctrl[2], 1u);                // ring larger than the shared-memory capacity
    }

    int e_total = 0, s_total = 0;
    if (active) {
        // C. stable gather of the ring into shared memory
        for (int c = w; c < ncand; c += NW) {
            const int tbase = tl_tile[c] * kTile;
            int pos = tl_off[c];
            unsigned m[kTile / 32];
#pragma unroll
            for (int k = 0; k < kTile / 32; ++k) {
                int i = tbase + k * 32 + lane;
                int b = i < n ? rid[i] : 255;
                m[k] = __ballot_sync(0xffffffffu, b == r);
            }
            float4 v[kTile / 32];
#pragma unroll
            for (int k = 0; k < kTile / 32; ++k)
                if (m[k] >> lane & 1u) v[k] = __ldg(pts + tbase + k * 32 + lane);
#pragma unroll
            for (int k = 0; k < kTile / 32; ++k) {
                if (m[k] >> lane & 1u) {
                    int d = pos + __popc(m[k] & lanemask_lt());
                    sp[d] = v[k];
                    ssrc[d] = tbase + k * 32 + lane;
                }
                pos += __popc(m[k]);
            }
        }
        __syncthreads();

        // D. curvature (:73-80): float sums left to right, squares and their sum in double
        for (int j = tid; j < nr; j += kExtractThreads) {
            sflag[j] = 0;
            if (j >= 5 && j < nr - 5) {
                float d[3];
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    auto at = [&](int q) { const float* f = reinterpret_cast<const float*>(&sp[q]); return f[a]; };
                    float t = __fadd_rn(at(j - 5), at(j - 4));
                    t = __fadd_rn(t, at(j - 3));
                    t = __fadd_rn(t, at(j - 2));
                    t = __fadd_rn(t, at(j - 1));
                    t = __fsub_rn(t, __fmul_rn(10.0f, at(j)));
                    t = __fadd_rn(t, at(j + 1));
                    t = __fadd_rn(t, at(j + 2));
                    t = __fadd_rn(t, at(j + 3));
                    t = __fadd_rn(t, at(j + 4));
                    t = __fadd_rn(t, at(j + 5));
                    d[a] = t;
                }
                double dx = d[0], dy = d[1], dz = d[2];
                double val = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                scurv[j] = (unsigned long long)__double_as_longlong(val);   // non-negative: bit pattern is order preserving
            }
        }
        __syncthreads();

        // E. one warp per sector (:81-92, :99-209)
        if (w < kSectors) {
            const int total = nr - 10, L = total / kSectors;
            const int lo = L * w, hi = (w == kSectors - 1) ? total - 1 : L * (w + 1) - 1;   // hi excluded (:83-88)
            const int a = lo + 5, b = hi + 5;
            int cnt = 0;
            while (true) {
                unsigned long long bk = 0ull;
                int bi = -1;
                for (int i = a + lane; i < b; i += 32) {
                    if (!sflag[i]) {
                        unsigned long long k = scurv[i];
                        if (bi < 0 || k > bk || (k == bk && i > bi)) { bk = k; bi = i; }
                    }
                }
#pragma unroll
                for (int o = 16; o >= 1; o >>= 1) {
                    unsigned long long ok = __shfl_xor_sync(0xffffffffu, bk, o);
                    int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                    if (oi >= 0 && (bi < 0 || ok > bk || (ok == bk && oi > bi))) { bk = ok; bi = oi; }
                }
                if (bi < 0) break;
                if (!(__longlong_as_double((long long)bk) > 0.1)) break;                    // :114
                cnt++;                                                                      // :118-119
                if (lane == 0) sflag[bi] = 1;
                if (cnt <= kEdgePerSector) {                                                // :121-126
                    if (lane == 0) s_edge_ids[w][cnt - 1] = bi;
                } else {
                    break;
                }
                if (lane == 0) {
                    for (int k = 1; k <= 5; ++k) {                                          // :128-136
                        if (step_d2(sp, bi + k, bi + k - 1) > 0.05) break;
                        if (bi + k < b) sflag[bi + k] = 1;
                    }
                    for (int k = -1; k >= -5; --k) {                                        // :137-145
                        if (step_d2(sp, bi + k, bi + k + 1) > 0.05) break;
                        if (bi + k >= a) sflag[bi + k] = 1;
                    }
                }
                __syncwarp();
            }
            __syncwarp();
            int sc = 0;
            for (int base = a; base < b; base += 32) {
                int i = base + lane;
                sc += __popc(__ballot_sync(0xffffffffu, i < b && !sflag[i]));
            }
            if (lane == 0) { s_ecnt[w] = min(cnt, kEdgePerSector); s_scnt[w] = sc; }
        }
        __syncthreads();
        for (int k = 0; k < kSectors; ++k) { e_total += s_ecnt[k]; s_total += s_scnt[k]; }
    }

    // F. publish this ring's counts, then look back over the earlier rings of the same scan
    unsigned long long* done = P.done + (size_t)s * kMaxLines;
    if (tid == 0)
        st_release_u64(done + r, ((unsigned long long)epoch << 32) | ((unsigned long long)e_total << 20) | (unsigned long long)s_total);
    if (w == 0) {
        int e = 0, su = 0;
        for (int q = lane; q < r; q += 32) {
            unsigned long long v;
            do { v = ld_acquire_u64(done + q); } while ((unsigned)(v >> 32) != epoch);
            e += (int)((v >> 20) & 0xfffu);
            su += (int)(v & 0xfffffu);
        }
        e = __reduce_add_sync(0xffffffffu, e);
        su = __reduce_add_sync(0xffffffffu, su);
        if (lane == 0) {
            s_eoff = e; s_soff = su;
            if (r == P.num_lines - 1) { P.n_edge[s] = e + e_total; P.n_surf[s] = su + s_total; }
        }
    }
    __syncthreads();
    if (!active) return;

    // G. write the compacted clouds straight from shared memory
    if (w < kSectors) {
        const int total = nr - 10, L = total / kSectors;
        const int lo = L * w, hi = (w == kSectors - 1) ? total - 1 : L * (w + 1) - 1;
        const int a = lo + 5, b = hi + 5;
        int eo = s_eoff, so = s_soff;
        for (int k = 0; k < w; ++k) { eo += s_ecnt[k]; so += s_scnt[k]; }
        uint8_t* label = P.label ? P.label + (size_t)s * P.stride : nullptr;
        if (lane < s_ecnt[w]) {
            int id = s_edge_ids[w][lane];
            P.edge[(size_t)s * P.edge_stride + eo + lane] = sp[id];
            if (label) label[ssrc[id]] = 1;
        }
        float4* surf = P.surf + (size_t)s * P.stride + so;
        for (int base = a; base < b; base += 32) {
            int i = base + lane;
            bool f = i < b && !sflag[i];
            unsigned m = __ballot_sync(0xffffffffu, f);
            if (f) {
                st_stream_f4(surf + __popc(m & lanemask_lt()), sp[i]);
                if (label) label[ssrc[i]] = 2;
            }
            surf += __popc(m);
        }
    }
}

}  // namespace pf

// ------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------
struct pf_extract {
    int device = 0;
    cudaStream_t stream = nullptr;
    pf_lidar_params lidar{};
    int stride = 0, tiles = 0, max_batch = 0, rcap = 0, maxtl = 0, edge_stride = 0;
    size_t smem = 0;
    int group = 1;
    // device
    float4* d_pts = nullptr;
    uint8_t* d_ringid = nullptr;
    uint16_t* d_tile_rng = nullptr;
    uint8_t* d_label = nullptr;
    float4* d_edge = nullptr;
    float4* d_surf = nullptr;
    int *d_n = nullptr, *d_n_edge = nullptr, *d_n_surf = nullptr;
    unsigned long long* d_done = nullptr;
    unsigned int* d_ctrl = nullptr;
    // pinned host
    int* h_counts = nullptr;       // [3 * max_batch]: n, n_edge, n_surf
    unsigned int* h_ctrl = nullptr;
    uint64_t launches = 0;
    // result of the last single-scan run (device resident hand-off to the odometry)
    int last_valid = 0;
    int last_n = 0;
};

namespace pf {

static int extract_launch(pf_extract* h, const float4* d_xyzi, const int* d_n, int batch, int stride, float4* d_edge,
                          int* d_n_edge, int edge_stride, float4* d_surf, int* d_n_surf, uint8_t* d_label) {
    PF_REQUIRE(stride % kTile == 0 && stride <= h->stride, "stride %d must be a multiple of %d and <= %d", stride, kTile, h->stride);
    PF_REQUIRE(batch >= 1 && batch <= h->max_batch, "batch %d outside 1..%d", batch, h->max_batch);
    PF_REQUIRE(edge_stride >= 120 * h->lidar.num_lines, "edge_stride %d < %d", edge_stride, 120 * h->lidar.num_lines);
    const int tiles = stride / kTile;
    for (int s0 = 0; s0 < batch; s0 += h->group) {
        const int nb = batch - s0 < h->group ? batch - s0 : h->group;
        ExtractParams P{};
        P.pts = d_xyzi + (size_t)s0 * stride;
        P.n = d_n + s0;
        P.ringid = h->d_ringid + (size_t)s0 * stride;
        P.tile_rng = h->d_tile_rng + (size_t)s0 * tiles;
        P.label = d_label ? d_label + (size_t)s0 * stride : nullptr;
        P.edge = d_edge + (size_t)s0 * edge_stride;
        P.surf = d_surf + (size_t)s0 * stride;
        P.n_edge = d_n_edge + s0;
        P.n_surf = d_n_surf + s0;
        P.done = h->d_done + (size_t)s0 * kMaxLines;
        P.ctrl = h->d_ctrl;
        P.stride = stride; P.tiles = tiles; P.edge_stride = edge_stride; P.batch = nb;
        P.num_lines = h->lidar.num_lines; P.rcap = h->rcap; P.maxtl = h->maxtl;
        P.min_d = h->lidar.min_distance; P.max_d = h->lidar.max_distance;
        k_ring_classify<<<dim3(tiles, nb), kTile, 0, h->stream>>>(P);
        k_ring_extract<<<nb * h->lidar.num_lines, kExtractThreads, h->smem, h->stream>>>(P);
        h->launches += 2;
    }
    PF_CUDA(cudaGetLastError());
    return PF_OK;
}

static int extract_check_ctrl(pf_extract* h) {
    // error bits are sticky on the device; read and clear
    PF_CUDA(cudaMemcpyAsync(h->h_ctrl, h->d_ctrl + 2, sizeof(unsigned), cudaMemcpyDeviceToHost, h->stream));
    PF_CUDA(cudaStreamSynchronize(h->stream));
    if (h->h_ctrl[0] & 1u) {
        PF_CUDA(cudaMemsetAsync(h->d_ctrl + 2, 0, sizeof(unsigned), h->stream));
        set_error("a ring holds more than max_ring_points = %d points", h->rcap);
        return PF_ERR_CAPACITY;
    }
    return PF_OK;
}

}  // namespace pf

using namespace pf;

extern "C" int pf_extract_create(const pf_lidar_params* lidar, const pf_extract_config* cfg, int device, pf_extract** out) {
    PF_REQUIRE(lidar && cfg && out, "null argument");
    PF_REQUIRE(lidar->num_lines == 16 || lidar->num_lines == 32 || lidar->num_lines == 64,
               "num_lines must be 16, 32 or 64 (src/laserProcessingClass.cpp:30-61), got %d", lidar->num_lines);
    PF_REQUIRE(cfg->max_points > 0 && cfg->max_batch > 0, "max_points and max_batch must be positive");
    int ndev = 0;
    PF_CUDA(cudaGetDeviceCount(&ndev));
    PF_REQUIRE(device >= 0 && device < ndev, "device %d not available (%d devices)", device, ndev);
    PF_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    PF_CUDA(cudaGetDeviceProperties(&prop, device));
    PF_REQUIRE(prop.major == 10, "pfilter_b200 needs an sm_100a device, found sm_%d%d", prop.major, prop.minor);

    pf_extract* h = new pf_extract();
    h->device = device;
    h->lidar = *lidar;
    h->stride = div_up(cfg->max_points, kTile) * kTile;
    h->tiles = h->stride / kTile;
    h->max_batch = cfg->max_batch;
    h->rcap = cfg->max_ring_points > 0 ? cfg->max_ring_points : 2560;
    h->rcap = div_up(h->rcap, 32) * 32;
    h->maxtl = h->tiles;
    h->edge_stride = 120 * lidar->num_lines;
    h->smem = (size_t)h->rcap * (16 + 8 + 4 + 1) + (size_t)(2 * h->maxtl + 1) * 4 + 64;
    if (h->smem > (size_t)prop.sharedMemPerBlockOptin) {
        set_error("max_ring_points %d / max_points %d need %zu B shared memory (> %zu)", h->rcap, cfg->max_points, h->smem,
                  (size_t)prop.sharedMemPerBlockOptin);
        delete h;
        return PF_ERR_INVALID;
    }
    PF_CUDA(cudaFuncSetAttribute(k_ring_extract, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem));
    // group of scans per launch pair: keep points + outputs of a group inside L2 (126 MB)
    {
        size_t per_scan = (size_t)h->stride * 16;
        int g = (int)((40u << 20) / per_scan);
        const char* env = getenv("PF_EXTRACT_GROUP");
        if (env) g = atoi(env);
        h->group = g < 1 ? 1 : g;
    }
    PF_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    const size_t np = (size_t)h->max_batch * h->stride;
    PF_CUDA(cudaMalloc(&h->d_pts, np * 16));
    PF_CUDA(cudaMalloc(&h->d_ringid, np));
    PF_CUDA(cudaMalloc(&h->d_tile_rng, (size_t)h->max_batch * h->tiles * 2));
    PF_CUDA(cudaMalloc(&h->d_label, np));
    PF_CUDA(cudaMalloc(&h->d_edge, (size_t)h->max_batch * h->edge_stride * 16));
    PF_CUDA(cudaMalloc(&h->d_surf, np * 16));
    PF_CUDA(cudaMalloc(&h->d_n, sizeof(int) * h->max_batch));
    PF_CUDA(cudaMalloc(&h->d_n_edge, sizeof(int) * h->max_batch));
    PF_CUDA(cudaMalloc(&h->d_n_surf, sizeof(int) * h->max_batch));
    PF_CUDA(cudaMalloc(&h->d_done, sizeof(unsigned long long) * h->max_batch * kMaxLines));
    PF_CUDA(cudaMalloc(&h->d_ctrl, sizeof(unsigned) * 4));
    PF_CUDA(cudaMemset(h->d_done, 0, sizeof(unsigned long long) * h->max_batch * kMaxLines));
    PF_CUDA(cudaMemset(h->d_ctrl, 0, sizeof(unsigned) * 4));
    PF_CUDA(cudaMallocHost(&h->h_counts, sizeof(int) * 3 * h->max_batch));
    PF_CUDA(cudaMallocHost(&h->h_ctrl, sizeof(unsigned) * 4));
    *out = h;
    return PF_OK;
}

extern "C" int pf_extract_destroy(pf_extract* h) {
    if (!h) return PF_OK;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    cudaFree(h->d_pts); cudaFree(h->d_ringid); cudaFree(h->d_tile_rng); cudaFree(h->d_label); cudaFree(h->d_edge);
    cudaFree(h->d_surf); cudaFree(h->d_n); cudaFree(h->d_n_edge); cudaFree(h->d_n_surf); cudaFree(h->d_done);
    cudaFree(h->d_ctrl);
    cudaFreeHost(h->h_counts); cudaFreeHost(h->h_ctrl);
    cudaStreamDestroy(h->stream);
    delete h;
    return PF_OK;
}

extern "C" int pf_extract_run_batch_device(pf_extract* h, const void* d_xyzi, const int* d_n, int batch, int stride,
                                           void* d_edge, int* d_n_edge, int edge_stride, void* d_surf, int* d_n_surf,
                                           uint8_t* d_label) {
    PF_REQUIRE(h && d_xyzi && d_n && d_edge && d_n_edge && d_surf && d_n_surf, "null argument");
    PF_CUDA(cudaSetDevice(h->device));
    return extract_launch(h, (const float4*)d_xyzi, d_n, batch, stride, (float4*)d_edge, d_n_edge, edge_stride,
                          (float4*)d_surf, d_n_surf, d_label);
}

extern "C" int pf_extract_sync(pf_extract* h) {
    PF_REQUIRE(h, "null handle");
    PF_CUDA(cudaSetDevice(h->device));
    return extract_check_ctrl(h);
}

extern "C" void* pf_extract_stream(pf_extract* h) { return h ? (void*)h->stream : nullptr; }

extern "C" int pf_extract_kernel_launches(pf_extract* h, uint64_t* launches) {
    PF_REQUIRE(h && launches, "null argument");
    *launches = h->launches;
    return PF_OK;
}

// Enqueue H2D + kernels for one scan; results stay on the device (used by pf_extract_run and the frame pipeline).
// device_input: xyzi is a device pointer (no copy).  The count travels as a kernel argument, so consecutive frames
// can be enqueued without a host synchronisation in between.
int pf_extract_enqueue_single(pf_extract* h, const float* xyzi, int n, int device_input) {
    PF_REQUIRE(h && (xyzi || n == 0), "null argument");
    PF_REQUIRE(n >= 0 && n <= h->stride, "scan of %d points exceeds max_points %d", n, h->stride);
    PF_CUDA(cudaSetDevice(h->device));
    k_set_int<<<1, 32, 0, h->stream>>>(h->d_n, n);
    h->launches += 1;
    const float4* src = h->d_pts;
    if (device_input) {
        src = reinterpret_cast<const float4*>(xyzi);
    } else if (n > 0) {
        PF_CUDA(cudaMemcpyAsync(h->d_pts, xyzi, (size_t)n * 16, cudaMemcpyHostToDevice, h->stream));
    }
    PF_CHECK(extract_launch(h, src ? src : h->d_pts, h->d_n, 1, h->stride, h->d_edge, h->d_n_edge, h->edge_stride, h->d_surf,
                            h->d_n_surf, h->d_label));
    h->last_valid = 1;
    h->last_n = n;
    return PF_OK;
}

// accessors for the device-resident hand-off (odom.cu)
void pf_extract_device_outputs(pf_extract* h, const float4** edge, const int** n_edge, const float4** surf, const int** n_surf,
                               cudaStream_t* stream, int* edge_cap, int* surf_cap) {
    *edge = h->d_edge; *n_edge = h->d_n_edge; *surf = h->d_surf; *n_surf = h->d_n_surf; *stream = h->stream;
    *edge_cap = h->edge_stride < h->last_n ? h->edge_stride : h->last_n; *surf_cap = h->last_n;   // upper bounds of the device counts
}

extern "C" int pf_extract_run_batch(pf_extract* h, const float* xyzi, const int* n, int batch, int stride, float* edge,
                                    int* n_edge, int edge_stride, float* surf, int* n_surf, uint8_t* label) {
    PF_REQUIRE(h && xyzi && n && edge && n_edge && surf && n_surf, "null argument");
    PF_REQUIRE(batch >= 1 && batch <= h->max_batch, "batch %d outside 1..%d", batch, h->max_batch);
    PF_REQUIRE(stride % kTile == 0 && stride <= h->stride, "stride %d must be a multiple of %d and <= %d", stride, kTile, h->stride);
    PF_CUDA(cudaSetDevice(h->device));
    for (int s = 0; s < batch; ++s) {
        PF_REQUIRE(n[s] >= 0 && n[s] <= stride, "scan %d: %d points exceed stride %d", s, n[s], stride);
        h->h_counts[s] = n[s];
    }
    PF_CUDA(cudaMemcpyAsync(h->d_n, h->h_counts, sizeof(int) * batch, cudaMemcpyHostToDevice, h->stream));
    for (int s = 0; s < batch; ++s)
        if (n[s] > 0)
            PF_CUDA(cudaMemcpyAsync(h->d_pts + (size_t)s * stride, xyzi + (size_t)s * stride * 4, (size_t)n[s] * 16,
                                    cudaMemcpyHostToDevice, h->stream));
    PF_CHECK(extract_launch(h, h->d_pts, h->d_n, batch, stride, h->d_edge, h->d_n_edge, edge_stride, h->d_surf, h->d_n_surf,
                            label ? h->d_label : nullptr));
    int* hc = h->h_counts + h->max_batch;
    PF_CUDA(cudaMemcpyAsync(hc, h->d_n_edge, sizeof(int) * batch, cudaMemcpyDeviceToHost, h->stream));
    PF_CUDA(cudaMemcpyAsync(hc + h->max_batch, h->d_n_surf, sizeof(int) * batch, cudaMemcpyDeviceToHost, h->stream));
    PF_CHECK(extract_check_ctrl(h));   // synchronises
    for (int s = 0; s < batch; ++s) {
        n_edge[s] = hc[s];
        n_surf[s] = hc[h->max_batch + s];
        if (n_edge[s] > 0)
            PF_CUDA(cudaMemcpyAsync(edge + (size_t)s * edge_stride * 4, h->d_edge + (size_t)s * edge_stride, (size_t)n_edge[s] * 16,
                                    cudaMemcpyDeviceToHost, h->stream));
        if (n_surf[s] > 0)
            PF_CUDA(cudaMemcpyAsync(surf + (size_t)s * stride * 4, h->d_surf + (size_t)s * stride, (size_t)n_surf[s] * 16,
                                    cudaMemcpyDeviceToHost, h->stream));
        if (label && n[s] > 0)
            PF_CUDA(cudaMemcpyAsync(label + (size_t)s * stride, h->d_label + (size_t)s * stride, (size_t)n[s], cudaMemcpyDeviceToHost,
                                    h->stream));
    }
    PF_CUDA(cudaStreamSynchronize(h->stream));
    return PF_OK;
}

extern "C" int pf_extract_run(pf_extract* h, const float* xyzi, int n, float* edge, int* n_edge, float* surf, int* n_surf,
                              uint8_t* label) {
    PF_REQUIRE(h && edge && n_edge && surf && n_surf, "null argument");
    PF_CHECK(pf_extract_enqueue_single(h, xyzi, n, 0));
    int* hc = h->h_counts + h->max_batch;
    PF_CUDA(cudaMemcpyAsync(hc, h->d_n_edge, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    PF_CUDA(cudaMemcpyAsync(hc + 1, h->d_n_surf, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    PF_CHECK(extract_check_ctrl(h));
    *n_edge = hc[0];
    *n_surf = hc[1];
    if (*n_edge > 0) PF_CUDA(cudaMemcpyAsync(edge, h->d_edge, (size_t)*n_edge * 16, cudaMemcpyDeviceToHost, h->stream));
    if (*n_surf > 0) PF_CUDA(cudaMemcpyAsync(surf, h->d_surf, (size_t)*n_surf * 16, cudaMemcpyDeviceToHost, h->stream));
    if (label && n > 0) PF_CUDA(cudaMemcpyAsync(label, h->d_label, (size_t)n, cudaMemcpyDeviceToHost, h->stream));
    PF_CUDA(cudaStreamSynchronize(h->stream));
    return PF_OK;
}
