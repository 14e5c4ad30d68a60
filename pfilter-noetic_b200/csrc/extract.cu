// K1: LiDAR feature extraction on sm_100a.
//
// Replaces LaserProcessingClass::featureExtraction / featureExtractionFromSector
// (/root/reference/src/laserProcessingClass.cpp:10-96, :99-209).  Arithmetic spec: SURVEY.md appendix A.1.
//
// Two kernels per batch of scans:
//   k_ring_classify : one thread per point, coalesced float4 loads; ring id (elevation angle, :25-61) -> 1 byte.  The
//                     angle is first evaluated in fp32 (polynomial atan); only points whose bin position lies within
//                     1e-3 of a decision boundary take the reference's exact double-precision atan path, so the ring
//                     ids are the reference's.  Per (scan, ring) the first / last 256-point tile holding the ring.
//   k_ring_extract  : one CTA (6 warps) per (scan, ring).  Stable gather of the ring's points (:62) from its tile
//                     range into shared memory; 11-tap float curvature (:73-80) from a 17-point register window per
//                     lane; "short step" link bits for the neighbour suppression (:128-145); then one warp per sector
//                     (:81-92): candidates (curvature > 0.1, :114) are compacted into registers as
//                     (23-bit monotone key | 9-bit index) words and the greedy descending walk of :110-148 becomes at
//                     most 21 rounds of { redux.max, range kill }; two candidates in the same key bucket are resolved
//                     with the exact double curvature (ties: higher index first, = the descending walk over an
//                     ascending sort with lower index first).  surf = unflagged (:198-205).  Output offsets of the
//                     rings of a scan are chained through a decoupled look-back on (epoch, counts) words, so the
//                     compacted edge/surf clouds are written once, straight from shared memory.
// HBM traffic per point: 16 B read + 16 B written (+ ring id / label bytes); the second read of the points by
// k_ring_extract hits L2.
#include <math.h>

#include <vector>

#include "common.cuh"

namespace pf {

constexpr int kTile = 256;          // points per classify tile
constexpr int kExtractThreads = 192;
constexpr int kExtractWarps = kExtractThreads / 32;
constexpr int kMaxLines = 64;
constexpr int kEdgePerSector = 20;  // src/laserProcessingClass.cpp:121
constexpr int kSectors = 6;         // :81
constexpr int kWin = 5;             // curvature values per lane per pass (odd: conflict-free 16-byte shared loads)
constexpr int kMaxRingCap = 3040;   // sector length <= 512 (9-bit index in the candidate word)
static_assert(kExtractWarps == kSectors, "one warp per sector");

struct ExtractParams {
    const float4* pts;        // [batch][stride]
    const int* n;             // [batch]
    uint8_t* ringid;          // [batch][stride]
    int2* ring_tiles;         // [batch][64] first / last tile holding the ring (reset to {INT_MAX, -1} by the consumer)
    uint8_t* tile_pure;       // [batch][tiles] ring id when the tile is a full tile of one ring, else 255
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

// ring id of one point or 255 (dropped); src/laserProcessingClass.cpp:25-61, exact arithmetic
__device__ __noinline__ int ring_id_exact(float x, float y, float z, int num_lines, double min_d, double max_d) {
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

// fp32 evaluation of the same decision; returns -1 when the point is too close to a decision boundary to call.
// atan(t) = t P(t^2) on |t| <= 0.75 (least-squares fit, |error| < 6e-6 deg); the bin position u is then known to
// better than 1e-4, and anything within kEps of an integer / a gate goes to the exact path.
__device__ __forceinline__ int ring_id_fast_t(float t, int num_lines) {
    constexpr float kEps = 1e-3f;
    if (!(fabsf(t) <= 0.75f)) return -1;
    const float u2 = __fmul_rn(t, t);
    float p = 0x1.2dafa8p-6f;
    p = __fmaf_rn(p, u2, -0x1.df9d38p-5f);
    p = __fmaf_rn(p, u2, 0x1.9c48a2p-4f);
    p = __fmaf_rn(p, u2, -0x1.20c29ap-3f);
    p = __fmaf_rn(p, u2, 0x1.9943ecp-3f);
    p = __fmaf_rn(p, u2, -0x1.5553e4p-2f);
    p = __fmaf_rn(p, u2, 0x1.fffffep-1f);
    const float ang = __fmul_rn(__fmul_rn(p, t), 57.29577951308232f);
    float u;
    int base = 0;
    if (num_lines == 64) {
        if (fabsf(ang + 8.83f) < kEps) return -1;
        if (ang >= -8.83f) {
            if (ang > 2.0f - kEps) return ang > 2.0f + kEps ? 255 : -1;
            u = __fmaf_rn(2.0f - ang, 3.0f, 0.5f);
        } else {
            if (ang < -24.33f + kEps) return ang < -24.33f - kEps ? 255 : -1;
            u = __fmaf_rn(-8.83f - ang, 2.0f, 0.5f);
            base = 32;
        }
    } else if (num_lines == 32) {
        u = __fmul_rn(ang + 30.666666f, 0.75f);
        if (u < kEps) return u < -1.0f - kEps ? 255 : -1;    // int() truncates towards zero: (-1, 0] is ring 0
    } else {
        u = __fmaf_rn(ang + 15.0f, 0.5f, 0.5f);
        if (u < kEps) return u < -1.0f - kEps ? 255 : -1;
    }
    const float fl = floorf(u), fr = u - fl;
    if (fr < kEps || fr > 1.0f - kEps) return -1;
    const int id = base + (int)fl;
    return id < num_lines ? id : 255;
}

// fp32 fast path on top of one MUFU.RSQ: dist = s * rsqrt(s) and z / dist = z * rsqrt(s) are good to ~3 ulp, i.e. the bin position is
// still known to ~1e-4 (ring_id_fast calls everything within 1e-3 of a decision "unsure"); the range gate is pulled in by 1e-6
// relative so that the approximate distance can never decide a point the reference's (double)sqrtf comparison would not.
__device__ __forceinline__ int ring_id_dev(float x, float y, float z, int num_lines, double min_d, double max_d, float gate_lo,
                                           float gate_hi) {
    const float s = __fmaf_rn(x, x, __fmul_rn(y, y));
    const float r = rsqrtf(s);
    const float dist = __fmul_rn(s, r);
    int id = -1;
    if (dist > gate_lo && dist < gate_hi && fabsf(z) < 3.0e38f) id = ring_id_fast_t(__fmul_rn(z, r), num_lines);
    if (id < 0) {
        if (!(isfinite(x) && isfinite(y) && isfinite(z))) return 255;   // x86 int(NaN) = INT_MIN: the reference drops these
        id = ring_id_exact(x, y, z, num_lines, min_d, max_d);
    }
    return id;
}

__global__ void __launch_bounds__(kTile) k_ring_classify(ExtractParams P, float gate_lo, float gate_hi) {
    const int s = blockIdx.y, tile = blockIdx.x, tid = threadIdx.x;
    if (blockIdx.x == 0 && blockIdx.y == 0 && tid == 0) {
        P.ctrl[0] = 0;                // ticket for the extract kernel that follows in stream order
        atomicAdd(&P.ctrl[1], 1u);    // new epoch invalidates all look-back words of earlier launches
    }
    const int n = P.n[s];
    if (tile * kTile >= n) return;
    __shared__ int s_wring[kTile / 32];
    __shared__ unsigned s_lo[kTile / 32], s_hi[kTile / 32], s_drop[kTile / 32];
    const int i = tile * kTile + tid;
    const size_t g = (size_t)s * P.stride + i;
    int ring = 255;
    if (i < n) {
        float4 p = ld_stream_f4(P.pts + g);
        ring = ring_id_dev(p.x, p.y, p.z, P.num_lines, P.min_d, P.max_d, gate_lo, gate_hi);
        if (P.label) P.label[g] = 0;
    }
    P.ringid[g] = (uint8_t)ring;
    // rings present in this tile -> per (scan, ring) tile range; a full tile of one ring is "pure" (copied without ballots).
    // Common case first (ring-major input: 3 tiles out of 4 hold a single ring): one match per warp, one atomic pair per tile.
    int same;
    __match_all_sync(0xffffffffu, ring, &same);
    if ((tid & 31) == 0) s_wring[tid >> 5] = same ? ring : 256;
    __syncthreads();
    const int r0 = s_wring[0];
    bool pure = r0 < kMaxLines;
#pragma unroll
    for (int w = 1; w < kTile / 32; ++w) pure = pure && s_wring[w] == r0;
    if (pure) {      // uniform over the CTA
        if (tid == 0) {
            int2* rt = P.ring_tiles + (size_t)s * kMaxLines + r0;
            atomicMin(&rt->x, tile);
            atomicMax(&rt->y, tile);
            P.tile_pure[(size_t)s * P.tiles + tile] = (uint8_t)r0;
        }
        return;
    }
    const unsigned bit = 1u << (ring & 31);
    const unsigned lo = __reduce_or_sync(0xffffffffu, ring < 32 ? bit : 0u);
    const unsigned hi = __reduce_or_sync(0xffffffffu, (ring >= 32 && ring < 64) ? bit : 0u);
    if ((tid & 31) == 0) { s_lo[tid >> 5] = lo; s_hi[tid >> 5] = hi; }
    __syncthreads();
    if (tid < kMaxLines) {
        unsigned ml = 0u, mh = 0u;
#pragma unroll
        for (int w = 0; w < kTile / 32; ++w) { ml |= s_lo[w]; mh |= s_hi[w]; }
        const unsigned m = tid < 32 ? ml : mh;
        if (m >> (tid & 31) & 1u) {
            int2* rt = P.ring_tiles + (size_t)s * kMaxLines + tid;
            atomicMin(&rt->x, tile);
            atomicMax(&rt->y, tile);
        }
        if (tid == 0) P.tile_pure[(size_t)s * P.tiles + tile] = (uint8_t)255;
    }
}

__global__ void k_set_int(int* p, int v) {
    if (threadIdx.x == 0) *p = v;
}

// exclusive scan of a[0..n) in shared memory by the whole CTA; returns the total
__device__ int block_excl_scan(int* a, int n, int* warp_tmp /*[kExtractWarps]*/) {
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
        for (int k = 0; k < kExtractWarps; ++k) {
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

// exact curvature of ring element j (:73-77): float sums left to right, squares and their sum in double
__device__ __noinline__ double curvature_exact(const float4* sp, int j) {
    double sq[3];
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
        const double d = (double)t;
        sq[a] = __dmul_rn(d, d);
    }
    return __dadd_rn(__dadd_rn(sq[0], sq[1]), sq[2]);
}

// Greedy edge pick of one sector [a, b) by one warp (:110-148); NPL candidate words per lane.  Returns the pick count.
// A candidate word is (23-bit key << 9) | (index - a): the maximum word is the next element of the descending walk unless
// another live candidate shares its key bucket; then the exact double values decide (ties: higher index first).
template <int NPL>
__device__ __forceinline__ int sector_pick(const float4* sp, const unsigned* scand, int C, const uint8_t* slinkb, uint8_t* sflag,
                                           int* edge_ids, int a, int b, int lane) {
    unsigned word[NPL];
#pragma unroll
    for (int k = 0; k < NPL; ++k) word[k] = (k * 32 + lane < C) ? scand[k * 32 + lane] : 0u;
    // lanes holding two words of one bucket (rare) force the exact comparison whenever their bucket is on top
    bool dup = false;
#pragma unroll
    for (int k = 0; k < NPL; ++k)
#pragma unroll
        for (int l = k + 1; l < NPL; ++l) dup |= (word[k] != 0u) && ((word[k] ^ word[l]) >> 9) == 0u;
    int cnt = 0;
    while (true) {
        unsigned m = word[0];
#pragma unroll
        for (int k = 1; k < NPL; ++k) m = max(m, word[k]);
        const unsigned M = __reduce_max_sync(0xffffffffu, m);
        if (M == 0u) break;
        const unsigned bucket = M >> 9;
        const unsigned top = __ballot_sync(0xffffffffu, (m >> 9) == bucket);
        int rel = (int)(M & 511u);
        if ((top & (top - 1u)) != 0u || __any_sync(0xffffffffu, dup && (m >> 9) == bucket)) {
            unsigned long long bk = 0ull;
            int bi = -1;
#pragma unroll
            for (int k = 0; k < NPL; ++k) {
                if ((word[k] >> 9) == bucket) {
                    const int r = (int)(word[k] & 511u);
                    const unsigned long long v = (unsigned long long)__double_as_longlong(curvature_exact(sp, a + r));
                    if (bi < 0 || v > bk || (v == bk && r > bi)) { bk = v; bi = r; }
                }
            }
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) {
                const unsigned long long ok = __shfl_xor_sync(0xffffffffu, bk, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (oi >= 0 && (bi < 0 || ok > bk || (ok == bk && oi > bi))) { bk = ok; bi = oi; }
            }
            rel = bi;
        }
        const int i = a + rel;
        cnt++;                                                                      // :118-119
        if (cnt > kEdgePerSector) {                                                 // :121-126 the 21st is picked, not emitted
            if (lane == 0) sflag[i] = 1;
            break;
        }
        if (lane == 0) edge_ids[cnt - 1] = i;
        // neighbour suppression (:128-145): forward while the steps i->i+1->... are short (at most 5), same backward.
        // lanes 0..9 fetch the step bytes i-5 .. i+4
        const unsigned steps = __ballot_sync(0xffffffffu, slinkb[i - 5 + min(lane, 9)] != 0) & 0x3ffu;
        const int fwd = __ffs(~(steps >> 5) | 32u) - 1;
        const int back = __clz(~steps & 31u) - 27;
        const int lo = max(i - back, a), hi = min(i + fwd, b - 1);
        if (lane <= hi - lo) sflag[lo + lane] = 1;
        const unsigned rlo = (unsigned)(lo - a), span = (unsigned)(hi - lo);
#pragma unroll
        for (int k = 0; k < NPL; ++k)
            if ((word[k] & 511u) - rlo <= span) word[k] = 0u;
    }
    return cnt;
}

__global__ void __launch_bounds__(kExtractThreads, 4) k_ring_extract(ExtractParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* sp = reinterpret_cast<float4*>(smem_raw);                       // [rcap + 16] ring points
    unsigned* scand = reinterpret_cast<unsigned*>(sp + P.rcap + 16);        // [rcap] candidate words, one list per sector
    int* tl_off = reinterpret_cast<int*>(scand + P.rcap);                   // [maxtl + 1]
    uint8_t* sflag = reinterpret_cast<uint8_t*>(tl_off + P.maxtl + 1);      // [rcap] picked / suppressed
    uint8_t* slinkb = sflag + P.rcap;                                       // [rcap + 16] short step q -> q+1
    int* ssrc = reinterpret_cast<int*>(slinkb + P.rcap + 16);               // [rcap] source index (label output only)

    __shared__ int s_work;
    __shared__ int s_warp[kExtractWarps];
    __shared__ int s_edge_ids[kSectors][kEdgePerSector];
    __shared__ int s_ecnt[kSectors], s_scnt[kSectors], s_ccnt[kSectors];
    __shared__ int s_eoff, s_soff;

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    constexpr int NW = kExtractWarps;

    // work item by ticket: CTAs that hold ticket t only ever wait on tickets < t, which are already running
    if (tid == 0) s_work = (int)atomicAdd(&P.ctrl[0], 1u);
    if (tid < kSectors) s_ccnt[tid] = 0;
    __syncthreads();
    // ring-major over the batch: ring r of a scan starts a whole wave of CTAs after its rings 0..r-1, so by the time it looks
    // back for their counts they have long been published (scan-major order made ring 63 wait for 63 siblings started with it)
    const int s = s_work % P.batch, r = s_work / P.batch;
    const unsigned epoch = *reinterpret_cast<volatile unsigned int*>(&P.ctrl[1]);
    const int n = P.n[s];
    const uint8_t* rid = P.ringid + (size_t)s * P.stride;
    const float4* pts = P.pts + (size_t)s * P.stride;
    const uint8_t* pure = P.tile_pure + (size_t)s * P.tiles;
    const bool want_label = P.label != nullptr;

    // A. tile range of this ring (from the classify kernel); hand the slot back for the next launch
    int2* rtp = P.ring_tiles + (size_t)s * kMaxLines + r;
    const int2 rt = *rtp;
    __syncthreads();
    if (tid == 0) *rtp = make_int2(0x7fffffff, -1);
    const int tlo = rt.x;
    int ncand = rt.y >= rt.x ? rt.y - rt.x + 1 : 0;
    if (ncand > P.maxtl) ncand = P.maxtl;   // cannot happen (maxtl = tiles per scan)

    // B. matches per tile (a pure tile holds 256 points of one ring)
    for (int c = w; c < ncand; c += NW) {
        const int tbase = (tlo + c) * kTile;
        const int pr = pure[tlo + c];
        int cnt = 0;
        if (pr == r) {
            cnt = min(kTile, n - tbase);
        } else if (pr == 255) {
#pragma unroll
            for (int k = 0; k < kTile / 32; ++k) {
                int i = tbase + k * 32 + lane;
                int b = i < n ? rid[i] : 255;
                cnt += __popc(__ballot_sync(0xffffffffu, b == r));
            }
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
            const int tbase = (tlo + c) * kTile;
            const int pr = pure[tlo + c];
            int pos = tl_off[c];
            if (pr == r) {
                const int cnt = min(kTile, n - tbase);
                float4 v[kTile / 32];
#pragma unroll
                for (int k = 0; k < kTile / 32; ++k)
                    if (k * 32 + lane < cnt) v[k] = __ldg(pts + tbase + k * 32 + lane);
#pragma unroll
                for (int k = 0; k < kTile / 32; ++k) {
                    if (k * 32 + lane < cnt) {
                        sp[pos + k * 32 + lane] = v[k];
                        if (want_label) ssrc[pos + k * 32 + lane] = tbase + k * 32 + lane;
                    }
                }
            } else if (pr == 255) {
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
                        if (want_label) ssrc[d] = tbase + k * 32 + lane;
                    }
                    pos += __popc(m[k]);
                }
            }
        }
        // flags of the ring, cleared by words
        for (int q = tid; q < (nr + 3) / 4; q += kExtractThreads) reinterpret_cast<unsigned*>(sflag)[q] = 0u;
        __syncthreads();

        // D. curvature (:73-80) and short-step bytes (:129-132, :138-141) from a register window: a lane owns kWin
        //    consecutive ring positions per pass; candidates (value > 0.1, :114) go straight into their sector's list.
        const int total = nr - 10, L = total / kSectors;
        const float inv_l = 1.0f / (float)L;
        for (int base = 5 + w * (32 * kWin); base < nr - 1; base += NW * 32 * kWin) {
            const int j0 = base + lane * kWin;
            if (j0 < nr - 1) {
                float wx[kWin + 10], wy[kWin + 10], wz[kWin + 10];
#pragma unroll
                for (int k = 0; k < kWin + 10; ++k) {
                    const float4 p = sp[j0 - 5 + k];
                    wx[k] = p.x; wy[k] = p.y; wz[k] = p.z;
                }
                // short steps q -> q+1 for the own positions (and 0..4 by the lane that owns position 5): float differences,
                // double squares, decided in fp32 unless within 1e-6 of the threshold 0.05
                auto short_step = [&](int k) {   // window slots k, k+1
                    const float dx = __fsub_rn(wx[k + 1], wx[k]), dy = __fsub_rn(wy[k + 1], wy[k]), dz = __fsub_rn(wz[k + 1], wz[k]);
                    const float sf = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
                    if (sf < 0.04999995f) return true;
                    if (sf > 0.05000005f) return false;
                    const double ddx = dx, ddy = dy, ddz = dz;
                    return !(__dadd_rn(__dadd_rn(__dmul_rn(ddx, ddx), __dmul_rn(ddy, ddy)), __dmul_rn(ddz, ddz)) > 0.05);
                };
#pragma unroll
                for (int k = 0; k < kWin; ++k)
                    if (j0 + k + 1 < nr) slinkb[j0 + k] = short_step(k + 5) ? 1 : 0;
                if (j0 == 5) {
#pragma unroll
                    for (int k = 0; k < 5; ++k) slinkb[k] = short_step(k) ? 1 : 0;
                }
                int t = j0 - 5;
                int sec = min(kSectors - 1, __float2int_rd(((float)t + 0.5f) * inv_l));
                int bound = L * (sec + 1);
#pragma unroll
                for (int k = 0; k < kWin; ++k, ++t) {
                    auto tap = [&](const float* a) {
                        float u = __fadd_rn(a[k], a[k + 1]);
                        u = __fadd_rn(u, a[k + 2]);
                        u = __fadd_rn(u, a[k + 3]);
                        u = __fadd_rn(u, a[k + 4]);
                        u = __fsub_rn(u, __fmul_rn(10.0f, a[k + 5]));
                        u = __fadd_rn(u, a[k + 6]);
                        u = __fadd_rn(u, a[k + 7]);
                        u = __fadd_rn(u, a[k + 8]);
                        u = __fadd_rn(u, a[k + 9]);
                        u = __fadd_rn(u, a[k + 10]);
                        return (double)u;
                    };
                    if (sec < kSectors - 1 && t >= bound) { ++sec; bound += L; }
                    // sector slice [L sec, hi) with hi = L (sec+1) - 1, or total - 1 for the last: hi itself is dropped (:83-88)
                    const int hi = sec == kSectors - 1 ? total - 1 : bound - 1;
                    if (t < hi) {
                        const double dx = tap(wx), dy = tap(wy), dz = tap(wz);
                        const double val = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                        if (val > 0.1) {
                            // key monotone in the exact value (round towards zero), index relative to the sector start
                            const unsigned key = __float_as_uint(__double2float_rz(val)) >> 8;
                            const int pos = atomicAdd(&s_ccnt[sec], 1);
                            scand[L * sec + 5 + pos] = (key << 9) | (unsigned)(t - L * sec);
                        }
                    }
                }
            }
        }
        __syncthreads();

        // E. one warp per sector (:81-92, :99-209)
        {
            const int lo = L * w, hi = (w == kSectors - 1) ? total - 1 : L * (w + 1) - 1;   // hi excluded (:83-88)
            const int a = lo + 5, b = hi + 5;
            const int C = s_ccnt[w];
            int cnt;
            const unsigned* sc = scand + a;
            if (C <= 32) cnt = sector_pick<1>(sp, sc, C, slinkb, sflag, s_edge_ids[w], a, b, lane);
            else if (C <= 64) cnt = sector_pick<2>(sp, sc, C, slinkb, sflag, s_edge_ids[w], a, b, lane);
            else if (C <= 128) cnt = sector_pick<4>(sp, sc, C, slinkb, sflag, s_edge_ids[w], a, b, lane);
            else if (C <= 256) cnt = sector_pick<8>(sp, sc, C, slinkb, sflag, s_edge_ids[w], a, b, lane);
            else cnt = sector_pick<16>(sp, sc, C, slinkb, sflag, s_edge_ids[w], a, b, lane);
            __syncwarp();
            int nsurf = 0;
            for (int base = a; base < b; base += 32) {
                int i = base + lane;
                nsurf += __popc(__ballot_sync(0xffffffffu, i < b && !sflag[i]));
            }
            if (lane == 0) { s_ecnt[w] = min(cnt, kEdgePerSector); s_scnt[w] = nsurf; }
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
            unsigned long long v = ld_acquire_u64(done + q);
            while ((unsigned)(v >> 32) != epoch) {
                __nanosleep(200);
                v = ld_acquire_u64(done + q);
            }
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
    {
        const int total = nr - 10, L = total / kSectors;
        const int lo = L * w, hi = (w == kSectors - 1) ? total - 1 : L * (w + 1) - 1;
        const int a = lo + 5, b = hi + 5;
        int eo = s_eoff, so = s_soff;
        for (int k = 0; k < w; ++k) { eo += s_ecnt[k]; so += s_scnt[k]; }
        uint8_t* label = want_label ? P.label + (size_t)s * P.stride : nullptr;
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
    int2* d_ring_tiles = nullptr;
    uint8_t* d_tile_pure = nullptr;
    float gate_lo = 0, gate_hi = 0;
    size_t smem_label = 0;
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
    // result of the last single-scan run (device resident hand-off to the odometry).  Single-scan outputs are double
    // buffered (slot toggles per scan) so that the extraction of the next frame can run while the odometry still reads this one.
    int last_valid = 0;
    int last_n = 0;
    int slot = 0;
    float4* d_edge1 = nullptr;       // slot 1 (slot 0 = d_edge / d_surf / d_n_edge / d_n_surf, shared with the batched path)
    float4* d_surf1 = nullptr;
    int* d_cnt1 = nullptr;           // [2] n_edge, n_surf of slot 1
    float4* out_edge() const { return slot ? d_edge1 : d_edge; }
    float4* out_surf() const { return slot ? d_surf1 : d_surf; }
    int* out_n_edge() const { return slot ? d_cnt1 : d_n_edge; }
    int* out_n_surf() const { return slot ? d_cnt1 + 1 : d_n_surf; }
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
        P.ring_tiles = h->d_ring_tiles + (size_t)s0 * kMaxLines;
        P.tile_pure = h->d_tile_pure + (size_t)s0 * tiles;
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
        k_ring_classify<<<dim3(tiles, nb), kTile, 0, h->stream>>>(P, h->gate_lo, h->gate_hi);
        k_ring_extract<<<nb * h->lidar.num_lines, kExtractThreads, d_label ? h->smem_label : h->smem, h->stream>>>(P);
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
    h->rcap = cfg->max_ring_points > 0 ? cfg->max_ring_points : 2304;
    h->rcap = div_up(h->rcap, 32) * 32;
    if (h->rcap > kMaxRingCap) {
        set_error("max_ring_points %d exceeds the supported %d", h->rcap, kMaxRingCap);
        delete h;
        return PF_ERR_INVALID;
    }
    h->maxtl = h->tiles;
    h->edge_stride = 120 * lidar->num_lines;
    // shared memory of k_ring_extract: points, keys, link bits, tile offsets, flags (+ source indices for the label output)
    h->smem = (size_t)(h->rcap + 16) * 16 + (size_t)h->rcap * 4 + (size_t)(h->maxtl + 1) * 4 + h->rcap + (h->rcap + 16);
    h->smem = (h->smem + 15) / 16 * 16;
    h->smem_label = h->smem + (size_t)h->rcap * 4;
    if (h->smem_label > (size_t)prop.sharedMemPerBlockOptin) {
        set_error("max_ring_points %d / max_points %d need %zu B shared memory (> %zu)", h->rcap, cfg->max_points, h->smem_label,
                  (size_t)prop.sharedMemPerBlockOptin);
        delete h;
        return PF_ERR_INVALID;
    }
    PF_CUDA(cudaFuncSetAttribute(k_ring_extract, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_label));
    // fp32 range gate strictly inside [min_distance, max_distance]: everything else takes the exact double comparison
    h->gate_lo = (float)lidar->min_distance;
    if (!((double)h->gate_lo > lidar->min_distance)) h->gate_lo = nextafterf(h->gate_lo, INFINITY);
    h->gate_hi = (float)lidar->max_distance;
    if (!((double)h->gate_hi < lidar->max_distance)) h->gate_hi = nextafterf(h->gate_hi, -INFINITY);
    // the fast path sees the distance only to a few ulp (rsqrt): pull both limits in by 1e-6 relative
    h->gate_lo = h->gate_lo * (1.0f + 1e-6f) + 1e-30f;
    h->gate_hi = h->gate_hi * (1.0f - 1e-6f);
    // scans per launch pair.  Measured on B200 (128 scans of 114 k points): one pair for the whole batch is fastest (0.444 ms
    // vs 0.605 ms in L2-sized groups of 21): the kernels are bound by instruction issue, not by DRAM, so the second read of the
    // points missing L2 costs less than the extra launches and partial waves of small groups.  PF_EXTRACT_GROUP overrides.
    {
        int g = h->max_batch;
        const char* env = getenv("PF_EXTRACT_GROUP");
        if (env) g = atoi(env);
        h->group = g < 1 ? 1 : g;
    }
    PF_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    const size_t np = (size_t)h->max_batch * h->stride;
    PF_CUDA(cudaMalloc(&h->d_pts, np * 16));
    PF_CUDA(cudaMalloc(&h->d_ringid, np));
    PF_CUDA(cudaMalloc(&h->d_ring_tiles, sizeof(int2) * (size_t)h->max_batch * kMaxLines));
    PF_CUDA(cudaMalloc(&h->d_tile_pure, (size_t)h->max_batch * h->tiles));
    {
        std::vector<int2> init((size_t)h->max_batch * kMaxLines, make_int2(0x7fffffff, -1));
        PF_CUDA(cudaMemcpy(h->d_ring_tiles, init.data(), sizeof(int2) * init.size(), cudaMemcpyHostToDevice));
    }
    PF_CUDA(cudaMalloc(&h->d_label, np));
    PF_CUDA(cudaMalloc(&h->d_edge, (size_t)h->max_batch * h->edge_stride * 16));
    PF_CUDA(cudaMalloc(&h->d_surf, np * 16));
    PF_CUDA(cudaMalloc(&h->d_n, sizeof(int) * h->max_batch));
    PF_CUDA(cudaMalloc(&h->d_edge1, (size_t)h->edge_stride * 16));
    PF_CUDA(cudaMalloc(&h->d_surf1, (size_t)h->stride * 16));
    PF_CUDA(cudaMalloc(&h->d_cnt1, sizeof(int) * 2));
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
    cudaFree(h->d_pts); cudaFree(h->d_ringid); cudaFree(h->d_ring_tiles); cudaFree(h->d_tile_pure); cudaFree(h->d_label); cudaFree(h->d_edge);
    cudaFree(h->d_surf); cudaFree(h->d_n); cudaFree(h->d_n_edge); cudaFree(h->d_n_surf); cudaFree(h->d_done);
    cudaFree(h->d_edge1); cudaFree(h->d_surf1); cudaFree(h->d_cnt1);
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
int pf_extract_enqueue_single(pf_extract* h, const float* xyzi, int n, int device_input, int want_label) {
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
    h->slot ^= 1;
    PF_CHECK(extract_launch(h, src ? src : h->d_pts, h->d_n, 1, h->stride, h->out_edge(), h->out_n_edge(), h->edge_stride, h->out_surf(),
                            h->out_n_surf(), want_label ? h->d_label : nullptr));
    h->last_valid = 1;
    h->last_n = n;
    return PF_OK;
}

// accessors for the device-resident hand-off (odom.cu)
void pf_extract_device_outputs(pf_extract* h, const float4** edge, const int** n_edge, const float4** surf, const int** n_surf,
                               cudaStream_t* stream, int* edge_cap, int* surf_cap, int* slot) {
    *edge = h->out_edge(); *n_edge = h->out_n_edge(); *surf = h->out_surf(); *n_surf = h->out_n_surf(); *stream = h->stream;
    *slot = h->slot;
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
    PF_CHECK(pf_extract_enqueue_single(h, xyzi, n, 0, label != nullptr));
    int* hc = h->h_counts + h->max_batch;
    PF_CUDA(cudaMemcpyAsync(hc, h->out_n_edge(), sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    PF_CUDA(cudaMemcpyAsync(hc + 1, h->out_n_surf(), sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    PF_CHECK(extract_check_ctrl(h));
    *n_edge = hc[0];
    *n_surf = hc[1];
    if (*n_edge > 0) PF_CUDA(cudaMemcpyAsync(edge, h->out_edge(), (size_t)*n_edge * 16, cudaMemcpyDeviceToHost, h->stream));
    if (*n_surf > 0) PF_CUDA(cudaMemcpyAsync(surf, h->out_surf(), (size_t)*n_surf * 16, cudaMemcpyDeviceToHost, h->stream));
    if (label && n > 0) PF_CUDA(cudaMemcpyAsync(label, h->d_label, (size_t)n, cudaMemcpyDeviceToHost, h->stream));
    PF_CUDA(cudaStreamSynchronize(h->stream));
    return PF_OK;
}
