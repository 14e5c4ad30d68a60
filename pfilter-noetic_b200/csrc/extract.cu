// K1: LiDAR feature extraction on sm_100a.
//
// Replaces LaserProcessingClass::featureExtraction / featureExtractionFromSector
// (/root/reference/src/laserProcessingClass.cpp:10-96, :99-209).  Arithmetic spec: SURVEY.md appendix A.1.
//
// Three kernels per batch of scans:
//   k_ring_classify  : one warp per 256-point tile, coalesced float4 loads; ring id (elevation angle, :25-61) -> 1 byte.  The
//                      angle is first evaluated in fp32 (polynomial atan); only points whose bin position lies within
//                      2.5e-4 of a decision boundary take the reference's exact double-precision atan path, so the ring
//                      ids are the reference's.  Per (scan, ring) the first / last tile holding the ring; per tile whether
//                      it holds one ring only ("pure": the ring-major order LiDAR drivers emit makes most tiles pure).
//   k_ring_index     : one warp per (scan, ring): ring position at which each tile of the ring's range starts, ring size.
//   k_sector_extract : one WARP per (scan, ring, sector) (:81-92), persistent, work by ticket, no CTA barrier.  Stable gather of
//                      the sector's points (+5 either side, :62) into the warp's slice of shared memory; 11-tap float
//                      curvature (:73-80) from a 20-point register window per lane; "short step" link bits for the
//                      neighbour suppression (:128-145); candidates (curvature > 0.1, :114) are compacted into registers as
//                      (23-bit monotone key | 9-bit index) words and the greedy descending walk of :110-148 becomes at
//                      most 21 rounds of { redux.max, flag-range update, range kill }; two candidates in the same key
//                      bucket are resolved with the exact double curvature (ties: higher index first, = the descending
//                      walk over an ascending sort with lower index first).  surf = unflagged (:198-205).  Output offsets
//                      are chained through (epoch, counts) words per sector and per ring, so the compacted edge / surf
//                      clouds are written once, straight from shared memory.
// HBM traffic per point: 16 B read + 16 B written (+ ring id / label bytes); the second read of the points by
// k_sector_extract hits L2 when the batch fits there.
#include <math.h>

#include <algorithm>
#include <atomic>
#include <vector>

#include "common.cuh"

namespace pf {

constexpr int kTile = 256;          // points per classify tile (one warp)
constexpr int kClassifyThreads = 256;
constexpr int kSecWarps = 2;        // warps per CTA of k_sector_extract; every warp works on its own
constexpr int kMaxLines = 64;
constexpr int kEdgePerSector = 20;  // src/laserProcessingClass.cpp:121
constexpr int kSectors = 6;         // :81
constexpr int kWin = 10;            // curvature positions per lane per pass
constexpr int kMaxRingCap = 3040;   // sector length <= 512 (9-bit index in the candidate word)
constexpr unsigned kFull = 0xffffffffu;

struct ExtractParams {
    const float4* pts;        // [batch][stride]
    const int* n;             // [batch]
    uint8_t* ringid;          // [batch][stride]
    int2* ring_tiles;         // [batch][64] first / last tile holding the ring (reset to {INT_MAX, -1} by k_ring_index)
    uint8_t* tile_pure;       // [batch][tiles] ring id when the tile is a full tile of one ring, else 255
    int* tile_off;            // [batch][64][tiles + 1] ring positions at which the tiles of the ring's range start
    int4* ring_info;          // [batch][64] {first tile, tiles in range, points of the ring, -}
    int4* sec_desc;           // [batch][64][6] {local points (0 = ring not processed), first ring position, first source index when
                              //                 the sector is one contiguous run of the input (else -1), first tile of the range}
    uint8_t* label;           // [batch][stride] or null
    float4* edge;             // [batch][edge_stride]
    float4* surf;             // [batch][stride]
    int* n_edge;              // [batch]
    int* n_surf;              // [batch]
    unsigned long long* done_sec;   // [batch][64 * 6] (epoch | edges | surfs) of a sector
    unsigned long long* done_ring;  // [batch][64]     the same summed over the ring, published by its last sector
    unsigned int* ctrl;       // [0] ticket, [1] epoch, [2] error bits
    int stride, tiles, edge_stride, batch;
    int num_lines, rcap, scap, warp_smem;
    int sval_off;             // byte offset of the exact-curvature array inside a warp's shared memory (reference surf order only)
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

// fp32 evaluation of the same decision, straight-line; returns -1 when the point is too close to a decision boundary to call.
// atan(t) = t P(t^2) on |t| <= 0.75 (least-squares fit, |error| < 6e-6 deg).  Error budget of the bin position u (bins are 1/3,
// 1/2, 4/3 or 2 degrees wide): t is known to ~4e-7 relative (fma, MUFU.RSQ at 2 ulp, one product) = 1e-5 deg at 25 deg, the
// polynomial adds 6e-6 deg and its evaluation ~5e-6 deg: < 7e-5 bins in the worst case (3 bins per degree); anything within
// kEps = 2.5e-4 of an integer or within 4 kEps of a validity gate goes to the exact path.
template <int LINES>
__device__ __forceinline__ int ring_id_fast_t(float t) {
    constexpr float kEps = 2.5e-4f;
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
    bool drop = false, unsure = !(fabsf(t) <= 0.75f);
    if (LINES == 64) {
        // upper block (angle >= -8.83): u = (2 - a) 3 + 0.5, valid while a <= 2 (u >= 0.5); lower block: u = (-8.83 - a) 2 + 0.5,
        // valid while a >= -24.33 (u <= 31.5).  Either side of the split gives ring 32, away from any integer of u.
        const bool lower = ang < -8.83f;
        u = lower ? __fmaf_rn(ang, -2.0f, -17.16f) : __fmaf_rn(ang, -3.0f, 6.5f);
        base = lower ? 32 : 0;
        drop = lower ? u > 31.5f : u < 0.5f;
        unsure = unsure || fabsf(u - (lower ? 31.5f : 0.5f)) < 4.0f * kEps;
    } else if (LINES == 32) {
        u = __fmul_rn(ang + 30.666666f, 0.75f);
        drop = u < -1.0f;                               // int() truncates towards zero: (-1, 0] is ring 0, left to the exact path
        unsure = unsure || fabsf(u + 0.5f) < 0.5f + kEps;
    } else {
        u = __fmaf_rn(ang + 15.0f, 0.5f, 0.5f);
        drop = u < -1.0f;
        unsure = unsure || fabsf(u + 0.5f) < 0.5f + kEps;
    }
    const float fl = floorf(u), fr = u - fl;
    unsure = unsure || fabsf(fr - 0.5f) > 0.5f - kEps;
    int id = base + (int)fl;
    id = (drop || id >= LINES) ? 255 : id;
    return unsure ? -1 : id;
}

// fp32 fast path on top of one MUFU.RSQ: dist = s * rsqrt(s) and z / dist = z * rsqrt(s) are good to ~3 ulp; the range gate is
// pulled in by 1e-6 relative so that the approximate distance can never decide a point the reference's (double)sqrtf comparison
// would not.
// (blo, bhi): band of t = tan(elevation) that lies strictly inside the ring r0 of the lane's first point (c_ring_band): the other
// seven points of the lane are in the same ring in ring-major input, and two compares confirm it without the polynomial.
template <int LINES>
__device__ __forceinline__ int ring_id_dev(float x, float y, float z, double min_d, double max_d, float gate_lo, float gate_hi,
                                           float blo, float bhi, int r0) {
    const float s = __fmaf_rn(x, x, __fmul_rn(y, y));
    const float r = rsqrtf(s);
    const float dist = __fmul_rn(s, r);
    const float t = __fmul_rn(z, r);
    const bool gate_ok = dist > gate_lo && dist < gate_hi && fabsf(z) < 3.0e38f;
    if (gate_ok && t > blo && t < bhi) return r0;
    int id = gate_ok ? ring_id_fast_t<LINES>(t) : -1;
    if (id < 0) {
        if (!(isfinite(x) && isfinite(y) && isfinite(z))) return 255;   // x86 int(NaN) = INT_MIN: the reference drops these
        id = ring_id_exact(x, y, z, LINES, min_d, max_d);
    }
    return id;
}

// [ring][0 / 1]: tan of the ring's elevation interval pulled in by 1e-3 degrees on either side (host: ring_band_table), for the
// sensor of the handle that launched last on this device (all handles of one sensor type share it)
__constant__ float c_ring_band[3][kMaxLines][2];

// One warp per 256-point tile (8 coalesced 512-byte rows, all loads in flight before the first use): ring id byte per point,
// and the tile summary (one ring only -> "pure"; else the set of rings present) from registers and warp votes alone.
template <int LINES, bool kLabel>
__global__ void __launch_bounds__(kClassifyThreads) k_ring_classify(ExtractParams P, float gate_lo, float gate_hi) {
    PF_PDL_ENTRY();
    const int s = blockIdx.y, lane = threadIdx.x & 31;
    const int tile = blockIdx.x * (kClassifyThreads / 32) + (threadIdx.x >> 5);
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
        P.ctrl[0] = 0;                // ticket for the extract kernel that follows in stream order
        atomicAdd(&P.ctrl[1], 1u);    // new epoch invalidates all look-back words of earlier launches
    }
    const int n = P.n[s];
    if (tile >= P.tiles || tile * kTile >= n) return;
    constexpr int R = kTile / 32;
    const int i0 = tile * kTile + lane;
    const size_t g = (size_t)s * P.stride + i0;
    float4 p[R];
#pragma unroll
    for (int k = 0; k < R; ++k)
        if (i0 + 32 * k < n) p[k] = ld_stream_f4(P.pts + g + 32 * k);
    int ring[R];
    constexpr int kCfg = LINES == 64 ? 2 : LINES == 32 ? 1 : 0;
    float blo = 2.0f, bhi = -2.0f;       // empty band: the first point always takes the full path
#pragma unroll
    for (int k = 0; k < R; ++k) {
        ring[k] = 255;
        if (i0 + 32 * k < n) {
            ring[k] = ring_id_dev<LINES>(p[k].x, p[k].y, p[k].z, P.min_d, P.max_d, gate_lo, gate_hi, blo, bhi, ring[0]);
            if (k == 0 && ring[0] < LINES) { blo = c_ring_band[kCfg][ring[0]][0]; bhi = c_ring_band[kCfg][ring[0]][1]; }
            if (kLabel) P.label[g + 32 * k] = 0;
        }
        P.ringid[g + 32 * k] = (uint8_t)ring[k];
    }
    bool same = true;
#pragma unroll
    for (int k = 1; k < R; ++k) same = same && ring[k] == ring[0];
    int all_lanes;
    __match_all_sync(kFull, ring[0], &all_lanes);
    const bool pure = __all_sync(kFull, same) && all_lanes && ring[0] < kMaxLines;
    if (pure) {      // the common case with ring-major input
        if (lane == 0) {
            int2* rt = P.ring_tiles + (size_t)s * kMaxLines + ring[0];
            atomicMin(&rt->x, tile);
            atomicMax(&rt->y, tile);
            P.tile_pure[(size_t)s * P.tiles + tile] = (uint8_t)ring[0];
        }
        return;
    }
    unsigned lo = 0u, hi = 0u;
#pragma unroll
    for (int k = 0; k < R; ++k) {
        if (ring[k] < 32) lo |= 1u << ring[k];
        else if (ring[k] < 64) hi |= 1u << (ring[k] - 32);
    }
    lo = __reduce_or_sync(kFull, lo);
    hi = __reduce_or_sync(kFull, hi);
    int2* rt = P.ring_tiles + (size_t)s * kMaxLines;
    if (lo >> lane & 1u) { atomicMin(&rt[lane].x, tile); atomicMax(&rt[lane].y, tile); }
    if (hi >> lane & 1u) { atomicMin(&rt[32 + lane].x, tile); atomicMax(&rt[32 + lane].y, tile); }
    if (lane == 0) P.tile_pure[(size_t)s * P.tiles + tile] = (uint8_t)255;
}

__global__ void k_set_int(int* p, int v) {
    PF_PDL_ENTRY();
    if (threadIdx.x == 0) *p = v;
}

// bytes of x equal to the byte b (0..255)
__device__ __forceinline__ int count_bytes_eq(unsigned x, unsigned b) {
    x ^= b * 0x01010101u;
    unsigned t = (x & 0x7f7f7f7fu) + 0x7f7f7f7fu;
    t = ~(t | x | 0x7f7f7f7fu);     // 0x80 in every byte of x that is zero
    return __popc(t);
}

// One warp per (scan, ring): ring position at which every tile of the ring's tile range starts (exclusive prefix of the
// per-tile point counts) and the ring's size, so that any sector of the ring can be located without touching the others.
__global__ void __launch_bounds__(256) k_ring_index(ExtractParams P) {
    PF_PDL_ENTRY();
    const int lane = threadIdx.x & 31;
    const int item = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (item >= P.batch * P.num_lines) return;
    const int s = item / P.num_lines, r = item - s * P.num_lines;
    int2* rtp = P.ring_tiles + (size_t)s * kMaxLines + r;
    const int2 rt = *rtp;
    __syncwarp();
    if (lane == 0) *rtp = make_int2(0x7fffffff, -1);     // hand the slot back for the next launch
    const int tlo = rt.x;
    const int ncand = rt.y >= rt.x ? rt.y - rt.x + 1 : 0;
    const uint8_t* pure = P.tile_pure + (size_t)s * P.tiles;
    const uint8_t* rid = P.ringid + (size_t)s * P.stride;
    int* off = P.tile_off + ((size_t)s * kMaxLines + r) * (P.tiles + 1);
    int carry = 0;
    int off_incl0 = 0;              // first chunk: ring position at which tile `lane` ENDS (kept in registers for the sector search)
    unsigned pure0 = 0u;            // first chunk: tiles that are pure tiles of this ring
    for (int c0 = 0; c0 < ncand; c0 += 32) {
        const int c = c0 + lane;
        const int pr = c < ncand ? pure[tlo + c] : 254;
        int cnt = pr == r ? kTile : 0;
        unsigned mixed = __ballot_sync(kFull, pr == 255);
        while (mixed) {       // tiles with several rings (or dropped points): count this ring's bytes, 8 per lane, four tiles in flight
            int b[4];
            uint2 v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                b[j] = mixed ? __ffs(mixed) - 1 : -1;
                if (mixed) mixed &= mixed - 1u;
                v[j] = make_uint2(0xffffffffu, 0xffffffffu);
                if (b[j] >= 0) v[j] = *reinterpret_cast<const uint2*>(rid + (size_t)(tlo + c0 + b[j]) * kTile + lane * 8);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (b[j] < 0) break;       // uniform
                const int m = __reduce_add_sync(kFull, count_bytes_eq(v[j].x, (unsigned)r) + count_bytes_eq(v[j].y, (unsigned)r));
                if (lane == b[j]) cnt = m;
            }
        }
        int x = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(kFull, x, o);
            if (lane >= o) x += y;
        }
        if (c < ncand) off[c] = carry + x - cnt;
        if (c0 == 0) { off_incl0 = x; pure0 = __ballot_sync(kFull, pr == r); }
        carry += __shfl_sync(kFull, x, 31);
    }
    if (lane == 0) {
        off[ncand] = carry;
        P.ring_info[(size_t)s * kMaxLines + r] = make_int4(tlo, ncand, carry, 0);
        if (carry > P.rcap) atomicOr(&P.ctrl[2], 1u);     // ring larger than the configured capacity
    }
    // sector descriptors (:81-92): sector k is located in the tile range; when every tile it touches is a pure tile of this ring,
    // its points are ONE contiguous run of the input and the extract kernel copies them without looking at any other table
    const int nr = carry;
    const bool ring_ok = nr >= 131 && nr <= P.rcap;       // :67
    const int total = nr - 10, L = total / kSectors;
    if (ncand <= 32) {
        // the whole tile range sits in the lanes' registers: six ballots instead of dependent loads
        int4 d = make_int4(0, 0, -1, 0);
#pragma unroll
        for (int k = 0; k < kSectors; ++k) {
            const int lo = L * k, hi = (k == kSectors - 1) ? total - 1 : lo + L - 1;   // hi excluded (:83-88)
            const int n_loc = hi - lo + 10, p0 = lo, p1 = lo + n_loc;
            const unsigned before = __ballot_sync(kFull, lane < ncand && off_incl0 <= p0);   // tiles that end at or before p0
            const unsigned reach = __ballot_sync(kFull, lane < ncand && off_incl0 < p1);    // tiles that end before p1 (never the last)
            const int a = __popc(before);                                               // first tile holding p0
            const int e = min(__popc(reach), ncand - 1);                                // last tile holding a point below p1
            const unsigned span = (e >= a) ? ((e - a == 31 ? kFull : ((1u << (e - a + 1)) - 1u)) << a) : 0u;
            const int off_a = __shfl_sync(kFull, off_incl0, max(a - 1, 0));
            if (lane == k && ring_ok) {
                const bool contiguous = (pure0 & span) == span;
                d = make_int4(n_loc, p0, contiguous ? (tlo + a) * kTile + (p0 - (a > 0 ? off_a : 0)) : -1, a);
            }
        }
        if (lane < kSectors) P.sec_desc[((size_t)s * kMaxLines + r) * kSectors + lane] = d;
        return;
    }
    __syncwarp();
    if (lane < kSectors) {
        const int k = lane;
        int4 d = make_int4(0, 0, -1, 0);
        if (ring_ok) {
            const int lo = L * k, hi = (k == kSectors - 1) ? total - 1 : lo + L - 1;   // hi excluded (:83-88)
            const int n_loc = hi - lo + 10, p0 = lo, p1 = lo + n_loc;
            int a = 0, b = ncand;                         // first tile with off[c + 1] > p0
            while (a < b) {
                const int m = (a + b) >> 1;
                if (off[m + 1] <= p0) a = m + 1; else b = m;
            }
            bool contiguous = true;
            for (int c = a; c < ncand && off[c] < p1; ++c) contiguous = contiguous && pure[tlo + c] == r;
            d = make_int4(n_loc, p0, contiguous ? (tlo + a) * kTile + (p0 - off[a]) : -1, a);
        }
        P.sec_desc[((size_t)s * kMaxLines + r) * kSectors + k] = d;
    }
}

// exact curvature of sector-local element j (:73-77): float sums left to right, squares and their sum in double
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

// 16-byte asynchronous copy global -> shared (L2 only): the gather keeps every row of a sector in flight without staging registers
__device__ __forceinline__ void cp_async_16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// the reference's comparison (:129-132, :138-141) for the five steps between the six points at p: float differences, double
// squares; out of line, it is taken only when an fp32 value is within 1e-7 of the threshold
__device__ __noinline__ unsigned short_steps_exact(const float4* p) {
    unsigned bits = 0u;
    for (int j = 0; j < 5; ++j) {
        const double ddx = (double)__fsub_rn(p[j + 1].x, p[j].x), ddy = (double)__fsub_rn(p[j + 1].y, p[j].y),
                     ddz = (double)__fsub_rn(p[j + 1].z, p[j].z);
        if (!(__dadd_rn(__dadd_rn(__dmul_rn(ddx, ddx), __dmul_rn(ddy, ddy)), __dmul_rn(ddz, ddz)) > 0.05)) bits |= 1u << j;
    }
    return bits;
}

// shifts that give 0 for amounts >= 32 (PTX semantics), also for "negative" amounts seen as large unsigned numbers
__device__ __forceinline__ unsigned shl_clamp(unsigned v, unsigned s) {
    unsigned r;
    asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(s));
    return r;
}
__device__ __forceinline__ unsigned shr_clamp(unsigned v, unsigned s) {
    unsigned r;
    asm("shr.u32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(s));
    return r;
}

// Greedy edge pick of one sector by one warp (:110-148).  Positions are sector-relative (rel = ring position - sector start;
// the point of rel sits at sp[rel + 5]).  A candidate word is (23-bit key << 9) | rel: the maximum word is the next element of
// the descending walk unless another live candidate shares its key bucket; then the exact double values decide (ties: higher
// index first).  NPL words per lane live in registers; every round is one redux.max, the flag update of the suppressed range
// (lane w keeps the flag bits of rel 32 w .. 32 w + 31) and a range kill.  Returns the pick count (<= 21).
// Flag bits of the picks made so far (:128-145), all at once: lane l < 20 ORs the suppressed range of pick l into the sector's
// flag words in shared memory, lane 20 the position of a 21st pick; lane w gets back the word of rel 32 w .. 32 w + 31.
__device__ __forceinline__ unsigned pick_flags(const uint16_t* srng, unsigned* sflagw, int cnt, int myedge, int lane) {
    if (lane < cnt) {
        unsigned rlo = (unsigned)myedge, m11 = 1u;
        if (lane < kEdgePerSector) {
            const unsigned rs = srng[myedge];
            rlo = rs & 511u;
            m11 = (2u << (rs >> 9)) - 1u;
        }
        const unsigned sh = rlo & 31u;
        atomicOr(&sflagw[rlo >> 5], m11 << sh);
        if (sh > 21u) atomicOr(&sflagw[(rlo >> 5) + 1], m11 >> (32u - sh));     // 11 bits at most: spills over from bit 22 on
    }
    __syncwarp();
    return lane < 16 ? sflagw[lane] : 0u;
}

template <int NPL>
__device__ __forceinline__ int sector_pick(unsigned* scand, int C, const unsigned* slink, uint16_t* srng, unsigned* sflagw, int Ls,
                                           int lane, unsigned& flagw, int& myedge) {
    unsigned word[NPL];
#pragma unroll
    for (int k = 0; k < NPL; ++k) word[k] = (k * 32 + lane < C) ? scand[k * 32 + lane] : 0u;
    __syncwarp();      // srng overlays scand from here on
    // suppressed range of every candidate, should it get picked (:128-145): forward while the steps rel -> rel+1 -> ... are
    // short (at most 5), same backward, clipped to the sector
#pragma unroll
    for (int k = 0; k < NPL; ++k) {
        if (word[k] != 0u) {
            const int rel = (int)(word[k] & 511u);
            // link bit q = step between local points q and q+1; the point of rel is local rel + 5: bits rel .. rel + 9
            const unsigned steps = __funnelshift_r(slink[rel >> 5], slink[(rel >> 5) + 1], rel & 31) & 0x3ffu;
            const int fwd = __ffs(~(steps >> 5) | 32u) - 1;
            const int back = __clz(~steps & 31u) - 27;
            const int lo = max(rel - back, 0), hi = min(rel + fwd, Ls - 1);
            srng[rel] = (uint16_t)(lo | ((hi - lo) << 9));
        }
    }
    // lanes holding two words of one bucket (rare) force the exact comparison whenever their bucket is on top
    bool dup = false;
#pragma unroll
    for (int k = 0; k < NPL; ++k)
#pragma unroll
        for (int l = k + 1; l < NPL; ++l) dup |= (word[k] != 0u) && ((word[k] ^ word[l]) >> 9) == 0u;
    const bool anydup = __any_sync(kFull, dup);
    __syncwarp();
    int cnt = 0;
#pragma unroll 1
    while (true) {
        unsigned m = word[0];
#pragma unroll
        for (int k = 1; k < NPL; ++k) m = max(m, word[k]);
        const unsigned M = __reduce_max_sync(kFull, m);
        if (M == 0u) break;
        const unsigned bucket = M >> 9;
        const unsigned top = __ballot_sync(kFull, (m >> 9) == bucket);
        const int rel = (int)(M & 511u);
        if ((top & (top - 1u)) != 0u || (anydup && __any_sync(kFull, dup && (m >> 9) == bucket))) {
            // two live candidates share the top bucket: hand the live words back and let the exact (slow) walk finish the sector
            flagw = pick_flags(srng, sflagw, cnt, myedge, lane);
#pragma unroll
            for (int k = 0; k < NPL; ++k) scand[k * 32 + lane] = word[k];
            __syncwarp();
            return -1 - cnt;
        }
        cnt++;                                                                      // :118-119
        if (lane == cnt - 1) myedge = rel;                                          // lane l keeps the l-th pick
        if (cnt > kEdgePerSector) break;                                            // :121-126 the 21st is picked, not emitted
        const unsigned rs = srng[rel];
        const unsigned rlo = rs & 511u, span = rs >> 9;
#pragma unroll
        for (int k = 0; k < NPL; ++k)
            if ((word[k] & 511u) - rlo <= span) word[k] = 0u;
    }
    flagw = pick_flags(srng, sflagw, cnt, myedge, lane);
    return cnt;
}

// more than 256 candidates in a sector: the words stay in shared memory (slow, never seen on LiDAR-like data)
__device__ __noinline__ int sector_pick_large(const float4* sp, unsigned* scand, int C, const unsigned* slink, int Ls, int lane,
                                              unsigned& flagw, int& myedge, int cnt) {
    const int wbase = lane * 32;
    while (true) {
        unsigned m = 0u;
        for (int j = lane; j < C; j += 32) m = max(m, scand[j]);
        const unsigned M = __reduce_max_sync(kFull, m);
        if (M == 0u) break;
        const unsigned bucket = M >> 9;
        // always the exact comparison inside the top bucket
        unsigned long long bk = 0ull;
        int bi = -1;
        for (int j = lane; j < C; j += 32) {
            const unsigned wd = scand[j];
            if ((wd >> 9) == bucket) {
                const int r = (int)(wd & 511u);
                const unsigned long long v = (unsigned long long)__double_as_longlong(curvature_exact(sp, r + 5));
                if (bi < 0 || v > bk || (v == bk && r > bi)) { bk = v; bi = r; }
            }
        }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            const unsigned long long ok = __shfl_xor_sync(kFull, bk, o);
            const int oi = __shfl_xor_sync(kFull, bi, o);
            if (oi >= 0 && (bi < 0 || ok > bk || (ok == bk && oi > bi))) { bk = ok; bi = oi; }
        }
        const int rel = bi;
        cnt++;
        if (cnt > kEdgePerSector) {
            if ((rel >> 5) == lane) flagw |= 1u << (rel & 31);
            break;
        }
        if (lane == cnt - 1) myedge = rel;
        const unsigned steps = __funnelshift_r(slink[rel >> 5], slink[(rel >> 5) + 1], rel & 31) & 0x3ffu;
        const int fwd = __ffs(~(steps >> 5) | 32u) - 1;
        const int back = __clz(~steps & 31u) - 27;
        const unsigned rlo = (unsigned)max(rel - back, 0), span = (unsigned)min(rel + fwd, Ls - 1) - rlo;
        const unsigned m11 = (2u << span) - 1u;
        const int sh = (int)rlo - wbase;
        flagw |= shl_clamp(m11, (unsigned)sh) | shr_clamp(m11, (unsigned)(-sh));
        for (int j = lane; j < C; j += 32)
            if ((scand[j] & 511u) - rlo <= span) scand[j] = 0u;
        __syncwarp();
    }
    return cnt;
}

// One warp per sector, handed out by ticket ((ring, sector)-major over the batch, so the sectors a warp has to look back on
// were started a whole wave earlier).  No CTA-wide barrier anywhere: gather the sector's points (+5 on either side) from the
// ring's tiles into the warp's slice of shared memory; 11-tap float curvature in the reference's order from a 20-point
// register window per lane (:73-80), squares in double; short-step link bits (:129-132, :138-141); candidates (> 0.1, :114)
// compacted by ballot; greedy pick (sector_pick); surf = unflagged (:198-205).  Output offsets: every sector publishes
// (epoch | edges | surfs), the last sector of a ring to finish publishes the ring's sum, and a sector's offset is the sum of
// the earlier rings' words plus the earlier sectors of its own ring: the compacted clouds are written once, from shared memory.
// kRefOrder: surf points leave in the reference's order -- ascending exact curvature inside the sector, the order its sorted walk at
// :198-205 produces (equal values: lower ring position first; std::sort leaves that case open) -- instead of ring position.  The
// exact doubles of all positions are kept in shared memory and every surf point counts the surf points in front of it: O(n^2 / 32)
// per lane, about five times the cost of the default order, for callers that want pcl::VoxelGrid downstream to see the very point
// order the reference's extraction would have handed it.
template <bool kLabel, bool kRefOrder>
__global__ void __launch_bounds__(kSecWarps * 32, kRefOrder ? 8 : 14) k_sector_extract(ExtractParams P) {
    PF_PDL_ENTRY();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int S = P.scap;
    unsigned char* wsm = smem_raw + (size_t)(threadIdx.x >> 5) * P.warp_smem;
    float4* sp = reinterpret_cast<float4*>(wsm);                         // [S + 24] sector points, local index q (+ window slack)
    unsigned* scand = reinterpret_cast<unsigned*>(sp + S + 24);          // [S] candidate words; once they sit in registers the
    uint16_t* srng = reinterpret_cast<uint16_t*>(scand);                 //     same bytes hold the suppressed range per position
    unsigned* slink = scand + S;                                         // [S / 32 + 2] link bits: bit q = short step q -> q+1
    unsigned* sflagw = slink + S / 32 + 2;                               // [18] flag bits of the sector (picked / suppressed)
    int* ssrc = reinterpret_cast<int*>(sflagw + 18);                     // [S] source index (label output only)
    double* sval = reinterpret_cast<double*>(wsm + P.sval_off);          // [S] exact curvature per sector position (kRefOrder only)

    const unsigned epoch = *reinterpret_cast<volatile unsigned int*>(&P.ctrl[1]);
    const int nsec = P.batch * P.num_lines * kSectors;       // < 2^23 (host): the float division below is exact
    const float inv_batch = 1.0f / (float)P.batch;
    const unsigned lt = lanemask_lt();

    // (tickets are NOT fetched ahead: a ticket held back by a busy warp starts late and everything behind it in the scan waits)
    while (true) {
        int ticket = 0;
        if (lane == 0) ticket = (int)atomicAdd(&P.ctrl[0], 1u);
        ticket = __shfl_sync(kFull, ticket, 0);
        if (ticket >= nsec) break;
        const int rk = __float2int_rd(((float)ticket + 0.5f) * inv_batch), s = ticket - rk * P.batch;
        const int r = rk / kSectors, k = rk - r * kSectors;
        const int4 desc = P.sec_desc[(size_t)(s * kMaxLines + r) * kSectors + k];
        const int n_loc = desc.x;
        const bool active = n_loc > 0;
        const int Ls = active ? n_loc - 10 : 0;
        int ne = 0, myedge = 0;
        unsigned flagw = 0u;
        if (active) {
            const float4* pts = P.pts + (size_t)s * P.stride;
            if (desc.z >= 0) {
                // A. ring-major input: the sector is one contiguous run of the scan
                const float4* src = pts + desc.z;
                for (int q = lane; q < n_loc; q += 32) {
                    cp_async_16(sp + q, src + q);
                    if (kLabel) ssrc[q] = desc.z + q;
                }
            } else {
                // B. general order: stable gather (:62) from the ring's tile range, tile by tile
                const int p0 = desc.y, p1 = p0 + n_loc;
                const int4 info = P.ring_info[(size_t)s * kMaxLines + r];
                const int tlo = info.x, ncand = info.y;
                const uint8_t* rid = P.ringid + (size_t)s * P.stride;
                const uint8_t* pure = P.tile_pure + (size_t)s * P.tiles;
                const int* off = P.tile_off + ((size_t)s * kMaxLines + r) * (P.tiles + 1);
                for (int c = desc.w; c < ncand; ++c) {
                    const int beg = off[c];
                    if (beg >= p1) break;
                    const int end = off[c + 1];
                    if (end == beg) continue;
                    const int tbase = (tlo + c) * kTile;
                    if (pure[tlo + c] == r) {
                        const int ps = max(beg, p0), pe = min(end, p1);
                        for (int pos = ps + lane; pos < pe; pos += 32) {
                            cp_async_16(sp + (pos - p0), pts + tbase + (pos - beg));
                            if (kLabel) ssrc[pos - p0] = tbase + (pos - beg);
                        }
                    } else {
                        // the ring ids of the lane's eight rows first: one load latency per tile instead of eight in a row
                        unsigned long long rb = 0ull;
#pragma unroll
                        for (int row = 0; row < kTile / 32; ++row)
                            rb |= (unsigned long long)rid[tbase + row * 32 + lane] << (8 * row);
                        int run = beg;
                        for (int row = 0; row < kTile / 32 && run < p1; ++row) {
                            const int i = tbase + row * 32 + lane;
                            const bool mt = (int)((rb >> (8 * row)) & 0xffull) == r;
                            const unsigned mm = __ballot_sync(kFull, mt);
                            const int pos = run + __popc(mm & lt);
                            if (mt && pos >= p0 && pos < p1) {
                                cp_async_16(sp + (pos - p0), pts + i);
                                if (kLabel) ssrc[pos - p0] = i;
                            }
                            run += __popc(mm);
                        }
                    }
                }
            }
            for (int i = lane; i < S / 32 + 2 + 18; i += 32) slink[i] = 0u;      // link bits and flag words
            cp_async_wait_all();
            __syncwarp();

            // C. curvature, link bits, candidates
            int C = 0;
            // A lane owns the curvature positions q0 .. q0 + 9 and the ten steps q0 - 5 .. q0 + 4 (the first ten of its window), so the
            // steps 0 .. 4 in front of the first position need no special case; one more lane than there are positions may be active
            for (int qb = 5; qb - 5 < n_loc - 1; qb += 32 * kWin) {
                const int q0 = qb + lane * kWin;
                const bool act = q0 - 5 < n_loc - 1;
                const int qa = act ? q0 : 5;
                unsigned lm = 0u;
                // two half windows of 15 points (5 positions each): keeps the register window at 45 floats
#pragma unroll 1
                for (int h = 0; h < 2; ++h) {
                    constexpr int HW = kWin / 2;
                    // packed fp32 pairs (add.f32x2 / mul.f32x2 / fma.rn.f32x2: IEEE round-to-nearest per half, so bit-identical to
                    // the scalar operations): (x, y) and (z, intensity) as they come out of the 16-byte shared load
                    float2 wxy[HW + 10], wzw[HW + 10];
#pragma unroll
                    for (int j = 0; j < HW + 10; ++j) {
                        const float4 p = sp[qa - 5 + h * HW + j];
                        wxy[j] = make_float2(p.x, p.y); wzw[j] = make_float2(p.z, p.w);
                    }
                    const float2 neg1 = make_float2(-1.0f, -1.0f);
                    const int qh = q0 + h * HW;
                    // short steps between window slots j, j+1 = local points qh - 5 + j, + 1: float differences, double squares
                    // (:129-132), decided in fp32; a value within 1e-7 of the threshold 0.05 sends the five steps to the exact test
                    unsigned sb = 0u;
                    bool unsure = false;
#pragma unroll
                    for (int j = 0; j < HW; ++j) {
                        const float2 dxy = __ffma2_rn(wxy[j], neg1, wxy[j + 1]), dzw = __ffma2_rn(wzw[j], neg1, wzw[j + 1]);   // b - a
                        const float2 sq = __fmul2_rn(dxy, dxy);
                        const float sf = __fmaf_rn(dzw.x, dzw.x, __fadd_rn(sq.x, sq.y));
                        unsure = unsure || fabsf(sf - 0.05f) <= 1e-7f;
                        if (sf < 0.05f) sb |= 1u << j;
                    }
                    if (unsure) sb = short_steps_exact(sp + (qa - 5 + h * HW));
                    // steps that end beyond the local range do not exist (bits qh - 5 + j with qh - 5 + j + 1 < n_loc)
                    const int nvalid = n_loc - 1 - (qh - 5);
                    if (nvalid < HW) sb &= nvalid > 0 ? (1u << nvalid) - 1u : 0u;
                    lm |= sb << (h * HW);
#pragma unroll
                    for (int j = 0; j < HW; ++j) {
                        auto tap = [&](const float2* a) {     // left to right, t - 10 p = t + (-10 p) (:73-75)
                            float2 u = __fadd2_rn(a[j], a[j + 1]);
                            u = __fadd2_rn(u, a[j + 2]);
                            u = __fadd2_rn(u, a[j + 3]);
                            u = __fadd2_rn(u, a[j + 4]);
                            // scalar products: ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 (one rounding)
                            u = __fadd2_rn(u, make_float2(__fmul_rn(-10.0f, a[j + 5].x), __fmul_rn(-10.0f, a[j + 5].y)));
                            u = __fadd2_rn(u, a[j + 6]);
                            u = __fadd2_rn(u, a[j + 7]);
                            u = __fadd2_rn(u, a[j + 8]);
                            u = __fadd2_rn(u, a[j + 9]);
                            u = __fadd2_rn(u, a[j + 10]);
                            return u;
                        };
                        const float2 sxy = tap(wxy), szw = tap(wzw);
                        const double dx = (double)sxy.x, dy = (double)sxy.y, dz = (double)szw.x;
                        const double val = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                        const int rel = qh + j - 5;
                        if (kRefOrder && act && rel < Ls) sval[rel] = val;
                        const bool isc = act && rel < Ls && val > 0.1;
                        const unsigned cm = __ballot_sync(kFull, isc);
                        if (isc) {
                            // key monotone in the exact value: exponent and 18 mantissa bits of the double above 2^-4 (saturating)
                            const unsigned key = min(((unsigned)__double2hiint(val) - 0x3FB00000u) >> 2, 0x7FFFFFu);
                            scand[C + __popc(cm & lt)] = (key << 9) | (unsigned)rel;
                        }
                        C += __popc(cm);
                    }
                }
                if (act && lm != 0u) {
                    const int q = q0 - 5;
                    atomicOr(&slink[q >> 5], lm << (q & 31));
                    if ((q & 31) != 0 && (lm >> (32 - (q & 31))) != 0u) atomicOr(&slink[(q >> 5) + 1], lm >> (32 - (q & 31)));
                }
            }
            __syncwarp();

            // D. greedy pick (:99-209)
            int cnt, cw = C;
            if (C <= 64) { cnt = sector_pick<2>(scand, C, slink, srng, sflagw, Ls, lane, flagw, myedge); cw = 64; }
#ifndef PF_NO_NPL4
            else if (C <= 128) { cnt = sector_pick<4>(scand, C, slink, srng, sflagw, Ls, lane, flagw, myedge); cw = 128; }
#endif
            else if (C <= 256) { cnt = sector_pick<8>(scand, C, slink, srng, sflagw, Ls, lane, flagw, myedge); cw = 256; }
            else cnt = -1;
            if (cnt < 0) cnt = sector_pick_large(sp, scand, cw, slink, Ls, lane, flagw, myedge, -1 - cnt);
            ne = min(cnt, kEdgePerSector);
        }
        // surf = unflagged (:198-205): lane w owns the bits of rel 32 w .. 32 w + 31; base = surf points before its word
        const int remw = Ls - lane * 32;
        const unsigned keep = ~flagw & (remw >= 32 ? kFull : (remw > 0 ? (1u << remw) - 1u : 0u));
        int sbase = __popc(keep);
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(kFull, sbase, o);
            if (lane >= o) sbase += y;
        }
        const int ns = __shfl_sync(kFull, sbase, 31);
        sbase -= __popc(keep);

        // E. publish this sector's counts (the word is its own message: relaxed accesses, no fence)
        unsigned long long* dsec = P.done_sec + ((size_t)s * kMaxLines + r) * kSectors;
        unsigned long long* dring = P.done_ring + (size_t)s * kMaxLines;
        if (lane == 0)
            st_relaxed_u64(dsec + k, ((unsigned long long)epoch << 32) | ((unsigned long long)ne << 20) | (unsigned long long)ns);
        // F. look back.  Own ring first: its last sector has then seen all siblings and publishes the ring's word (before it
        //    waits for anything else, so ring words never chain); then the earlier rings of the scan, one word per ring.
        //    The polls are warp-uniform loops (every lane stays until all words have arrived): a per-lane spin leaves the warp
        //    split for the rest of the iteration and every later instruction is issued once per fragment.
        int e_in = 0, su_in = 0;
        {
            unsigned long long v = 0ull;
            bool ok = lane >= k;
            unsigned nap = 100u;        // polls were 12 % of the kernel's instructions at a fixed 200 ns: back off
            while (true) {
                if (!ok) {
                    v = ld_relaxed_u64(dsec + lane);
                    ok = (unsigned)(v >> 32) == epoch;
                }
                if (__all_sync(kFull, ok)) break;
                __nanosleep(nap);
                if (nap < 1600u) nap <<= 1;
            }
            if (lane < k) {
                e_in = (int)((v >> 20) & 0xfffu);
                su_in = (int)(v & 0xfffffu);
            }
        }
        e_in = __reduce_add_sync(kFull, e_in);
        su_in = __reduce_add_sync(kFull, su_in);
        if (k == kSectors - 1 && lane == 0)
            st_relaxed_u64(dring + r, ((unsigned long long)epoch << 32) | ((unsigned long long)(e_in + ne) << 20) |
                                          (unsigned long long)(su_in + ns));
        int e = 0, su = 0;
        {
            unsigned long long v0 = 0ull, v1 = 0ull;      // r <= 63: at most two ring words per lane
            bool ok0 = lane >= r, ok1 = lane + 32 >= r;
            unsigned nap = 100u;
            while (true) {
                if (!ok0) {
                    v0 = ld_relaxed_u64(dring + lane);
                    ok0 = (unsigned)(v0 >> 32) == epoch;
                }
                if (!ok1) {
                    v1 = ld_relaxed_u64(dring + lane + 32);
                    ok1 = (unsigned)(v1 >> 32) == epoch;
                }
                if (__all_sync(kFull, ok0 && ok1)) break;
                __nanosleep(nap);
                if (nap < 1600u) nap <<= 1;
            }
            if (lane < r) { e += (int)((v0 >> 20) & 0xfffu); su += (int)(v0 & 0xfffffu); }
            if (lane + 32 < r) { e += (int)((v1 >> 20) & 0xfffu); su += (int)(v1 & 0xfffffu); }
        }
        e = __reduce_add_sync(kFull, e) + e_in;
        su = __reduce_add_sync(kFull, su) + su_in;
        if (r == P.num_lines - 1 && k == kSectors - 1 && lane == 0) { P.n_edge[s] = e + ne; P.n_surf[s] = su + ns; }

        // G. write the compacted clouds straight from shared memory
        if (active) {
            uint8_t* label = kLabel ? P.label + (size_t)s * P.stride : nullptr;
            if (lane < ne) {
                P.edge[(size_t)s * P.edge_stride + e + lane] = sp[myedge + 5];
                if (kLabel) label[ssrc[myedge + 5]] = 1;
            }
            float4* surf = P.surf + (size_t)s * P.stride + su;
            if (kRefOrder) {
                __syncwarp();
                if (lane < 16) sflagw[lane] = flagw;
                __syncwarp();
                for (int rel = lane; rel < Ls; rel += 32) {
                    if ((sflagw[rel >> 5] >> (rel & 31)) & 1u) continue;         // picked or suppressed: not a surf point
                    const double v = sval[rel];
                    int rank = 0;
                    for (int j = 0; j < Ls; ++j) {                               // same j in every lane: broadcast reads
                        const double u = sval[j];
                        const bool surf_j = ((sflagw[j >> 5] >> (j & 31)) & 1u) == 0u;
                        rank += (surf_j && (u < v || (u == v && j < rel))) ? 1 : 0;
                    }
                    st_stream_f4(surf + rank, sp[rel + 5]);
                    if (kLabel) label[ssrc[rel + 5]] = 2;
                }
            } else
            for (int row = 0; row * 32 < Ls; ++row) {
                const unsigned kw = __shfl_sync(kFull, keep, row);
                const int bw = __shfl_sync(kFull, sbase, row);
                if (kw >> lane & 1u) {
                    st_stream_f4(surf + bw + __popc(kw & lt), sp[row * 32 + lane + 5]);
                    if (kLabel) label[ssrc[row * 32 + lane + 5]] = 2;
                }
            }
        }
        __syncwarp();
    }
}

}  // namespace pf

// ------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------
struct pf_extract {
    int device = 0;
    uint64_t uid = 0;                 // unique per handle ever created in this process (a captured graph is tied to it, not to the address)
    cudaStream_t stream = nullptr;
    pf_lidar_params lidar{};
    int stride = 0, tiles = 0, max_batch = 0, rcap = 0, scap = 0, edge_stride = 0;
    int warp_smem = 0, warp_smem_label = 0;   // bytes of shared memory per warp of k_sector_extract (without / with label output)
    int ctas = 0, ctas_label = 0;             // resident CTAs of the persistent kernel on the whole device
    int ref_order = 0;                        // surf emission order: 0 ring position, 1 the reference's (ascending curvature)
    int sval_off = 0;
    int group = 1;
    // device
    float4* d_pts = nullptr;
    uint8_t* d_ringid = nullptr;
    int2* d_ring_tiles = nullptr;
    uint8_t* d_tile_pure = nullptr;
    int* d_tile_off = nullptr;
    int4* d_ring_info = nullptr;
    int4* d_sec_desc = nullptr;
    float gate_lo = 0, gate_hi = 0;
    uint8_t* d_label = nullptr;
    float4* d_edge = nullptr;
    float4* d_surf = nullptr;
    int *d_n = nullptr, *d_n_edge = nullptr, *d_n_surf = nullptr;
    unsigned long long* d_done_sec = nullptr;
    unsigned long long* d_done_ring = nullptr;
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

// ring id of an elevation angle in degrees, the reference's arithmetic (src/laserProcessingClass.cpp:30-61) in double
static int ring_of_angle(double angle, int num_lines) {
    int id;
    if (num_lines == 64) {
        if (angle >= -8.83) id = (int)((2 - angle) * 3.0 + 0.5);
        else id = 32 + (int)((-8.83 - angle) * 2.0 + 0.5);
        if (angle > 2 || angle < -24.33 || id > 63 || id < 0) return -1;
    } else if (num_lines == 32) {
        id = (int)((angle + 92.0 / 3.0) * 3.0 / 4.0);
        if (id > 31 || id < 0) return -1;
    } else {
        id = (int)((angle + 15) / 2 + 0.5);
        if (id > 15 || id < 0) return -1;
    }
    return id;
}

// For every ring the interval of tan(elevation) that maps to it, pulled in by 1e-3 degrees (fifty times the fp32 error of the
// kernel's t): all decision angles (integer bin positions of either block, the validity gates, the block split) are collected,
// the ring between two neighbours is the ring of their midpoint, adjacent intervals of one ring are merged.
static void ring_band_table(int num_lines, float band[kMaxLines][2]) {
    std::vector<double> cut;
    for (int k = -2; k <= 66; ++k) {
        if (num_lines == 64) { cut.push_back(2.0 - (k - 0.5) / 3.0); cut.push_back(-8.83 - (k - 0.5) / 2.0); }
        else if (num_lines == 32) cut.push_back(k * 4.0 / 3.0 - 92.0 / 3.0);
        else cut.push_back((k - 0.5) * 2.0 - 15.0);
    }
    if (num_lines == 64) { cut.push_back(2.0); cut.push_back(-24.33); cut.push_back(-8.83); }
    std::sort(cut.begin(), cut.end());
    double lo[kMaxLines], hi[kMaxLines];
    bool seen[kMaxLines], closed[kMaxLines];
    for (int r = 0; r < kMaxLines; ++r) { seen[r] = closed[r] = false; lo[r] = hi[r] = 0; }
    int prev = -1;
    for (size_t i = 0; i + 1 < cut.size(); ++i) {
        if (!(cut[i + 1] > cut[i])) continue;
        const int r = ring_of_angle(0.5 * (cut[i] + cut[i + 1]), num_lines);
        if (r >= 0) {
            if (!seen[r]) { seen[r] = true; lo[r] = cut[i]; hi[r] = cut[i + 1]; }
            else if (prev == r && !closed[r]) hi[r] = cut[i + 1];
            else closed[r] = true;                 // a second, separate interval of the same ring: keep the first only
        }
        if (prev >= 0 && prev != r) closed[prev] = true;
        prev = r;
    }
    const double kMargin = 1e-3, kRad = 3.14159265358979323846 / 180.0;
    for (int r = 0; r < kMaxLines; ++r) {
        band[r][0] = 2.0f; band[r][1] = -2.0f;
        if (r < num_lines && seen[r] && hi[r] - lo[r] > 4 * kMargin) {
            band[r][0] = (float)tan((lo[r] + kMargin) * kRad);
            band[r][1] = (float)tan((hi[r] - kMargin) * kRad);
        }
    }
}

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
        P.tile_off = h->d_tile_off + (size_t)s0 * kMaxLines * (h->tiles + 1);
        P.ring_info = h->d_ring_info + (size_t)s0 * kMaxLines;
        P.sec_desc = h->d_sec_desc + (size_t)s0 * kMaxLines * kSectors;
        P.done_sec = h->d_done_sec + (size_t)s0 * kMaxLines * kSectors;
        P.done_ring = h->d_done_ring + (size_t)s0 * kMaxLines;
        P.ctrl = h->d_ctrl;
        P.stride = stride; P.tiles = tiles; P.edge_stride = edge_stride; P.batch = nb;
        P.num_lines = h->lidar.num_lines; P.rcap = h->rcap; P.scap = h->scap;
        P.min_d = h->lidar.min_distance; P.max_d = h->lidar.max_distance;
        {
            const dim3 grid(div_up(tiles, kClassifyThreads / 32), nb);
            auto launch = [&](auto kern) { launch_pdl(kern, grid, dim3(kClassifyThreads), 0, h->stream, P, h->gate_lo, h->gate_hi); };
            const int nl = h->lidar.num_lines;
            if (d_label) {
                if (nl == 64) launch(k_ring_classify<64, true>);
                else if (nl == 32) launch(k_ring_classify<32, true>);
                else launch(k_ring_classify<16, true>);
            } else {
                if (nl == 64) launch(k_ring_classify<64, false>);
                else if (nl == 32) launch(k_ring_classify<32, false>);
                else launch(k_ring_classify<16, false>);
            }
        }
        launch_pdl(k_ring_index, dim3(div_up(nb * h->lidar.num_lines, 8)), dim3(256), 0, h->stream, P);
        const int want = div_up(nb * h->lidar.num_lines * kSectors, kSecWarps);
        P.sval_off = h->sval_off;
        {
            const int wsm = d_label ? h->warp_smem_label : h->warp_smem, ctas = d_label ? h->ctas_label : h->ctas;
            P.warp_smem = wsm;
            const dim3 grid(want < ctas ? want : ctas), block(kSecWarps * 32);
            const size_t smem = (size_t)kSecWarps * wsm;
            if (h->ref_order) {
                if (d_label) launch_pdl(k_sector_extract<true, true>, grid, block, smem, h->stream, P);
                else launch_pdl(k_sector_extract<false, true>, grid, block, smem, h->stream, P);
            } else {
                if (d_label) launch_pdl(k_sector_extract<true, false>, grid, block, smem, h->stream, P);
                else launch_pdl(k_sector_extract<false, false>, grid, block, smem, h->stream, P);
            }
        }
        h->launches += 3;
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
    PF_REQUIRE(cfg->max_batch <= 16384, "max_batch %d > 16384", cfg->max_batch);   // sectors per launch stay below 2^23
    PF_REQUIRE(cfg->surf_order == 0 || cfg->surf_order == 1, "surf_order %d: 0 (ring position) or 1 (the reference's: ascending curvature)", cfg->surf_order);
    int ndev = 0;
    PF_CUDA(cudaGetDeviceCount(&ndev));
    PF_REQUIRE(device >= 0 && device < ndev, "device %d not available (%d devices)", device, ndev);
    PF_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    PF_CUDA(cudaGetDeviceProperties(&prop, device));
    PF_REQUIRE(prop.major == 10, "pfilter_b200 needs an sm_100a device, found sm_%d%d", prop.major, prop.minor);

    pf_extract* h = new pf_extract();
    { static std::atomic<uint64_t> next_uid{1}; h->uid = next_uid.fetch_add(1); }
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
    h->edge_stride = 120 * lidar->num_lines;
    // shared memory of one warp of k_sector_extract: a sector of a ring of rcap points (+5 points either side, +16 of slack for
    // the register windows), candidate words, link bits, suppressed ranges (+ source indices for the label output)
    h->scap = div_up((h->rcap - 10) / kSectors + 14, 32) * 32;
    if (h->scap < 256) h->scap = 256;     // sector_pick hands up to 256 live words back through the candidate array
    h->ref_order = cfg->surf_order;
    h->warp_smem = (h->scap + 24) * 16 + h->scap * 4 + (h->scap / 32 + 2 + 18) * 4;
    h->warp_smem = div_up(h->warp_smem, 16) * 16;
    h->warp_smem_label = h->warp_smem + h->scap * 4;
    if (h->ref_order) {                        // + the exact curvature of every sector position (behind the label array's place)
        h->sval_off = h->warp_smem_label;
        h->warp_smem = h->warp_smem_label = h->sval_off + h->scap * 8;
    }
    if ((size_t)kSecWarps * h->warp_smem_label > (size_t)prop.sharedMemPerBlockOptin) {
        set_error("max_ring_points %d needs %d B shared memory (> %zu)", h->rcap, kSecWarps * h->warp_smem_label,
                  (size_t)prop.sharedMemPerBlockOptin);
        delete h;
        return PF_ERR_INVALID;
    }
    {
        auto setup = [&](auto kern, int wsm, int* ctas) -> int {
            PF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSecWarps * wsm));
            int occ = 0;
            PF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kSecWarps * 32, (size_t)kSecWarps * wsm));
            PF_REQUIRE(occ >= 1, "k_sector_extract does not fit on an SM");
            *ctas = occ * prop.multiProcessorCount;
            return PF_OK;
        };
        int rc;
        if (h->ref_order) {
            rc = setup(k_sector_extract<false, true>, h->warp_smem, &h->ctas);
            if (rc == PF_OK) rc = setup(k_sector_extract<true, true>, h->warp_smem_label, &h->ctas_label);
        } else {
            rc = setup(k_sector_extract<false, false>, h->warp_smem, &h->ctas);
            if (rc == PF_OK) rc = setup(k_sector_extract<true, false>, h->warp_smem_label, &h->ctas_label);
        }
        if (rc != PF_OK) { delete h; return rc; }
    }
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
    {
        float band[kMaxLines][2];
        ring_band_table(lidar->num_lines, band);
        const int cfg_idx = lidar->num_lines == 64 ? 2 : lidar->num_lines == 32 ? 1 : 0;
        PF_CUDA(cudaMemcpyToSymbol(c_ring_band, band, sizeof(band), sizeof(band) * cfg_idx));
    }
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
    PF_CUDA(cudaMalloc(&h->d_tile_off, sizeof(int) * (size_t)h->max_batch * kMaxLines * (h->tiles + 1)));
    PF_CUDA(cudaMalloc(&h->d_ring_info, sizeof(int4) * (size_t)h->max_batch * kMaxLines));
    PF_CUDA(cudaMalloc(&h->d_sec_desc, sizeof(int4) * (size_t)h->max_batch * kMaxLines * kSectors));
    PF_CUDA(cudaMalloc(&h->d_done_sec, sizeof(unsigned long long) * h->max_batch * kMaxLines * kSectors));
    PF_CUDA(cudaMalloc(&h->d_done_ring, sizeof(unsigned long long) * h->max_batch * kMaxLines));
    PF_CUDA(cudaMalloc(&h->d_ctrl, sizeof(unsigned) * 4));
    PF_CUDA(cudaMemset(h->d_done_sec, 0, sizeof(unsigned long long) * h->max_batch * kMaxLines * kSectors));
    PF_CUDA(cudaMemset(h->d_done_ring, 0, sizeof(unsigned long long) * h->max_batch * kMaxLines));
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
    cudaFree(h->d_surf); cudaFree(h->d_n); cudaFree(h->d_n_edge); cudaFree(h->d_n_surf); cudaFree(h->d_done_sec);
    cudaFree(h->d_done_ring); cudaFree(h->d_tile_off); cudaFree(h->d_ring_info); cudaFree(h->d_sec_desc);
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

// The two halves of pf_extract_enqueue_single: `pre` hands over what changes from scan to scan (the count, the H2D copy, the output
// slot), `kernels` enqueues the three kernels, whose launch geometry and arguments depend on the handle and the slot only -- the
// frame pipeline captures them in a CUDA graph together with the down-sampling that follows (odom.cu: front graph).
int pf_extract_enqueue_pre(pf_extract* h, const float* xyzi, int n, int device_input, const float4** src_out) {
    PF_REQUIRE(h && (xyzi || n == 0), "null argument");
    PF_REQUIRE(n >= 0 && n <= h->stride, "scan of %d points exceeds max_points %d", n, h->stride);
    PF_CUDA(cudaSetDevice(h->device));
    PF_CUDA(launch_pdl(k_set_int, dim3(1), dim3(32), 0, h->stream, h->d_n, n));
    h->launches += 1;
    const float4* src = h->d_pts;
    if (device_input) {
        src = reinterpret_cast<const float4*>(xyzi);
    } else if (n > 0) {
        PF_CUDA(cudaMemcpyAsync(h->d_pts, xyzi, (size_t)n * 16, cudaMemcpyHostToDevice, h->stream));
    }
    h->slot ^= 1;
    h->last_valid = 1;
    h->last_n = n;
    *src_out = src ? src : h->d_pts;
    return PF_OK;
}
int pf_extract_enqueue_kernels(pf_extract* h, const float4* src, int want_label) {
    return extract_launch(h, src, h->d_n, 1, h->stride, h->out_edge(), h->out_n_edge(), h->edge_stride, h->out_surf(), h->out_n_surf(),
                          want_label ? h->d_label : nullptr);
}
void pf_extract_count_launches(pf_extract* h, int n) { h->launches += (uint64_t)n; }
uint64_t pf_extract_uid(const pf_extract* h) { return h->uid; }

// Enqueue H2D + kernels for one scan; results stay on the device (used by pf_extract_run and the frame pipeline).
// device_input: xyzi is a device pointer (no copy).  The count travels as a kernel argument, so consecutive frames
// can be enqueued without a host synchronisation in between.
int pf_extract_enqueue_single(pf_extract* h, const float* xyzi, int n, int device_input, int want_label) {
    const float4* src = nullptr;
    PF_CHECK(pf_extract_enqueue_pre(h, xyzi, n, device_input, &src));
    return pf_extract_enqueue_kernels(h, src, want_label);
}

// accessors for the device-resident hand-off (odom.cu)
void pf_extract_device_outputs(pf_extract* h, const float4** edge, const int** n_edge, const float4** surf, const int** n_surf,
                               cudaStream_t* stream, int* edge_cap, int* surf_cap, int* slot, const unsigned** err_word) {
    *err_word = h->d_ctrl + 2;
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
