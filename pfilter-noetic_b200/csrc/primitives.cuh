// Device-wide building blocks shared by the odometry kernels: stable LSD radix sort of (u32 key, u32 value)
// pairs and a single-pass chained scan (decoupled look-back) used for order-preserving compaction.
// All element counts live in DEVICE memory (no host round trip inside a frame); grids are sized by capacity.
#pragma once
#include "common.cuh"

namespace pf {

constexpr int kSortThreads = 256;
constexpr int kSortItems = 8;
constexpr int kSortTile = kSortThreads * kSortItems;   // 2048 keys per block
constexpr int kRadixBits = 8;
constexpr int kRadix = 1 << kRadixBits;
constexpr int kMinStatusStride = (1 << 23) / 2048 + 8;   // look-back words for a scan over the 2^23 cells of a search grid (knn.cu)
constexpr int kSmallSort = 8192;   // inputs up to this size are sorted by one CTA (k_sort_small), larger ones cooperatively

// control words (device): [0] epoch (bumped once per pipeline step), [1..31] tickets of the chained scans,
// [32..] small per-step state slots (zeroed by k_begin_step), kSlotWords words each
constexpr int kCtrlWords = 256;
constexpr int kSlotBase = 32;
constexpr int kSlotWords = 32;
// Word 15 of a state slot carries the error bits of the stage that owns the slot.  Slot use inside an odometry handle: 0 / 1 VoxelGrid
// down-sampling of a kind pair (voxel.cu), 2 search-grid build (knn.cu), 3 streaming map update (merge.cu).  The last control word is
// sticky: k_begin_step folds the slots' error bits into it before it wipes them.
constexpr int kStickyErrWord = kCtrlWords - 1;
constexpr unsigned kErrMergeMask = 30u;   // map update: 2 voxel coordinate range, 4 exception capacity, 8 / 16 internal (merge.cu)
constexpr unsigned kErrGrid = 32u;        // map extent exceeds the search grid (kGridCellCap cells of 1 m)
constexpr unsigned kErrVoxel = 64u;       // VoxelGrid index space exceeds 31 bits (pcl::VoxelGrid's "leaf size too small" case)
constexpr unsigned kErrRing = 128u;       // a scan ring exceeded the extractor's max_ring_points (the ring was dropped)
constexpr unsigned kErrPose = 256u;       // the solved pose is not finite: tracking was lost and the prediction ran away

struct Workspace {
    cudaStream_t stream = nullptr;
    int cap = 0;        // max number of items in a sort / scan
    int nb_cap = 0;     // cap / kSortTile (rounded up)
    uint32_t* keys[2] = {nullptr, nullptr};
    uint32_t* vals[2] = {nullptr, nullptr};
    uint32_t* hist = nullptr;                 // [kRadix][nb_cap]
    uint32_t* totals = nullptr;               // [kRadix]
    unsigned long long* scan_status = nullptr;   // [4][nb_cap * 8]   (tiles of 256 items, up to 4 concurrent scans)
    int status_stride = 0;
    unsigned int* ctrl = nullptr;             // [kCtrlWords]
    uint64_t launches = 0;
    int coop_blocks = 0;                      // grid of the cooperative sort (SMs x resident CTAs), set on first use
    cudaEvent_t ev_a = nullptr, ev_b = nullptr;   // optional (timed stage taps): recorded around the dominant kernel of a pipeline
};

int workspace_create(Workspace& ws, int cap, cudaStream_t stream);
void workspace_destroy(Workspace& ws);
// zero the tickets and bump the epoch: first kernel of every pipeline step that uses chained scans
int workspace_begin_step(Workspace& ws);
// Sorts keys[0]/vals[0][0 .. *n_dev) by key, stable, `passes` 8-bit digits starting at bit 0; the result lands in
// keys[*result_buf]/vals[*result_buf] (= passes & 1).  vals_iota: treat the input values as 0,1,2,... (vals[0] need not be filled).
int radix_sort(Workspace& ws, const int* n_dev, int n_cap, int passes, bool vals_iota, int* result_buf);

#ifdef __CUDACC__
// error bits of a workspace as seen right now: sticky word | live slot words
__device__ __forceinline__ unsigned ws_error_bits(const unsigned* ctrl) {
    unsigned e = ctrl[kStickyErrWord];
    if (ctrl[kSlotBase + 15] | ctrl[kSlotBase + kSlotWords + 15]) e |= kErrVoxel;
    if (ctrl[kSlotBase + 2 * kSlotWords + 15]) e |= kErrGrid;
    e |= ctrl[kSlotBase + 3 * kSlotWords + 15] & kErrMergeMask;
    return e;
}
// ------------------------------------------------------------------------------------------------------------
// block-level helpers (256 threads)
// ------------------------------------------------------------------------------------------------------------
// exclusive scan of one int per thread over a 256-thread CTA; *total receives the CTA sum. tmp: 9 ints of smem.
__device__ __forceinline__ int block_scan_excl_256(int v, int* tmp, int* total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) tmp[w] = x;
    __syncthreads();
    int woff = 0, tot = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        int t = tmp[k];
        if (k < w) woff += t;
        tot += t;
    }
    __syncthreads();
    *total = tot;
    return woff + x - v;
}

// Chained scan across the tiles of one kernel (decoupled look-back, single pass), for 256-thread CTAs.
// status word: epoch[63:34] | flag[33:32] | value[31:0]; flag 1 = tile aggregate, 2 = inclusive prefix.
// The status word carries its whole message (no other memory is published through it), so relaxed accesses suffice.
// The look-back is done by the WHOLE CTA: thread j inspects tile (look - j), so one round covers 256 predecessors.  With
// hundreds of tiles in flight at once the nearest finished prefix is typically several hundred tiles back; a 32-wide window
// made every tile pay ~20 dependent L2 round trips (measured: 65% of all warp samples stalled on the barrier behind it).
// Must be called by all 256 threads of the CTA; `tile` must come from an atomic ticket so that every predecessor tile is
// already running.  `sm`: kScanSmemWords words of shared memory.  Returns the exclusive prefix of this tile's aggregate.
constexpr int kScanSmemWords = 20;
__device__ __forceinline__ unsigned chained_scan_exclusive(unsigned long long* status, unsigned epoch, int tile, unsigned aggregate,
                                                           unsigned* sm) {
    const unsigned long long ep = ((unsigned long long)(epoch & 0x3fffffffu)) << 34;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid == 0) st_relaxed_u64(status + tile, ep | ((tile == 0 ? 2ull : 1ull) << 32) | aggregate);
    unsigned excl = 0;
    int look = tile - 1;
    while (look >= 0) {       // uniform over the CTA
        const int t = look - tid;
        unsigned flag = 2, val = 0;          // virtual tiles before tile 0: inclusive prefix 0
        if (t >= 0) {
            unsigned long long v;
            do { v = ld_relaxed_u64(status + t); } while ((v >> 34) != (ep >> 34) || ((v >> 32) & 3ull) == 0);
            flag = (unsigned)((v >> 32) & 3ull);
            val = (unsigned)v;
        }
        const unsigned pmask = __ballot_sync(0xffffffffu, flag == 2);
        const int first_p = __ffs(pmask) - 1;          // nearest predecessor of this warp's window holding an inclusive prefix
        const unsigned contrib = (pmask == 0 || lane <= first_p) ? val : 0;
        const unsigned wsum = __reduce_add_sync(0xffffffffu, contrib);
        if (lane == 0) { sm[w] = wsum; sm[8 + w] = pmask != 0 ? 1u : 0u; }
        __syncthreads();
        bool found = false;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (!found) { excl += sm[k]; found = sm[8 + k] != 0; }
        }
        __syncthreads();
        if (found) break;
        look -= 256;
    }
    if (tid == 0 && tile != 0) st_relaxed_u64(status + tile, ep | (2ull << 32) | (excl + aggregate));
    return excl;
}
#endif

}  // namespace pf
