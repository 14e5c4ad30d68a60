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

// control words (device): [0] epoch (bumped once per pipeline step), [1..31] tickets of the chained scans,
// [32..] small per-step state slots (zeroed by k_begin_step), kSlotWords words each
constexpr int kCtrlWords = 256;
constexpr int kSlotBase = 32;
constexpr int kSlotWords = 32;

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
};

int workspace_create(Workspace& ws, int cap, cudaStream_t stream);
void workspace_destroy(Workspace& ws);
// zero the tickets and bump the epoch: first kernel of every pipeline step that uses chained scans
int workspace_begin_step(Workspace& ws);
// Sorts keys[0]/vals[0][0 .. *n_dev) by key, stable, `passes` 8-bit digits starting at bit 0; the result lands in
// keys[*result_buf]/vals[*result_buf] (= passes & 1).  vals_iota: treat the input values as 0,1,2,... (vals[0] need not be filled).
int radix_sort(Workspace& ws, const int* n_dev, int n_cap, int passes, bool vals_iota, int* result_buf);

#ifdef __CUDACC__
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

// Chained scan across the tiles of one kernel (decoupled look-back, single pass).
// status word: epoch[63:34] | flag[33:32] | value[31:0]; flag 1 = tile aggregate, 2 = inclusive prefix.
// Must be called by all threads of the CTA; `tile` must come from an atomic ticket so that every predecessor
// tile is already running.  Returns the exclusive prefix of this tile's aggregate.
__device__ __forceinline__ unsigned chained_scan_exclusive(unsigned long long* status, unsigned epoch, int tile, unsigned aggregate,
                                                           unsigned* smem_bcast) {
    const unsigned long long ep = ((unsigned long long)(epoch & 0x3fffffffu)) << 34;
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        unsigned excl = 0;
        if (tile == 0) {
            if (lane == 0) st_release_u64(status, ep | (2ull << 32) | aggregate);
        } else {
            if (lane == 0) st_release_u64(status + tile, ep | (1ull << 32) | aggregate);
            int look = tile - 1;
            while (true) {
                const int t = look - lane;
                unsigned long long v = 0;
                unsigned flag = 2, val = 0;
                if (t >= 0) {
                    do { v = ld_acquire_u64(status + t); } while ((v >> 34) != (ep >> 34) || ((v >> 32) & 3ull) == 0);
                    flag = (unsigned)((v >> 32) & 3ull);
                    val = (unsigned)v;
                } else {
                    val = 0;   // virtual tiles before tile 0: inclusive prefix 0
                }
                const unsigned pmask = __ballot_sync(0xffffffffu, flag == 2);
                const int first_p = __ffs(pmask) - 1;   // nearest predecessor (smallest lane) holding an inclusive prefix
                unsigned contrib = (pmask == 0 || lane <= first_p) ? val : 0;
                excl += __reduce_add_sync(0xffffffffu, contrib);
                if (pmask != 0) break;
                look -= 32;
            }
            if (lane == 0) st_release_u64(status + tile, ep | (2ull << 32) | (excl + aggregate));
        }
        if (lane == 0) *smem_bcast = excl;
    }
    __syncthreads();
    unsigned r = *smem_bcast;
    __syncthreads();
    return r;
}
#endif

}  // namespace pf
