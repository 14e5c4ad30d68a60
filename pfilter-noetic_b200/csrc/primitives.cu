// Stable LSD radix sort (8-bit digits) of (u32 key, u32 value) pairs with a device-resident element count,
// plus workspace management.  Three kernels per digit: per-tile digit histogram, per-digit scan across tiles,
// stable scatter (warp-private digit counters keep the input order inside a tile).
#include <cooperative_groups.h>

#include "primitives.cuh"

namespace pf {

__device__ __forceinline__ void d_sort_hist(const uint32_t* __restrict__ keys, int n, int shift, uint32_t* __restrict__ hist, int nb_cap) {
    const int ntiles = (n + kSortTile - 1) / kSortTile;
    __shared__ unsigned h[kRadix];
    for (int b = blockIdx.x; b < ntiles; b += gridDim.x) {   // persistent: the grid is bounded, tiles are not
        const int start = b * kSortTile;
        h[threadIdx.x] = 0;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kSortItems; ++k) {
            int i = start + k * kSortThreads + threadIdx.x;
            const bool ok = i < n;
            const unsigned act = __ballot_sync(0xffffffffu, ok);
            if (ok) {
                unsigned d = (keys[i] >> shift) & (kRadix - 1);
                // warp-aggregated shared atomic: one add per distinct digit in the warp
                unsigned m = __match_any_sync(act, d);
                if ((int)lane_id() == __ffs(m) - 1) atomicAdd(&h[d], (unsigned)__popc(m));
            }
        }
        __syncthreads();
        hist[(size_t)threadIdx.x * nb_cap + b] = h[threadIdx.x];
        __syncthreads();
    }
}

// one warp per digit: exclusive scan of hist[d][0..nb) in place, totals[d] = sum
__device__ __forceinline__ void d_sort_scan(uint32_t* __restrict__ hist, int n, int nb_cap, uint32_t* __restrict__ totals) {
    const int nb = (n + kSortTile - 1) / kSortTile;
    const int lane = threadIdx.x & 31;
    for (int d = blockIdx.x * 8 + (threadIdx.x >> 5); d < kRadix; d += gridDim.x * 8) {
    uint32_t* row = hist + (size_t)d * nb_cap;
    unsigned carry = 0;
    for (int base = 0; base < nb; base += 32) {
        int i = base + lane;
        unsigned v = i < nb ? row[i] : 0u, x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (i < nb) row[i] = carry + x - v;
        carry += __shfl_sync(0xffffffffu, x, 31);
    }
    if (lane == 0) totals[d] = carry;
    }
}

__device__ __forceinline__ void d_sort_scatter(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                                               uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, int n, int shift,
                                               const uint32_t* __restrict__ hist, const uint32_t* __restrict__ totals, int nb_cap, int fused_scan) {
    const int ntiles = (n + kSortTile - 1) / kSortTile;
    __shared__ unsigned wc[kSortThreads / 32][kRadix];   // warp-private digit counters -> exclusive bases
    __shared__ int tmp[9];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    for (int b = blockIdx.x; b < ntiles; b += gridDim.x) {
    const int start = b * kSortTile;
    __syncthreads();
    for (int k = 0; k < kSortThreads / 32; ++k) wc[k][tid] = 0;
    __syncthreads();
    // warp w owns the contiguous sub-chunk [start + w*256, +256), processed in 8 ordered rounds of 32 keys
    unsigned key[kSortItems], val[kSortItems], rank[kSortItems];
    const int wstart = start + w * (kSortTile / (kSortThreads / 32));
#pragma unroll
    for (int k = 0; k < kSortItems; ++k) {
        int i = wstart + k * 32 + lane;
        bool ok = i < n;
        key[k] = ok ? keys_in[i] : 0xffffffffu;
        val[k] = ok ? (vals_in ? vals_in[i] : (unsigned)i) : 0u;
        unsigned d = (key[k] >> shift) & (kRadix - 1);
        unsigned act = __ballot_sync(0xffffffffu, ok);
        rank[k] = 0;
        if (ok) {
            unsigned m = __match_any_sync(act, d);
            unsigned before = wc[w][d];
            rank[k] = before + __popc(m & lanemask_lt());
            __syncwarp(act);
            if (lane == __ffs(m) - 1) wc[w][d] = before + __popc(m);
        }
        __syncwarp();
    }
    __syncthreads();
    // digit tid: exclusive prefix over the 8 warps, plus the global base of the digit for this tile
    {
        int tot;
        unsigned t, before;
        if (fused_scan) {   // few tiles: every CTA scans the raw per-tile histogram itself (saves the k_sort_scan launch)
            t = 0; before = 0;
            const uint32_t* row = hist + (size_t)tid * nb_cap;
            for (int bb = 0; bb < ntiles; ++bb) {
                const unsigned v = row[bb];
                t += v;
                if (bb < b) before += v;
            }
        } else {
            t = totals[tid];
            before = hist[(size_t)tid * nb_cap + b];
        }
        int dbase = block_scan_excl_256((int)t, tmp, &tot);
        unsigned run = (unsigned)dbase + before;
#pragma unroll
        for (int k = 0; k < kSortThreads / 32; ++k) {
            unsigned c = wc[k][tid];
            wc[k][tid] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kSortItems; ++k) {
        int i = wstart + k * 32 + lane;
        if (i < n) {
            unsigned d = (key[k] >> shift) & (kRadix - 1);
            unsigned pos = wc[w][d] + rank[k];
            keys_out[pos] = key[k];
            vals_out[pos] = val[k];
        }
    }
    }
}

// One cooperative launch = the whole sort: per 8-bit digit { per-tile histogram | grid barrier | [digit scan | grid barrier] |
// stable scatter | grid barrier }.  All CTAs are co-resident (grid <= SMs x occupancy, cudaLaunchCooperativeKernel), so the
// barriers are cooperative-groups grid syncs and nothing returns to the host between digits.
__global__ void __launch_bounds__(kSortThreads) k_sort_coop(uint32_t* keys0, uint32_t* vals0, uint32_t* keys1, uint32_t* vals1,
                                                            const int* __restrict__ n_dev, int passes, int vals_iota, uint32_t* hist,
                                                            uint32_t* totals, int nb_cap) {
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    const int n = *n_dev;
    if (n <= kSmallSort) return;      // sorted by k_sort_small (uniform over the grid: no barrier is skipped by a subset)
    const int ntiles = (n + kSortTile - 1) / kSortTile;
    const int fused = ntiles <= 96 ? 1 : 0;
    for (int p = 0; p < passes; ++p) {
        uint32_t* kin = (p & 1) ? keys1 : keys0;
        uint32_t* vin = (p & 1) ? vals1 : vals0;
        uint32_t* kout = (p & 1) ? keys0 : keys1;
        uint32_t* vout = (p & 1) ? vals0 : vals1;
        const int shift = p * kRadixBits;
        d_sort_hist(kin, n, shift, hist, nb_cap);
        grid.sync();
        if (!fused) {
            d_sort_scan(hist, n, nb_cap, totals);
            grid.sync();
        }
        d_sort_scatter(kin, (p == 0 && vals_iota) ? nullptr : vin, kout, vout, n, shift, hist, totals, nb_cap, fused);
        if (p + 1 < passes) grid.sync();
    }
}

// Small inputs (n <= kSmallSort): the whole sort in ONE CTA of 1024 threads, passes separated by __syncthreads only.  The frame
// loop sorts a few thousand new map points every frame; four cooperative passes with two grid barriers each cost ~50 us for
// them, this kernel ~10 us.  Warp w owns the contiguous keys [256 w, 256 w + 256) and ranks them in 8 ordered rounds of 32
// (warp-private digit counters keep the sort stable).  Between the passes keys and values live in shared memory (128 KB
// ping-pong): only the first pass reads and only the last pass writes global memory.
constexpr size_t kSmallSortSmem = sizeof(unsigned) * 4 * kSmallSort;
__global__ void __launch_bounds__(1024) k_sort_small(uint32_t* keys0, uint32_t* vals0, uint32_t* keys1, uint32_t* vals1,
                                                     const int* __restrict__ n_dev, int passes, int vals_iota) {
    PF_PDL_ENTRY();
    const int n = *n_dev;
    if (n > kSmallSort || n <= 0) return;      // large inputs are sorted by k_sort_coop
    __shared__ unsigned wc[32][kRadix];
    __shared__ unsigned dbase[kRadix];
    __shared__ unsigned wtot[8];
    extern __shared__ unsigned s_pp[];          // [2][2][kSmallSort]: keys and values ping-pong between the passes in shared memory
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    for (int p = 0; p < passes; ++p) {          // pass 0 reads the global input, the last pass writes the global output
        const uint32_t* kin = p == 0 ? keys0 : s_pp + ((p - 1) & 1) * 2 * kSmallSort;
        const uint32_t* vin = p == 0 ? vals0 : s_pp + ((p - 1) & 1) * 2 * kSmallSort + kSmallSort;
        const bool last = p + 1 == passes;
        uint32_t* kout = last ? ((p & 1) ? keys0 : keys1) : s_pp + (p & 1) * 2 * kSmallSort;
        uint32_t* vout = last ? ((p & 1) ? vals0 : vals1) : s_pp + (p & 1) * 2 * kSmallSort + kSmallSort;
        const int shift = p * kRadixBits;
#pragma unroll
        for (int k = 0; k < 8; ++k) wc[(tid >> 8) + 4 * k][tid & 255] = 0;
        __syncthreads();
        unsigned key[8], val[8], rank[8];
        const int wstart = w * 256;
#pragma unroll
        for (int k = 0; k < 8; ++k) {      // all loads first: one memory latency per pass, not eight
            const int i = wstart + k * 32 + lane;
            key[k] = i < n ? kin[i] : 0xffffffffu;
            val[k] = i < n ? ((p == 0 && vals_iota) ? (unsigned)i : vin[i]) : 0u;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int i = wstart + k * 32 + lane;
            const bool ok = i < n;
            const unsigned d = (key[k] >> shift) & (kRadix - 1);
            const unsigned act = __ballot_sync(0xffffffffu, ok);
            rank[k] = 0;
            if (ok) {
                const unsigned m = __match_any_sync(act, d);
                const unsigned before = wc[w][d];
                rank[k] = before + __popc(m & lanemask_lt());
                __syncwarp(act);
                if (lane == __ffs(m) - 1) wc[w][d] = before + __popc(m);
            }
            __syncwarp();
        }
        __syncthreads();
        if (tid < kRadix) {       // digit tid: exclusive prefix over the 32 warps, then over the digits
            unsigned run = 0;
#pragma unroll 8
            for (int k = 0; k < 32; ++k) { const unsigned c = wc[k][tid]; wc[k][tid] = run; run += c; }
            unsigned x = run;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const unsigned y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
            if (lane == 31) wtot[w] = x;
            dbase[tid] = x - run;     // exclusive inside the warp of 32 digits
        }
        __syncthreads();
        if (tid < kRadix) {
            unsigned off = 0;
            for (int k = 0; k < w; ++k) off += wtot[k];
            dbase[tid] += off;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int i = wstart + k * 32 + lane;
            if (i < n) {
                const unsigned d = (key[k] >> shift) & (kRadix - 1);
                const unsigned pos = dbase[d] + wc[w][d] + rank[k];
                kout[pos] = key[k];
                vout[pos] = val[k];
            }
        }
        __syncthreads();
    }
}

// The state slots are wiped at every step; the error bits they carry (ws_error_bits) are first folded into the sticky word, so an
// overflow seen by an early stage of a frame (or by the first kind pair of a BPF frame) is still there when the frame's last kernel
// collects the errors.
__global__ void k_begin_step(unsigned int* ctrl) {
    PF_PDL_ENTRY();
    __shared__ unsigned s_err;
    if (threadIdx.x == 0) s_err = ws_error_bits(ctrl);
    __syncthreads();
    if (threadIdx.x == 0) ctrl[0] += 1;
    else if (threadIdx.x == kStickyErrWord) ctrl[kStickyErrWord] = s_err;
    else if (threadIdx.x < kCtrlWords) ctrl[threadIdx.x] = 0;
}

int workspace_create(Workspace& ws, int cap, cudaStream_t stream) {
    PF_CUDA(cudaFuncSetAttribute(k_sort_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmallSortSmem));   // per device
    ws.stream = stream;
    ws.cap = cap;
    ws.nb_cap = div_up(cap, kSortTile);
    for (int k = 0; k < 2; ++k) {
        PF_CUDA(cudaMalloc(&ws.keys[k], sizeof(uint32_t) * cap));
        PF_CUDA(cudaMalloc(&ws.vals[k], sizeof(uint32_t) * cap));
    }
    PF_CUDA(cudaMalloc(&ws.hist, sizeof(uint32_t) * kRadix * ws.nb_cap));
    PF_CUDA(cudaMalloc(&ws.totals, sizeof(uint32_t) * kRadix));
    ws.status_stride = ws.nb_cap * 8 + 8;
    if (ws.status_stride < kMinStatusStride) ws.status_stride = kMinStatusStride;
    PF_CUDA(cudaMalloc(&ws.scan_status, sizeof(unsigned long long) * 4 * ws.status_stride));
    PF_CUDA(cudaMemset(ws.scan_status, 0, sizeof(unsigned long long) * 4 * ws.status_stride));
    PF_CUDA(cudaMalloc(&ws.ctrl, sizeof(unsigned) * kCtrlWords));
    PF_CUDA(cudaMemset(ws.ctrl, 0, sizeof(unsigned) * kCtrlWords));
    return PF_OK;
}

void workspace_destroy(Workspace& ws) {
    for (int k = 0; k < 2; ++k) { cudaFree(ws.keys[k]); cudaFree(ws.vals[k]); }
    cudaFree(ws.hist); cudaFree(ws.totals); cudaFree(ws.scan_status); cudaFree(ws.ctrl);
    ws = Workspace();
}

int workspace_begin_step(Workspace& ws) {
    PF_CUDA(launch_pdl(k_begin_step, dim3(1), dim3(kCtrlWords), 0, ws.stream, ws.ctrl));
    ws.launches += 1;
    PF_CUDA(cudaGetLastError());
    return PF_OK;
}

int radix_sort(Workspace& ws, const int* n_dev, int n_cap, int passes, bool vals_iota, int* result_buf) {
    PF_REQUIRE(n_cap <= ws.cap, "radix_sort: %d items exceed workspace capacity %d", n_cap, ws.cap);
    PF_REQUIRE(passes >= 1 && passes <= 4, "radix_sort: passes must be 1..4");
    int nb = div_up(n_cap, kSortTile);
    *result_buf = passes & 1;
    if (nb == 0) return PF_OK;
    if (ws.coop_blocks == 0) {
        int per_sm = 0;
        PF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_sort_coop, kSortThreads, 0));
        PF_REQUIRE(per_sm >= 1, "radix_sort: the cooperative sort kernel does not fit an SM");
        if (per_sm > 2) per_sm = 2;
        ws.coop_blocks = per_sm * kSMs;
    }
    if (nb > ws.coop_blocks) nb = ws.coop_blocks;
    int iota = vals_iota ? 1 : 0, nb_cap = ws.nb_cap;
    PF_CUDA(launch_pdl(k_sort_small, dim3(1), dim3(1024), kSmallSortSmem, ws.stream, ws.keys[0], ws.vals[0], ws.keys[1], ws.vals[1], n_dev, passes, iota));
    ws.launches += 1;
    void* args[] = {&ws.keys[0], &ws.vals[0], &ws.keys[1], &ws.vals[1], (void*)&n_dev, &passes, &iota, &ws.hist, &ws.totals, &nb_cap};
    PF_CUDA(cudaLaunchCooperativeKernel((const void*)k_sort_coop, dim3(nb), dim3(kSortThreads), args, 0, ws.stream));
    ws.launches += 1;
    return PF_OK;
}

}  // namespace pf
