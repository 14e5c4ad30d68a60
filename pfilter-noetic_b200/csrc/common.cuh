// Shared helpers for the pfilter_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "pfilter_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "pfilter_b200 is written for sm_100a (B200) only"
#endif

namespace pf {

void set_error(const char* fmt, ...);

#define PF_CUDA(call)                                                                              \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            pf::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_));    \
            return PF_ERR_CUDA;                                                                    \
        }                                                                                          \
    } while (0)

#define PF_CHECK(expr)                                                                             \
    do {                                                                                           \
        int s_ = (expr);                                                                           \
        if (s_ != PF_OK) return s_;                                                                \
    } while (0)

#define PF_REQUIRE(cond, ...)                                                                      \
    do {                                                                                           \
        if (!(cond)) {                                                                             \
            pf::set_error(__VA_ARGS__);                                                            \
            return PF_ERR_INVALID;                                                                 \
        }                                                                                          \
    } while (0)

constexpr int kSMs = 148;   // B200

// 16-byte map / feature point viewed as raw words: {x, y, z, rgba}
struct __align__(16) Pt {
    float x, y, z;
    uint32_t rgba;   // r | g << 8 | b << 16 | a << 24 (little endian layout of pf_point)
};
static_assert(sizeof(Pt) == 16, "Pt must be 16 bytes");
static_assert(sizeof(pf_point) == 16, "pf_point must be 16 bytes");

__host__ __device__ inline uint32_t pack_rgba(uint32_t r, uint32_t g, uint32_t b, uint32_t a) {
    return r | (g << 8) | (b << 16) | (a << 24);
}
__host__ __device__ inline uint32_t pt_r(uint32_t rgba) { return rgba & 0xffu; }
__host__ __device__ inline uint32_t pt_g(uint32_t rgba) { return (rgba >> 8) & 0xffu; }

inline int div_up(int a, int b) { return (a + b - 1) / b; }

// Programmatic dependent launch (PF_PDL=0 disables): a kernel launched through launch_pdl may be scheduled as soon as every CTA of
// the kernel in front of it in the stream has passed PF_PDL_ENTRY (or exited); its own PF_PDL_ENTRY then blocks until that kernel
// has completed and its memory is visible.  A frame is a chain of ~25 small dependent kernels: the edge hides the launch latency
// of every link (also inside the captured CUDA graph, where it becomes a programmatic edge).  Every kernel launched this way MUST
// start with PF_PDL_ENTRY(); kernels without it in front of one (cooperative sort, memcpy, events) behave as ordinary predecessors.
bool pdl_enabled();
void odom_handles_changed(int delta);   // odometry handles alive in this process (pdl_enabled's default looks at it)
inline int64_t div_up64(int64_t a, int64_t b) { return (a + b - 1) / b; }

#ifdef __CUDACC__
#define PF_PDL_ENTRY()                                  \
    do {                                                \
        cudaGridDependencySynchronize();                \
        cudaTriggerProgrammaticLaunchCompletion();      \
    } while (0)

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}
// streaming 16-byte load that does not pollute L1 (data is touched once per CTA)
__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream_f4(float4* p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}
// 16-byte loads / stores with an L2 eviction-priority policy (createpolicy): a pass that will be re-read soon marks its
// lines evict_last, the pass that consumes them (and its output stream) evict_first
__device__ __forceinline__ unsigned long long l2_policy_evict_last() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ unsigned long long l2_policy_evict_normal() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ unsigned long long l2_policy_evict_unchanged() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_unchanged.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
// 0 = evict_last, 1 = evict_normal, 2 = evict_first, 3 = evict_unchanged
__device__ __forceinline__ unsigned long long l2_policy_evict_first();
__device__ __forceinline__ unsigned long long l2_policy_by_code(int c);
__device__ __forceinline__ unsigned long long l2_policy_evict_first() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ unsigned long long l2_policy_by_code(int c) {
    return c == 0 ? l2_policy_evict_last() : c == 1 ? l2_policy_evict_normal() : c == 2 ? l2_policy_evict_first() : l2_policy_evict_unchanged();
}
__device__ __forceinline__ float4 ld_f4_l2hint(const void* p, unsigned long long pol) {
    float4 v;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void st_f4_l2hint(void* p, float4 v, unsigned long long pol) {
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w),
                 "l"(pol)
                 : "memory");
}
// ---- bulk (TMA, 1-D) copies global -> shared with mbarrier completion ----------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
// one thread: bulk copy of `bytes` (multiple of 16) global -> shared, completion counted on `bar`
__device__ __forceinline__ void bulk_load(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // earlier generic reads / writes of dst are ordered before the copy
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// relaxed 64-bit accesses at gpu scope: for words that carry their whole message themselves (epoch | flag | value), so no
// ordering against other memory is needed and neither a membar nor an L1 invalidate is paid per poll
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_release_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
#endif

}  // namespace pf
