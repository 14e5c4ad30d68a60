// Library-level C ABI entry points: version, error text, device count, pinned host memory.
#include <atomic>
#include <stdarg.h>
#include <stdlib.h>

#include "common.cuh"

namespace pf {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
// PF_PDL=0 / 1 forces the programmatic edges off / on.  Default: on while the process holds one odometry handle, off while it holds
// several.  An edge parks the CTAs of the next kernel on the SMs while its predecessor still runs; for one sequence that hides a launch
// latency per link (+1 % end to end), with several sequences sharing the GPU the parked CTAs hold slots that another sequence's
// kernel could be running in (8 sequences on one B200: 7.5 k scans/s with the edges, 8.6 k without).
static std::atomic<int> g_live_odoms{0};
void odom_handles_changed(int delta) { g_live_odoms.fetch_add(delta); }
bool pdl_enabled() {
    static const int mode = [] { const char* e = getenv("PF_PDL"); return e ? (e[0] == '0' ? 0 : 1) : 2; }();
    return mode == 2 ? g_live_odoms.load() <= 1 : mode == 1;
}
}  // namespace pf

extern "C" int pf_version(void) { return PF_VERSION; }
extern "C" const char* pf_last_error(void) { return pf::g_err; }

extern "C" int pf_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" int pf_host_alloc(void** p, uint64_t bytes) {
    PF_REQUIRE(p, "null argument");
    PF_CUDA(cudaMallocHost(p, bytes));
    return PF_OK;
}

extern "C" int pf_host_free(void* p) {
    if (p) PF_CUDA(cudaFreeHost(p));
    return PF_OK;
}
