// Library-level C ABI entry points: version, error text, device count, pinned host memory.
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
bool pdl_enabled() {
    static const bool on = [] { const char* e = getenv("PF_PDL"); return !(e && e[0] == '0'); }();
    return on;
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
