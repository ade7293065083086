// libpof.so — ABI bookkeeping: version, error string, device query.
#include <stdarg.h>
#include <string.h>

#include "pof_common.cuh"

namespace pof {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
    set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    return (int)e;
}

int sm_count() {
    // One attribute query per device; cached per device index (no mutable
    // state that changes results, only memoisation).
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

}  // namespace pof

extern "C" {

int pof_abi_version(void) { return POF_ABI_VERSION; }

const char* pof_last_error(void) { return pof::g_err; }

int pof_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0;
    POF_CUDA(cudaGetDevice(&dev));
    int v = 0;
    if (sm_count) {
        POF_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev));
        *sm_count = v;
    }
    if (cc_major) {
        POF_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev));
        *cc_major = v;
    }
    if (cc_minor) {
        POF_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, dev));
        *cc_minor = v;
    }
    return POF_OK;
}

}  // extern "C"
