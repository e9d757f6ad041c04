// Library-wide C-ABI plumbing: version, thread-local error string, device properties.
#include "sc_common.cuh"

namespace {
thread_local char g_err[512] = "";
}

void sc_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sc_num_sms() {
    static int cached = 0;
    if (cached) return cached;
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
        (void)cudaGetLastError();
        return SC_NUM_SMS_FALLBACK;
    }
    cached = n;
    return n;
}

int sc_max_smem_optin() {
    static int cached = 0;
    if (cached) return cached;
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess || n <= 0) {
        (void)cudaGetLastError();
        return 227 * 1024;  // sm_100 opt-in limit
    }
    cached = n;
    return n;
}

extern "C" int sc_version(void) { return 100; }
extern "C" const char* sc_last_error(void) { return g_err; }
