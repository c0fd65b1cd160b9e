// api.cu -- error reporting, device info and other small C-ABI entry points.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include "common.cuh"

namespace kfacb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static unsigned long long g_launches = 0;
void count_launch(int n) { __atomic_fetch_add(&g_launches, (unsigned long long)n, __ATOMIC_RELAXED); }

bool use_tc_gemm() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("KFACB_GEMM_SIMT"); v = (e && *e && atoi(e) != 0) ? 0 : 1; }
    return v == 1;
}

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        cached[dev] = v;
    }
    return cached[dev];
}

}  // namespace kfacb

#ifndef KFACB_SRC_FINGERPRINT
#define KFACB_SRC_FINGERPRINT "unknown"
#endif
extern "C" const char* kfacb_build_fingerprint(void) { return KFACB_SRC_FINGERPRINT; }

extern "C" int kfacb_version(void) { return 100; }   /* 0.1.0 */

extern "C" uint64_t kfacb_launch_count(void) { return __atomic_load_n(&kfacb::g_launches, __ATOMIC_RELAXED); }

extern "C" const char* kfacb_last_error(void) { return kfacb::g_err; }

extern "C" int kfacb_device_info(int32_t* sms, int32_t* cc_major, int32_t* cc_minor) {
    int dev = 0;
    KFACB_CHECK_CUDA(cudaGetDevice(&dev));
    int v = 0;
    KFACB_CHECK_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev));
    if (sms) *sms = v;
    KFACB_CHECK_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev));
    if (cc_major) *cc_major = v;
    KFACB_CHECK_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, dev));
    if (cc_minor) *cc_minor = v;
    return KFACB_OK;
}
