// C-ABI plumbing: version, thread-local error string, launch counter.
#include "common.cuh"
#include "../../include/playslot_b200.h"
#include <stdarg.h>

static thread_local char g_err[512] = "";
unsigned long long g_ps_launch_count = 0;

void ps_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* ps_last_error(void) { return g_err; }
extern "C" int ps_version(void) { return PLAYSLOT_B200_VERSION; }
extern "C" unsigned long long ps_launch_count(void) { return g_ps_launch_count; }
extern "C" void ps_reset_launch_count(void) { g_ps_launch_count = 0; }

// SMs left free by the persistent kernels (conv5x5_tc, dec_composite_tc): the decoder of one predicted frame runs on a side
// stream while the latency-bound cOCVP step of the next frame runs on the main stream; the persistent kernels otherwise
// occupy every SM (one CTA per SM with ~200 KB of shared memory) and the short kernels of the other stream could only start
// at their launch boundaries.
int g_ps_sm_reserve = 0;
extern "C" int ps_set_sm_reserve(int n_sms) {
    PS_REQUIRE(n_sms >= 0 && n_sms <= 64, "ps_set_sm_reserve: 0 <= n_sms <= 64 (got %d)", n_sms);
    g_ps_sm_reserve = n_sms;
    return PS_OK;
}
extern "C" int ps_get_sm_reserve(void) { return g_ps_sm_reserve; }

extern "C" int ps_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0;
    PS_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    PS_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    return PS_OK;
}
