// Library-level entry points: version and error reporting.
#include "common.cuh"
#include "../../include/holo_b200.h"
#include <cstdarg>

static thread_local char g_err[512] = "";

void holo_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* holo_last_error(void) { return g_err; }
extern "C" int holo_version(void) { return HOLO_B200_VERSION; }

extern "C" int holo_device_info(int* sm_major, int* sm_minor, int* n_sm) {
    int dev = 0;
    HOLO_CUDA(cudaGetDevice(&dev), "holo_device_info");
    cudaDeviceProp p;
    HOLO_CUDA(cudaGetDeviceProperties(&p, dev), "holo_device_info");
    if (sm_major) *sm_major = p.major;
    if (sm_minor) *sm_minor = p.minor;
    if (n_sm) *n_sm = p.multiProcessorCount;
    return HOLO_OK;
}
