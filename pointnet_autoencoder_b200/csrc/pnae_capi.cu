// pnae_capi.cu -- error plumbing and device queries behind include/pnae.h.
#include <stdarg.h>
#include <string.h>

#include "pnae_common.cuh"

namespace {
thread_local char g_err[512] = "";
}

void pnae_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int pnae_sm_count()
{
    // cached per device; a race only repeats the query
    static int cache[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cache[dev] == 0) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        cache[dev] = v;
    }
    return cache[dev];
}

extern "C" int pnae_version(void) { return PNAE_VERSION; }

extern "C" const char *pnae_last_error(void) { return g_err; }

extern "C" int pnae_device_info(int *sm_count, int *cc_major, int *cc_minor)
{
    int dev = 0;
    PNAE_CUDA_OK(cudaGetDevice(&dev));
    int sms = 0, maj = 0, mnr = 0;
    PNAE_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    PNAE_CUDA_OK(cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev));
    PNAE_CUDA_OK(cudaDeviceGetAttribute(&mnr, cudaDevAttrComputeCapabilityMinor, dev));
    if (sm_count) *sm_count = sms;
    if (cc_major) *cc_major = maj;
    if (cc_minor) *cc_minor = mnr;
    if (maj != 10) {
        pnae_set_error("device compute capability %d.%d is not sm_100", maj, mnr);
        return PNAE_ERR_UNSUPPORTED;
    }
    return PNAE_OK;
}
