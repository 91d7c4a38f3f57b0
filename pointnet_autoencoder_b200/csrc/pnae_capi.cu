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

// ---------------------------------------------------------------------------
// FP32 issue-rate probe: 8 independent FFMA chains per thread, 8 CTAs of 256 threads per SM.  bench.py times it
// next to the Chamfer kernels so the roofline can also be quoted against the FFMA rate measured in the same
// run, at the same clocks.
// ---------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256) fp32_probe_kernel(float *out, float a, float b, int iters)
{
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) acc[i] = __fmaf_rn(acc[i], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) s += acc[i];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}
}  // namespace

extern "C" int pnae_fp32_probe(int iters, float *out, size_t out_floats, long long *flop, void *stream)
{
    PNAE_REQUIRE(iters >= 1 && out != nullptr && flop != nullptr, "fp32_probe: invalid argument");
    const int blocks = pnae_sm_count() * 8;
    PNAE_REQUIRE(out_floats >= (size_t)blocks * 256, "fp32_probe: out needs %zu floats", (size_t)blocks * 256);
    fp32_probe_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(out, 1.0001f, 0.5f, iters);
    PNAE_CUDA_OK(cudaGetLastError());
    *flop = (long long)blocks * 256 * 8 * 2 * iters;
    return PNAE_OK;
}
