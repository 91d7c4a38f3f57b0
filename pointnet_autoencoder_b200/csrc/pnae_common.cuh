// pnae_common.cuh -- shared device helpers and host-side error plumbing (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/pnae.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "pointnet_autoencoder_b200 kernels are written for sm_100a only"
#endif

// ---- host side --------------------------------------------------------------
void pnae_set_error(const char *fmt, ...);

#define PNAE_REQUIRE(cond, ...)                         \
    do {                                                \
        if (!(cond)) {                                  \
            pnae_set_error(__VA_ARGS__);                \
            return PNAE_ERR_INVALID_ARG;                \
        }                                               \
    } while (0)

#define PNAE_CUDA_OK(expr)                                                              \
    do {                                                                                \
        cudaError_t e__ = (expr);                                                       \
        if (e__ != cudaSuccess) {                                                       \
            pnae_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__),     \
                           __FILE__, __LINE__);                                         \
            return PNAE_ERR_CUDA;                                                       \
        }                                                                               \
    } while (0)

int pnae_sm_count();   // cached per device

// Launch `kernel`, optionally as a programmatic dependent launch (include/pnae.h, PNAE_OVERLAP_PREVIOUS): the grid may
// then be scheduled while the kernel before it on the stream is still draining; inside the kernel everything up to
// `griddepcontrol.wait` overlaps that tail, everything after it sees the predecessor complete.
template <typename... KArgs, typename... Args>
inline cudaError_t pnae_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool overlap, Args... args)
{
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cfg.attrs = attr; cfg.numAttrs = overlap ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#ifdef __CUDACC__
// programmatic dependent launch, device side (both are no-ops in a grid launched the ordinary way)
__device__ __forceinline__ void pnae_pdl_release() { asm volatile("griddepcontrol.launch_dependents;"); }   // the next grid may be scheduled
__device__ __forceinline__ void pnae_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }        // the previous grid is complete and visible
#endif

static inline bool pnae_aligned(const void *p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

// ---- device side ------------------------------------------------------------
#ifdef __CUDACC__

// log2(e) rounded to float: the constant nvcc uses when it expands __expf
// (seen as `FMUL R, R, 1.4426950216293334961` in the reference's sm_100a SASS).
#define PNAE_LOG2E_F 1.4426950216293334961f

// Squared distance with the rounding the reference kernels compile to
// (tf_nndistance_g.cu:25-28, tf_approxmatch_g.cu:51): the compiler contracts
// x*x+y*y+z*z into FMUL(y) -> FFMA(x) -> FFMA(z).  Written with intrinsics so
// no compiler decision is involved here.
__device__ __forceinline__ float pnae_sqdist(float dx, float dy, float dz)
{
    return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

// 2^t on the SFU.  The reference's __expf compiles (no -ftz) to ex2.approx.f32,
// which ptxas wraps in a x0.5 / square fix-up to keep denormal results; here
// results below 2^-126 flush to zero instead.  The difference is < 1.2e-38
// absolute per term and vanishes in every sum of the algorithm.
__device__ __forceinline__ float pnae_ex2(float t)
{
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t));
    return r;
}

__device__ __forceinline__ float pnae_rsqrt(float t)
{
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t));
    return r;
}

// level_j * log2(e) for the 10 levels j = 7..-2 (level = -4^j, 0 at j=-2:
// tf_approxmatch_g.cu:21-25).  level is a power of two, so
// (d*level)*log2e == d*(level*log2e) bit for bit: one multiply instead of two.
__device__ __forceinline__ float pnae_level_scale(int lev)
{
    // lev 0..8 -> j = 7-lev ; lev 9 -> 0
    return lev >= 9 ? 0.0f : -PNAE_LOG2E_F * __int_as_float((127 + 2 * (7 - lev)) << 23);
}

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

#endif  // __CUDACC__
