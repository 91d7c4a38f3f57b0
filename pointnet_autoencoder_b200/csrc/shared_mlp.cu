// shared_mlp.cu -- layers 1-4 of the PointNet encoder (3 -> 64 -> 64 -> 64 -> 128) for sm_100a.
//
// Reference path: get_model, models/model.py:40-56 -- four tf_util.conv2d calls with [1,3] / [1,1] kernels, i.e.
// per-point linear maps, each followed by bias, BatchNorm (training mode: statistics over all B*N points) and
// ReLU (utils/tf_util.py:155-185, 514-533).  The reference graph runs every one of these as its own op and
// writes / re-reads the (B*N, C) activation between them; round 1 of this repository did the same through
// cuBLAS + library batch-norm kernels.
//
// Training-mode BatchNorm needs a layer's statistics over ALL points before the next layer may start, so the chain
// is one kernel per layer, each of which
//   * reads the previous layer's RAW output y (fp32) once and applies that layer's folded BatchNorm + ReLU on the way
//     into shared memory (a = relu(s*y + t); s, t per channel), so normalised activations never exist in HBM;
//   * multiplies by W on the tensor cores (tcgen05 kind::tf32 with the 3xTF32 split a = hi + lo, so the product keeps
//     fp32 accuracy: the reference computes these layers in fp32) and adds the bias;
//   * writes its own raw output once and accumulates the per-channel sum and sum of squares BatchNorm needs in
//     the epilogue (one register pair per channel, one atomic per channel per CTA).
// The layers are memory-bound (K <= 64: 16.8 MB in, 16.8 / 33.5 MB out at B*N = 65536).  Round 2 first ran them as
// warp-level mma.sync m16n8k8 MMAs (22 / 35 us per layer: twelve warps per SM, latency-bound, 3 x 192 MMAs per warp and
// tile on the legacy path); as a tcgen05 pipeline with producer warps they take 9 - 10 us (mlp_layer_tc_kernel below).
// The last kernel applies layer 4's BatchNorm + ReLU and emits the bf16 K-major operand of the conv5 tcgen05 kernel
// (encoder.cu) directly.
#include <cuda_bf16.h>

#include "pnae_common.cuh"
#include "pnae_tc.cuh"

namespace {

constexpr int kKin = 64;             // input channels of layers 2-4

// 3xTF32 split by TRUNCATION: hi = the top 19 bits (what the tensor core reads of an fp32 register anyway), lo = v - hi
// (exact).  One LOP3 + one FADD per element; cvt.rna.tf32 expands to ~5 instructions on sm_100 and the 2^-21
// relative difference is below the dropped lo*lo term.
__device__ __forceinline__ unsigned tf32_hi(float v) { return __float_as_uint(v) & 0xffffe000u; }
__device__ __forceinline__ unsigned tf32_lo(float v, unsigned hi) { return __float_as_uint(v - __uint_as_float(hi)); }

// The previous layer's BatchNorm as the consumer kernels see it: batch statistics from that layer's sums (training) or
// its moving statistics (inference), folded to scale s = gamma / sqrt(var + eps) and shift t = beta - mean * s.  Every
// CTA computes the same values from the same inputs; CTA (0,0) also performs TF's moving-average update
// (moving = decay * moving + (1 - decay) * batch, biased variance: tf.contrib.layers.batch_norm, tf_util.py:529-533).
struct BnPrev {
    const float *stats;          // (2, k) sum / sum of squares of the layer's raw output, or NULL in inference mode
    const float *gamma, *beta;
    float *moving_mean, *moving_var;
    float inv_count, eps, decay;
    int training;
};

__device__ __forceinline__ void bn_fold_channel(const BnPrev &bn, int k, int c, bool update, float &s, float &t)
{
    float mean, var;
    if (bn.training) {
        mean = bn.stats[c] * bn.inv_count;
        var = fmaxf(fmaf(-mean, mean, bn.stats[k + c] * bn.inv_count), 0.f);
        if (update) {
            bn.moving_mean[c] = fmaf(bn.decay, bn.moving_mean[c], (1.f - bn.decay) * mean);
            bn.moving_var[c] = fmaf(bn.decay, bn.moving_var[c], (1.f - bn.decay) * var);
        }
    } else {
        mean = bn.moving_mean[c]; var = bn.moving_var[c];
    }
    s = bn.gamma[c] / sqrtf(var + bn.eps);
    t = fmaf(-mean, s, bn.beta[c]);
}

// Layer 1: y[p, c] = xyz[p, :] . W[:, c] + bias[c]   (K = 3: three FMAs per output, write-bound)
// 16 threads per point (4 channels each, float4 stores); a thread's channel group is fixed, so its statistics stay
// in registers for the whole grid-stride loop.
__global__ void __launch_bounds__(256)
mlp_first_kernel(long long npts, const float *__restrict__ xyz, const float *__restrict__ w, const float *__restrict__ bias,
                 float *__restrict__ out, float *__restrict__ stats)
{
    __shared__ float4 s_red[2][16][16];        // [sum | sum of squares][point slot][channel group]
    pnae_pdl_release();
    const int cg = threadIdx.x & 15;           // channels 4*cg .. 4*cg+3
    float wx[4], wy[4], wz[4], bb[4], sum[4] = {0, 0, 0, 0}, sq[4] = {0, 0, 0, 0};
#pragma unroll
    for (int j = 0; j < 4; j++) {
        wx[j] = __ldg(w + 4 * cg + j); wy[j] = __ldg(w + 64 + 4 * cg + j); wz[j] = __ldg(w + 128 + 4 * cg + j);
        bb[j] = __ldg(bias + 4 * cg + j);
    }
    pnae_pdl_wait();
    // four points per thread and pass, their coordinates requested together: one memory round trip per 64 points of a CTA
    for (long long p0 = (long long)blockIdx.x * 64 + (threadIdx.x >> 4); p0 < npts; p0 += (long long)gridDim.x * 64) {
        float x[4], y[4], z[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const long long p = min(p0 + 16 * u, npts - 1);
            x[u] = __ldg(xyz + p * 3); y[u] = __ldg(xyz + p * 3 + 1); z[u] = __ldg(xyz + p * 3 + 2);
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const long long p = p0 + 16 * u;
            if (p < npts) {
                float v[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    v[j] = fmaf(z[u], wz[j], fmaf(y[u], wy[j], fmaf(x[u], wx[j], bb[j])));
                    sum[j] += v[j]; sq[j] = fmaf(v[j], v[j], sq[j]);
                }
                *reinterpret_cast<float4 *>(out + p * 64 + 4 * cg) = make_float4(v[0], v[1], v[2], v[3]);
            }
        }
    }
    // the 16 point slots of a channel group meet in shared memory and are added in a fixed order (shared-memory float
    // atomics are compare-and-swap loops and serialise badly): one global atomic per channel and CTA
    s_red[0][threadIdx.x >> 4][cg] = make_float4(sum[0], sum[1], sum[2], sum[3]);
    s_red[1][threadIdx.x >> 4][cg] = make_float4(sq[0], sq[1], sq[2], sq[3]);
    __syncthreads();
    if (threadIdx.x < 128) {
        const int which = threadIdx.x >> 6, c = threadIdx.x & 63;
        const float *r = reinterpret_cast<const float *>(&s_red[which][0][0]) + c;
        float a = 0.f;
#pragma unroll
        for (int i = 0; i < 16; i++) a += r[i * 64];
        atomicAdd(stats + threadIdx.x, a);
    }
}

#ifdef PNAE_MLP_TRACE                  // tuning builds only (tools/mlp_trace.py): SM-clock timestamps of two CTAs' phases
__device__ long long g_mlp_trace[2][32];
extern "C" __attribute__((visibility("default"))) int pnae_debug_mlp_trace(long long *host)
{
    return (int)cudaMemcpyFromSymbol(host, g_mlp_trace, sizeof(g_mlp_trace));
}
#endif

// Sums of x, y, z and of their six distinct products over all points, in double: what layer 1's BatchNorm statistics
// follow from (mlp_layer_tc_kernel<.., true>).  One atomicAdd(double) per value and CTA; `moments` must be zero.
__global__ void __launch_bounds__(256)
xyz_moments_kernel(long long npts, const float *__restrict__ xyz, double *__restrict__ moments)
{
    __shared__ double s_part[8][9];
    pnae_pdl_release();
    pnae_pdl_wait();
    double a[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npts; p += (long long)gridDim.x * blockDim.x) {
        const double x = __ldg(xyz + p * 3), y = __ldg(xyz + p * 3 + 1), z = __ldg(xyz + p * 3 + 2);
        a[0] += x; a[1] += y; a[2] += z;
        a[3] += x * x; a[4] += x * y; a[5] += x * z; a[6] += y * y; a[7] += y * z; a[8] += z * z;
    }
#pragma unroll
    for (int i = 0; i < 9; i++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a[i] += __shfl_xor_sync(0xffffffffu, a[i], o);
        if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5][i] = a[i];
    }
    __syncthreads();
    if (threadIdx.x < 9) {
        double t = 0.0;
#pragma unroll
        for (int wi = 0; wi < 8; wi++) t += s_part[wi][threadIdx.x];
        atomicAdd(moments + threadIdx.x, t);
    }
}

// Layers 2-4 on the fifth-generation tensor cores (tcgen05, kind::tf32, 3xTF32):
//   D[channel, point] = W^T[channel, :] . a[point, :],   a = relu(s_prev * in + t_prev)
// as  Whi.ahi + Wlo.ahi + Whi.alo  (hi = the value cut to TF32's 10 mantissa bits, lo = the remainder: fp32 accuracy,
// see tf32_hi / tf32_lo), 3 x 8 UMMAs of 128 x 256 x 8 per 256-point tile.  Channels are the M (TMEM lane) dimension:
// an epilogue thread owns one channel, so the statistics are private register sums and a warp's store of one point is
// 32 consecutive channels = one 128-byte segment of the row-major output.
//   warp 0      MMA issuer (+ TMEM allocation)
//   warps 1-8   producers: raw rows from global -> BatchNorm + ReLU -> hi / lo split -> the two 128B-swizzled K-major
//               operand tiles in shared memory (the layout a TMA load with SWIZZLE_128B would have produced); the
//               global loads of the NEXT tile are in registers while the tensor cores work on this one
//   warps 9-12  epilogue: TMEM -> registers -> bias -> global, sum and sum of squares per channel
// One operand stage (2 x 64 KB for 256 points) next to the resident weights (2 x 32 KB); accumulators double-buffered in
// TMEM, so the epilogue's stores overlap the next tile's MMAs.  Weights narrower than 128 channels are zero-padded.
constexpr int kTcPoints = 256;                 // points per tile (UMMA N)
constexpr int kTcM = 128;                      // channel rows of the weight operand (UMMA M)
#ifndef PNAE_TC_EPI_WARPS
#define PNAE_TC_EPI_WARPS 4
#endif
constexpr int kTcProducerWarps = 8, kTcEpiWarps = PNAE_TC_EPI_WARPS;   // epilogue: one or two warps per TMEM lane quadrant
constexpr int kTcThreads = 32 * (1 + kTcProducerWarps + kTcEpiWarps);
constexpr uint32_t kTcWBytes = 2 * kTcM * 128;         // one weight operand (two 32-element k boxes)
constexpr uint32_t kTcABytes = 2 * kTcPoints * 128;    // one activation operand
constexpr size_t kTcSmem = 1024 + 2 * kTcWBytes + 2 * kTcABytes + 256 + 6 * kKin * sizeof(float);

// byte offset of element (row, k) in a K-major SWIZZLE_128B operand of `rows` rows and 64 fp32 columns: two boxes of
// 32 columns; inside a box a row is 128 bytes and its 16-byte chunks are XOR-ed with the row's position in its group of 8
__device__ __forceinline__ uint32_t tc_offset(int rows, int row, int k)
{
    return (uint32_t)((k >> 5) * rows * 128 + row * 128 + ((((k & 31) >> 2) ^ (row & 7)) << 4) + (k & 3) * 4);
}

// XYZ: the input rows are not read but computed -- layer 1 (3 -> 64) folded into this kernel's producers:
//   y1 = xyz . W1 + b1  (three FMAs per channel, the arithmetic of mlp_first_kernel),  a = relu(s1 * y1 + t1),
// with layer 1's BatchNorm statistics derived from the moments of xyz (xyz_moments_kernel): y1 is affine in xyz, so
//   mean(y1_c) = w_c . mean(x) + b_c,   E[y1_c^2] = w_c^T E[x x^T] w_c + 2 b_c w_c . mean(x) + b_c^2.
// The (B*N, 64) tensor of layer 1 is then neither written nor read.
struct FirstLayer {
    const float *xyz, *w1, *b1;          // (npts,3), (3,64), (64)
    const double *moments;               // sum x, y, z, xx, xy, xz, yy, yz, zz over all points
};

template <int KOUT, bool XYZ>
__global__ void __launch_bounds__(kTcThreads, 1)
mlp_layer_tc_kernel(long long npts, const float *__restrict__ in, const FirstLayer first, const BnPrev bn, const float *__restrict__ w,
                    const float *__restrict__ bias, float *__restrict__ out, float *__restrict__ stats)
{
#ifdef PNAE_MLP_TRACE
    const int trace_cta = blockIdx.x == 0 ? 0 : blockIdx.x == 100 ? 1 : -1;
#define TC_TRACE(i) do { if (trace_cta >= 0 && (i) < 32) g_mlp_trace[trace_cta][i] = clock64(); } while (0)
#else
#define TC_TRACE(i) do { } while (0)
#endif
    if (threadIdx.x == 0) TC_TRACE(0);
    extern __shared__ __align__(1024) unsigned char tc_smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(tc_smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char *Whi = smem, *Wlo = Whi + kTcWBytes, *Ahi = Wlo + kTcWBytes, *Alo = Ahi + kTcABytes;
    uint64_t *bars = reinterpret_cast<uint64_t *>(Alo + kTcABytes);
    uint64_t *hi_full = bars, *lo_full = bars + 1, *hi_empty = bars + 2, *lo_empty = bars + 3, *t_full = bars + 4, *t_empty = bars + 6;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 8);
    float *sp = reinterpret_cast<float *>(bars) + 64, *tp = sp + kKin;        // folded BatchNorm of the previous layer
    float *w1s = tp + kKin;                                                   // (XYZ) layer 1's weights and bias: [4][64]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long ntiles = (npts + kTcPoints - 1) / kTcPoints;
    constexpr int kout = KOUT;

    pnae_pdl_release();
    if (threadIdx.x == 0) {
        mbar_init(hi_full, kTcProducerWarps); mbar_init(lo_full, kTcProducerWarps); mbar_init(hi_empty, 1); mbar_init(lo_empty, 1);
        for (int i = 0; i < 2; i++) { mbar_init(t_full + i, 1); mbar_init(t_empty + i, kTcEpiWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // W^T, hi and lo, zero rows past kout.  An item is one 16-byte chunk = 4 consecutive k of one channel; consecutive
    // threads take consecutive channels, so the four global reads are coalesced and the two 16-byte shared stores of a
    // quarter warp fall into eight different bank groups (the swizzle XORs the chunk with the row).
    for (int i = threadIdx.x; i < (kKin / 4) * kTcM; i += kTcThreads) {
        const int k4 = i / kTcM, c = i - k4 * kTcM;
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; e++) v[e] = c < kout ? __ldg(w + (size_t)(4 * k4 + e) * kout + c) : 0.f;
        uint4 hi, lo;
        hi.x = tf32_hi(v[0]); hi.y = tf32_hi(v[1]); hi.z = tf32_hi(v[2]); hi.w = tf32_hi(v[3]);
        lo.x = tf32_lo(v[0], hi.x); lo.y = tf32_lo(v[1], hi.y); lo.z = tf32_lo(v[2], hi.z); lo.w = tf32_lo(v[3], hi.w);
        const uint32_t off = tc_offset(kTcM, c, 4 * k4);
        *reinterpret_cast<uint4 *>(Whi + off) = hi;
        *reinterpret_cast<uint4 *>(Wlo + off) = lo;
    }
    fence_proxy_async_smem();
    if (threadIdx.x == 0) TC_TRACE(1);
    // everything above read only this layer's own parameters and may have run under the previous kernel's tail
    // (PNAE_OVERLAP_PREVIOUS); the input and its statistics belong to the time after it
    if (XYZ && threadIdx.x < kKin) {
        const int c = threadIdx.x;
        w1s[c] = __ldg(first.w1 + c); w1s[64 + c] = __ldg(first.w1 + 64 + c); w1s[128 + c] = __ldg(first.w1 + 128 + c);
        w1s[192 + c] = __ldg(first.b1 + c);
    }
    pnae_pdl_wait();
    if (threadIdx.x < kKin) {
        if (XYZ) {
            // layer 1's BatchNorm from the moments of xyz (double: the variance is a difference of second moments)
            const int c = threadIdx.x;
            float s, t;
            if (bn.training) {
                const double inv = 1.0 / (double)npts;
                const double wx = w1s[c], wy = w1s[64 + c], wz = w1s[128 + c], b = w1s[192 + c];
                const double *m = first.moments;
                const double mx = m[0] * inv, my = m[1] * inv, mz = m[2] * inv;
                const double lin = wx * mx + wy * my + wz * mz;
                const double quad = wx * wx * m[3] + wy * wy * m[6] + wz * wz * m[8] + 2.0 * (wx * wy * m[4] + wx * wz * m[5] + wy * wz * m[7]);
                const double mean = lin + b;
                const double var = fmax(quad * inv + 2.0 * b * lin + b * b - mean * mean, 0.0);
                if (blockIdx.x == 0) {
                    bn.moving_mean[c] = fmaf(bn.decay, bn.moving_mean[c], (1.f - bn.decay) * (float)mean);
                    bn.moving_var[c] = fmaf(bn.decay, bn.moving_var[c], (1.f - bn.decay) * (float)var);
                }
                s = bn.gamma[c] / sqrtf((float)var + bn.eps);
                t = fmaf(-(float)mean, s, bn.beta[c]);
            } else {
                s = bn.gamma[c] / sqrtf(bn.moving_var[c] + bn.eps);
                t = fmaf(-bn.moving_mean[c], s, bn.beta[c]);
            }
            sp[c] = s; tp[c] = t;
        } else {
            bn_fold_channel(bn, kKin, threadIdx.x, blockIdx.x == 0, sp[threadIdx.x], tp[threadIdx.x]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) TC_TRACE(2);

    if (warp == 0) {
        // ===== MMA issuer =====
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_tf32(kTcM, kTcPoints);
            int it = 0;
            for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++) {
                const int buf = it & 1;
                if (it >= 2) mbar_wait(t_empty + buf, ((it >> 1) - 1) & 1);      // the epilogue drained this accumulator
                // The two products that read a_hi first, a_lo's last: the producers write the next tile's a_hi while the
                // third product runs and its a_lo while the next tile's first two do, so the tensor pipe never waits for
                // a whole tile to be staged.
                const unsigned char *wa[3] = {Wlo, Whi, Whi};
                const unsigned char *ab[3] = {Ahi, Ahi, Alo};
#pragma unroll
                for (int pr = 0; pr < 3; pr++) {
                    if (pr == 0) mbar_wait(hi_full, it & 1);
                    if (pr == 2) mbar_wait(lo_full, it & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                    for (int ks = 0; ks < kKin / 8; ks++) {
                        const int kb = ks >> 2, kin = ks & 3;                      // 4 UMMA_K=8 steps per 128-byte row
                        const uint64_t ad = umma_desc_sw128(smem_u32(wa[pr] + (size_t)kb * kTcM * 128) + kin * 32);
                        const uint64_t bd = umma_desc_sw128(smem_u32(ab[pr] + (size_t)kb * kTcPoints * 128) + kin * 32);
                        umma_tf32(tmem_base + buf * kTcPoints, ad, bd, idesc, (pr | ks) != 0);
                    }
                    if (pr == 1) umma_commit(hi_empty);     // a_hi may be rewritten once the first two products retire
                }
                umma_commit(lo_empty);          // ... and a_lo once the third does
                umma_commit(t_full + buf);      // accumulator ready for the epilogue
                TC_TRACE(5 + 4 * it);
            }
        }
    } else if (warp <= kTcProducerWarps) {
        // ===== producers: a thread owns one 16-byte column chunk (4 input channels) of rows r0, r0 + 16, ... =====
        const int pt = threadIdx.x - 32;
        const int c4 = pt & 15, r0 = pt >> 4;
        float s4[4], t4[4], wx1[4], wy1[4], wz1[4], bb1[4];
#pragma unroll
        for (int e = 0; e < 4; e++) {
            s4[e] = sp[4 * c4 + e]; t4[e] = tp[4 * c4 + e];
            if (XYZ) { wx1[e] = w1s[4 * c4 + e]; wy1[e] = w1s[64 + 4 * c4 + e]; wz1[e] = w1s[128 + 4 * c4 + e]; bb1[e] = w1s[192 + 4 * c4 + e]; }
            else { wx1[e] = wy1[e] = wz1[e] = bb1[e] = 0.f; }
        }
        const uint32_t kbase = (uint32_t)((c4 >> 3) * kTcPoints * 128);
        int it = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++) {
            const long long p0 = tile * kTcPoints;
            // the activation of row j, chunk c4 from what was loaded: BatchNorm + ReLU of the raw input, or (XYZ) of layer 1
            // computed on the spot
            float4 v[XYZ ? 1 : kTcPoints / 16];
            float px[XYZ ? kTcPoints / 16 : 1], py[XYZ ? kTcPoints / 16 : 1], pz[XYZ ? kTcPoints / 16 : 1];
#pragma unroll
            for (int j = 0; j < kTcPoints / 16; j++) {
                const long long row = min(p0 + r0 + 16 * j, npts - 1);         // rows past the end repeat the last row; the epilogue ignores them
                if (XYZ) {
                    px[j] = __ldg(first.xyz + row * 3); py[j] = __ldg(first.xyz + row * 3 + 1); pz[j] = __ldg(first.xyz + row * 3 + 2);
                } else {
                    v[j] = __ldg(reinterpret_cast<const float4 *>(in + row * kKin) + c4);
                }
            }
            auto act = [&](int j) -> float4 {
                float4 a;
                if (XYZ) {
                    a.x = fmaf(pz[j], wz1[0], fmaf(py[j], wy1[0], fmaf(px[j], wx1[0], bb1[0])));
                    a.y = fmaf(pz[j], wz1[1], fmaf(py[j], wy1[1], fmaf(px[j], wx1[1], bb1[1])));
                    a.z = fmaf(pz[j], wz1[2], fmaf(py[j], wy1[2], fmaf(px[j], wx1[2], bb1[2])));
                    a.w = fmaf(pz[j], wz1[3], fmaf(py[j], wy1[3], fmaf(px[j], wx1[3], bb1[3])));
                } else {
                    a = v[j];
                }
                a.x = fmaxf(fmaf(a.x, s4[0], t4[0]), 0.f); a.y = fmaxf(fmaf(a.y, s4[1], t4[1]), 0.f);
                a.z = fmaxf(fmaf(a.z, s4[2], t4[2]), 0.f); a.w = fmaxf(fmaf(a.w, s4[3], t4[3]), 0.f);
                return a;
            };
            // the two operand tiles one after the other, each behind its own barriers (the activation is cheap enough to
            // form twice: keeping it would cost 64 registers)
            if (it >= 1) mbar_wait(hi_empty, (it - 1) & 1);
            if (pt == 0) TC_TRACE(3 + 4 * it);
#pragma unroll
            for (int j = 0; j < kTcPoints / 16; j++) {
                const int r = r0 + 16 * j;
                const float4 a = act(j);
                uint4 hi;
                hi.x = tf32_hi(a.x); hi.y = tf32_hi(a.y); hi.z = tf32_hi(a.z); hi.w = tf32_hi(a.w);
                *reinterpret_cast<uint4 *>(Ahi + kbase + (uint32_t)(r * 128 + (((c4 & 7) ^ (r & 7)) << 4))) = hi;
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(hi_full);
            if (it >= 1) mbar_wait(lo_empty, (it - 1) & 1);
#pragma unroll
            for (int j = 0; j < kTcPoints / 16; j++) {
                const int r = r0 + 16 * j;
                const float4 a = act(j);
                uint4 lo;
                lo.x = tf32_lo(a.x, tf32_hi(a.x)); lo.y = tf32_lo(a.y, tf32_hi(a.y));
                lo.z = tf32_lo(a.z, tf32_hi(a.z)); lo.w = tf32_lo(a.w, tf32_hi(a.w));
                *reinterpret_cast<uint4 *>(Alo + kbase + (uint32_t)(r * 128 + (((c4 & 7) ^ (r & 7)) << 4))) = lo;
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(lo_full);
            if (pt == 0) TC_TRACE(4 + 4 * it);
        }
    } else {
        // ===== epilogue: one thread per channel =====
        const int q = warp & 3;                                    // TMEM lane quadrant this warp may read
        const int ch = q * 32 + lane;
        const bool live = ch < kout;                               // (whole warps: kout is a multiple of 32)
        const float bs = live ? __ldg(bias + ch) : 0.f;
        float sum = 0.f, sq = 0.f;
        int it = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++) {
            const int buf = it & 1;
            const long long p0 = tile * kTcPoints;
            const int valid = (int)min((long long)kTcPoints, npts - p0);
            mbar_wait(t_full + buf, (it >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (q * 32 < kout) {                                   // (warp-uniform) this quadrant holds real channels
                constexpr int kPartsPerWarp = (kTcPoints / 64) / (kTcEpiWarps / 4);
                const int part0 = ((warp - 1 - kTcProducerWarps) >> 2) * kPartsPerWarp;      // with two warps per quadrant: its half of the columns
#pragma unroll 1
                for (int part = part0; part < part0 + kPartsPerWarp; part++) {        // 64 columns at a time: two TMEM loads in flight
                    if (part * 64 >= valid) break;
                    uint32_t r[2][32];
                    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * kTcPoints + part * 64;
                    tmem_ld32_issue(taddr, r[0]);
                    tmem_ld32_issue(taddr + 32, r[1]);
                    tmem_ld_wait();
                    float *op = out + (p0 + part * 64) * KOUT + ch;           // column j of this part: op[j * KOUT], a constant offset
                    if (part * 64 + 64 <= valid) {
#pragma unroll
                        for (int j = 0; j < 64; j++) {
                            const float y = __uint_as_float(r[j >> 5][j & 31]) + bs;
                            op[j * KOUT] = y;
                            sum += y; sq = fmaf(y, y, sq);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 64; j++)
                            if (part * 64 + j < valid) {
                                const float y = __uint_as_float(r[j >> 5][j & 31]) + bs;
                                op[j * KOUT] = y;
                                sum += y; sq = fmaf(y, y, sq);
                            }
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(t_empty + buf);
            if (threadIdx.x == 32 * (1 + kTcProducerWarps)) TC_TRACE(6 + 4 * it);
        }
        if (live) { atomicAdd(stats + ch, sum); atomicAdd(stats + kout + ch, sq); }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
    if (threadIdx.x == 0) TC_TRACE(31);
}

// relu(s * y + t) -> bf16, (npts, k) row-major = the K-major operand tile layout the conv5 kernel's TMA map expects
__global__ void __launch_bounds__(256)
mlp_apply_bf16_kernel(long long n4, int k, const float *__restrict__ in, const BnPrev bn, __nv_bfloat16 *__restrict__ out)
{
    __shared__ float s[256], t[256];
    pnae_pdl_release();
    pnae_pdl_wait();
    for (int c = threadIdx.x; c < k; c += blockDim.x) bn_fold_channel(bn, k, c, blockIdx.x == 0, s[c], t[c]);
    __syncthreads();
    const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
    auto emit = [&](long long i, const float4 &v, const float (&sc)[4], const float (&sh)[4]) {
        const float a0 = fmaxf(fmaf(v.x, sc[0], sh[0]), 0.f), a1 = fmaxf(fmaf(v.y, sc[1], sh[1]), 0.f);
        const float a2 = fmaxf(fmaf(v.z, sc[2], sh[2]), 0.f), a3 = fmaxf(fmaf(v.w, sc[3], sh[3]), 0.f);
        __nv_bfloat162 lo = __floats2bfloat162_rn(a0, a1), hi = __floats2bfloat162_rn(a2, a3);
        uint2 pk;
        pk.x = *reinterpret_cast<unsigned *>(&lo); pk.y = *reinterpret_cast<unsigned *>(&hi);
        reinterpret_cast<uint2 *>(out)[i] = pk;
    };
    if ((stride * 4) % k == 0) {
        // the usual case (k divides 1024): a thread meets the same four channels on every trip, so their scale / shift
        // live in registers and four loads are in flight per trip
        const int c = (int)((i0 * 4) % k);
        const float sc[4] = {s[c], s[c + 1], s[c + 2], s[c + 3]}, sh[4] = {t[c], t[c + 1], t[c + 2], t[c + 3]};
        long long i = i0;
        for (; i + 3 * stride < n4; i += 4 * stride) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; u++) v[u] = __ldg(reinterpret_cast<const float4 *>(in) + i + u * stride);
#pragma unroll
            for (int u = 0; u < 4; u++) emit(i + u * stride, v[u], sc, sh);
        }
        for (; i < n4; i += stride) emit(i, __ldg(reinterpret_cast<const float4 *>(in) + i), sc, sh);
    } else {
        for (long long i = i0; i < n4; i += stride) {
            const int c = (int)((i * 4) % k);
            const float sc[4] = {s[c], s[c + 1], s[c + 2], s[c + 3]}, sh[4] = {t[c], t[c + 1], t[c + 2], t[c + 3]};
            emit(i, __ldg(reinterpret_cast<const float4 *>(in) + i), sc, sh);
        }
    }
}

// BatchNorm bookkeeping of one layer in one launch: batch statistics from the layer kernel's sums (training) or the
// moving statistics (inference) -> folded scale s = gamma / sqrt(var + eps) and shift t = beta - mean * s for the NEXT
// kernel's prologue, plus TF's moving-average update (moving = decay * moving + (1 - decay) * batch; biased variance:
// tf.contrib.layers.batch_norm, utils/tf_util.py:529-533).
__global__ void bn_fold_kernel(int k, const float *__restrict__ stats, float inv_count, const float *__restrict__ gamma,
                               const float *__restrict__ beta, float eps, float decay, int training,
                               float *__restrict__ moving_mean, float *__restrict__ moving_var,
                               float *__restrict__ s_out, float *__restrict__ t_out)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= k) return;
    float mean, var;
    if (training) {
        mean = stats[c] * inv_count;
        var = fmaxf(fmaf(-mean, mean, stats[k + c] * inv_count), 0.f);
        moving_mean[c] = fmaf(decay, moving_mean[c], (1.f - decay) * mean);
        moving_var[c] = fmaf(decay, moving_var[c], (1.f - decay) * var);
    } else {
        mean = moving_mean[c]; var = moving_var[c];
    }
    const float sc = gamma[c] / sqrtf(var + eps);
    s_out[c] = sc;
    t_out[c] = fmaf(-mean, sc, beta[c]);
}

// conv5's BatchNorm + ReLU + max-pool finish on (B, C): from the tcgen05 kernel's per-(element, channel) max / min / sum /
// sum of squares of y0 = x @ w (no bias) to the pooled feature, in one launch.
//   pooled = relu((ext0 - mean0) * s + beta),  ext0 = max where gamma >= 0 else min,  s = gamma / sqrt(var + eps)
// Also leaves what the backward needs: inv (C), mean0 (C), ext0 (B,C), z (B,C).
// This is the last kernel of the encoder chain, so its latency is all exposed: a CTA owns 32 channels and spreads the
// batch over 8 thread rows, every load a thread needs (sums for the statistics, extrema for the output) is issued before
// the first use, and the 8 partial sums of a channel meet in shared memory in a fixed order -- two dependent memory
// round trips in all (one thread per channel walking the batch took 13 us).
constexpr int kFinRows = 8, kFinPer = 8;      // thread rows per CTA, batch elements per thread held in registers at once
__global__ void __launch_bounds__(32 * kFinRows)
conv5_finish_kernel(int b, int c, const float *__restrict__ vmax, const float *__restrict__ vmin,
                    const float *__restrict__ vsum, const float *__restrict__ vsq, const float *__restrict__ bias,
                    const float *__restrict__ gamma, const float *__restrict__ beta, float *__restrict__ moving_mean,
                    float *__restrict__ moving_var, float inv_count, float eps, float decay, int training,
                    float *__restrict__ pooled, float *__restrict__ inv_out, float *__restrict__ mean0_out,
                    float *__restrict__ ext0_out, float *__restrict__ z_out)
{
    __shared__ float s_sum[kFinRows][32], s_sq[kFinRows][32], s_mean[32], s_scale[32];
    pnae_pdl_release();
    pnae_pdl_wait();
    const int cx = threadIdx.x & 31, row = threadIdx.x >> 5;
    const int ch = blockIdx.x * 32 + cx;
    const bool live = ch < c;
    const float g = live ? gamma[ch] : 0.f, be = live ? beta[ch] : 0.f;
    float mean0 = 0.f, var = 1.f;
    if (!training && live) { mean0 = moving_mean[ch] - bias[ch]; var = moving_var[ch]; }
    for (int i0 = 0; i0 < b; i0 += kFinRows * kFinPer) {                      // (one pass for b <= 64)
        float e0[kFinPer], a1[kFinPer], a2[kFinPer];
#pragma unroll
        for (int j = 0; j < kFinPer; j++) {
            const int i = i0 + row + kFinRows * j;
            const bool ok = live && i < b;
            const size_t o = (size_t)i * c + ch;
            e0[j] = ok ? (g >= 0.f ? vmax[o] : vmin[o]) : 0.f;
            a1[j] = (ok && training) ? vsum[o] : 0.f;
            a2[j] = (ok && training) ? vsq[o] : 0.f;
        }
        if (training) {
            // statistics over the WHOLE batch are needed before any output: with b > 64 this first pass only gathers
            // them (the loop below re-reads the sums of the other passes)
            if (i0 == 0) {
                float p1 = 0.f, p2 = 0.f;
                for (int ib = 0; ib < b; ib += kFinRows * kFinPer) {
#pragma unroll
                    for (int j = 0; j < kFinPer; j++) {
                        const int i = ib + row + kFinRows * j;
                        if (ib == 0) { p1 += a1[j]; p2 += a2[j]; }
                        else if (live && i < b) { p1 += vsum[(size_t)i * c + ch]; p2 += vsq[(size_t)i * c + ch]; }
                    }
                }
                s_sum[row][cx] = p1; s_sq[row][cx] = p2;
                __syncthreads();
                if (row == 0) {
                    float s1 = 0.f, s2 = 0.f;
#pragma unroll
                    for (int r = 0; r < kFinRows; r++) { s1 += s_sum[r][cx]; s2 += s_sq[r][cx]; }
                    const float m = s1 * inv_count, v = fmaxf(fmaf(-m, m, s2 * inv_count), 0.f);
                    s_mean[cx] = m; s_scale[cx] = v;
                    if (live) {
                        moving_mean[ch] = fmaf(decay, moving_mean[ch], (1.f - decay) * (m + bias[ch]));
                        moving_var[ch] = fmaf(decay, moving_var[ch], (1.f - decay) * v);
                    }
                }
                __syncthreads();
                mean0 = s_mean[cx]; var = s_scale[cx];
            }
        }
        const float inv = 1.0f / sqrtf(var + eps), s = g * inv;
        if (i0 == 0 && row == 0 && live) { inv_out[ch] = inv; mean0_out[ch] = mean0; }
#pragma unroll
        for (int j = 0; j < kFinPer; j++) {
            const int i = i0 + row + kFinRows * j;
            if (live && i < b) {
                const size_t o = (size_t)i * c + ch;
                const float z = fmaf(e0[j] - mean0, s, be);
                ext0_out[o] = e0[j]; z_out[o] = z; pooled[o] = fmaxf(z, 0.f);
            }
        }
    }
}


}  // namespace

extern "C" int pnae_mlp_first(long long npts, const float *xyz, const float *w, const float *bias, float *out, float *stats, int flags,
                              void *stream)
{
    PNAE_REQUIRE(npts >= 1 && xyz && w && bias && out && stats, "mlp_first: invalid argument");
    PNAE_REQUIRE(pnae_aligned(out, 16), "mlp_first: out must be 16-byte aligned");
    PNAE_REQUIRE((flags & ~(PNAE_STATS_ZEROED | PNAE_OVERLAP_PREVIOUS)) == 0, "mlp_first: unknown flag bits 0x%x", flags);
    cudaStream_t st = (cudaStream_t)stream;
    if (!(flags & PNAE_STATS_ZEROED)) PNAE_CUDA_OK(cudaMemsetAsync(stats, 0, sizeof(float) * 2 * 64, st));
    const int blocks = (int)min((npts + 63) / 64, (long long)pnae_sm_count() * 8);
    PNAE_CUDA_OK(pnae_launch(mlp_first_kernel, dim3(blocks), dim3(256), 0, st, (flags & PNAE_OVERLAP_PREVIOUS) != 0, npts, xyz, w, bias, out, stats));
    return PNAE_OK;
}

static cudaError_t tc_configure()
{
    cudaError_t e = cudaFuncSetAttribute(mlp_layer_tc_kernel<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(mlp_layer_tc_kernel<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(mlp_layer_tc_kernel<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(mlp_layer_tc_kernel<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmem);
    return e;
}

static BnPrev make_bn(const float *stats, double count, const float *gamma, const float *beta, float *mm, float *mv, float eps,
                      float decay, int training)
{
    BnPrev bn;
    bn.stats = stats; bn.gamma = gamma; bn.beta = beta; bn.moving_mean = mm; bn.moving_var = mv;
    bn.inv_count = training ? (float)(1.0 / count) : 0.f; bn.eps = eps; bn.decay = decay; bn.training = training;
    return bn;
}

extern "C" int pnae_mlp_layer(long long npts, int kin, int kout, const float *in,
                              const float *stats_prev, const float *gamma_prev, const float *beta_prev,
                              float *moving_mean_prev, float *moving_var_prev, float eps, float decay, int training,
                              const float *w, const float *bias, float *out, float *stats, int flags, void *stream)
{
    PNAE_REQUIRE((flags & ~(PNAE_STATS_ZEROED | PNAE_OVERLAP_PREVIOUS)) == 0, "mlp_layer: unknown flag bits 0x%x", flags);
    PNAE_REQUIRE(npts >= 1 && in && gamma_prev && beta_prev && moving_mean_prev && moving_var_prev && w && bias && out && stats && (!training || stats_prev),
                 "mlp_layer: invalid argument");
    PNAE_REQUIRE(kin == kKin && (kout == 64 || kout == 128), "mlp_layer: needs 64 input channels and 64 or 128 output channels (got %d -> %d)", kin, kout);
    PNAE_REQUIRE(pnae_aligned(in, 16) && pnae_aligned(out, 8), "mlp_layer: in must be 16-byte and out 8-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    static bool configured[64] = {false};
    int dev = 0;
    PNAE_CUDA_OK(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        PNAE_CUDA_OK(tc_configure());
        configured[dev] = true;
    }
    if (!(flags & PNAE_STATS_ZEROED)) PNAE_CUDA_OK(cudaMemsetAsync(stats, 0, sizeof(float) * (2 * kout + kout / 64), st));
    const BnPrev bn = make_bn(stats_prev, (double)npts, gamma_prev, beta_prev, moving_mean_prev, moving_var_prev, eps, decay, training);
    const long long ntiles = (npts + kTcPoints - 1) / kTcPoints;
    const int gx = (int)min(ntiles, (long long)pnae_sm_count());            // persistent: one CTA per SM, tiles strided
    const bool pdl = (flags & PNAE_OVERLAP_PREVIOUS) != 0;
    const FirstLayer none = {nullptr, nullptr, nullptr, nullptr};
    if (kout == 64) PNAE_CUDA_OK(pnae_launch(mlp_layer_tc_kernel<64, false>, dim3(gx), dim3(kTcThreads), kTcSmem, st, pdl, npts, in, none, bn, w, bias, out, stats));
    else PNAE_CUDA_OK(pnae_launch(mlp_layer_tc_kernel<128, false>, dim3(gx), dim3(kTcThreads), kTcSmem, st, pdl, npts, in, none, bn, w, bias, out, stats));
    return PNAE_OK;
}

extern "C" int pnae_xyz_moments(long long npts, const float *xyz, double *moments, int flags, void *stream)
{
    PNAE_REQUIRE((flags & ~(PNAE_STATS_ZEROED | PNAE_OVERLAP_PREVIOUS)) == 0, "xyz_moments: unknown flag bits 0x%x", flags);
    PNAE_REQUIRE(npts >= 1 && xyz && moments && pnae_aligned(moments, 8), "xyz_moments: invalid argument");
    cudaStream_t st = (cudaStream_t)stream;
    if (!(flags & PNAE_STATS_ZEROED)) PNAE_CUDA_OK(cudaMemsetAsync(moments, 0, sizeof(double) * 9, st));
    const int blocks = (int)min((npts + 255) / 256, (long long)pnae_sm_count());
    PNAE_CUDA_OK(pnae_launch(xyz_moments_kernel, dim3(blocks), dim3(256), 0, st, (flags & PNAE_OVERLAP_PREVIOUS) != 0, npts, xyz, moments));
    return PNAE_OK;
}

extern "C" int pnae_mlp_layer_xyz(long long npts, const float *xyz, const double *moments, const float *w1, const float *b1,
                                  const float *gamma1, const float *beta1, float *moving_mean1, float *moving_var1,
                                  float eps, float decay, int training, int kout, const float *w, const float *bias,
                                  float *out, float *stats, int flags, void *stream)
{
    PNAE_REQUIRE((flags & ~(PNAE_STATS_ZEROED | PNAE_OVERLAP_PREVIOUS)) == 0, "mlp_layer_xyz: unknown flag bits 0x%x", flags);
    PNAE_REQUIRE(npts >= 1 && xyz && w1 && b1 && gamma1 && beta1 && moving_mean1 && moving_var1 && w && bias && out && stats && (!training || moments),
                 "mlp_layer_xyz: invalid argument");
    PNAE_REQUIRE(kout == 64 || kout == 128, "mlp_layer_xyz: needs 64 or 128 output channels (got %d)", kout);
    PNAE_REQUIRE(pnae_aligned(out, 8), "mlp_layer_xyz: out must be 8-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    static bool configured[64] = {false};
    int dev = 0;
    PNAE_CUDA_OK(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !configured[dev]) {
        PNAE_CUDA_OK(tc_configure());
        configured[dev] = true;
    }
    if (!(flags & PNAE_STATS_ZEROED)) PNAE_CUDA_OK(cudaMemsetAsync(stats, 0, sizeof(float) * (2 * kout + kout / 64), st));
    const BnPrev bn = make_bn(nullptr, (double)npts, gamma1, beta1, moving_mean1, moving_var1, eps, decay, training);
    const FirstLayer first = {xyz, w1, b1, moments};
    const long long ntiles = (npts + kTcPoints - 1) / kTcPoints;
    const int gx = (int)min(ntiles, (long long)pnae_sm_count());
    const bool pdl = (flags & PNAE_OVERLAP_PREVIOUS) != 0;
    if (kout == 64) PNAE_CUDA_OK(pnae_launch(mlp_layer_tc_kernel<64, true>, dim3(gx), dim3(kTcThreads), kTcSmem, st, pdl, npts, (const float *)nullptr, first, bn, w, bias, out, stats));
    else PNAE_CUDA_OK(pnae_launch(mlp_layer_tc_kernel<128, true>, dim3(gx), dim3(kTcThreads), kTcSmem, st, pdl, npts, (const float *)nullptr, first, bn, w, bias, out, stats));
    return PNAE_OK;
}

extern "C" int pnae_bn_fold(int k, const float *stats, double count, const float *gamma, const float *beta, float eps, float decay,
                            int training, float *moving_mean, float *moving_var, float *s_out, float *t_out, void *stream)
{
    PNAE_REQUIRE(k >= 1 && gamma && beta && moving_mean && moving_var && s_out && t_out && (!training || (stats && count >= 1.0)),
                 "bn_fold: invalid argument");
    bn_fold_kernel<<<(k + 127) / 128, 128, 0, (cudaStream_t)stream>>>(k, stats, training ? (float)(1.0 / count) : 0.f, gamma, beta, eps, decay,
                                                                      training, moving_mean, moving_var, s_out, t_out);
    PNAE_CUDA_OK(cudaGetLastError());
    return PNAE_OK;
}

extern "C" int pnae_mlp_apply_bf16(long long npts, int k, const float *in, const float *stats, const float *gamma, const float *beta,
                                   float *moving_mean, float *moving_var, float eps, float decay, int training, void *out_bf16, int flags,
                                   void *stream)
{
    PNAE_REQUIRE((flags & ~PNAE_OVERLAP_PREVIOUS) == 0, "mlp_apply_bf16: unknown flag bits 0x%x", flags);
    PNAE_REQUIRE(npts >= 1 && k >= 4 && k % 4 == 0 && k <= 256 && in && gamma && beta && moving_mean && moving_var && out_bf16 && (!training || stats),
                 "mlp_apply_bf16: invalid argument");
    PNAE_REQUIRE(pnae_aligned(in, 16) && pnae_aligned(out_bf16, 8), "mlp_apply_bf16: misaligned buffer");
    const long long n4 = npts * k / 4;
    const int blocks = (int)min((n4 + 255) / 256, (long long)pnae_sm_count() * 16);
    PNAE_CUDA_OK(pnae_launch(mlp_apply_bf16_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, (flags & PNAE_OVERLAP_PREVIOUS) != 0,
                             n4, k, in, make_bn(stats, (double)npts, gamma, beta, moving_mean, moving_var, eps, decay, training), (__nv_bfloat16 *)out_bf16));
    return PNAE_OK;
}

extern "C" int pnae_conv5_finish(int b, int c, double count, const float *vmax, const float *vmin, const float *vsum, const float *vsq,
                                 const float *bias, const float *gamma, const float *beta, float *moving_mean, float *moving_var,
                                 float eps, float decay, int training, float *pooled, float *inv, float *mean0, float *ext0, float *z, int flags,
                                 void *stream)
{
    PNAE_REQUIRE((flags & ~PNAE_OVERLAP_PREVIOUS) == 0, "conv5_finish: unknown flag bits 0x%x", flags);
    PNAE_REQUIRE(b >= 1 && c >= 1 && count >= 1.0 && vmax && vmin && vsum && vsq && bias && gamma && beta && moving_mean && moving_var && pooled && inv && mean0 && ext0 && z,
                 "conv5_finish: invalid argument");
    PNAE_CUDA_OK(pnae_launch(conv5_finish_kernel, dim3((c + 31) / 32), dim3(32 * kFinRows), 0, (cudaStream_t)stream, (flags & PNAE_OVERLAP_PREVIOUS) != 0,
                             b, c, vmax, vmin, vsum, vsq, bias, gamma, beta, moving_mean, moving_var, (float)(1.0 / count), eps, decay, training,
                             pooled, inv, mean0, ext0, z));
    return PNAE_OK;
}
