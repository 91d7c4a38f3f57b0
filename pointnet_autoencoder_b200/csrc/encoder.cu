// encoder.cu -- the PointNet encoder's dominant layer + pooling as ONE tensor-core kernel (sm_100a).
//
// Reference path: get_model, models/model.py:57-66 -- conv5 (1x1 conv 128 -> 1024, i.e. a per-point
// linear map: utils/tf_util.py:155-185) -> bias -> BatchNorm -> ReLU -> max over the points
// (tf_util.max_pool2d, :368-391).  89% of the encoder's FLOPs are this one GEMM, and the reference
// writes and re-reads the (B, N, 1024) activation (268 MB at B=32) three or more times.
//
// Here:  D[channel, point] = W5^T[channel, :] . X[point, :]   (bf16 operands, fp32 accumulate)
//  * tcgen05.mma (UMMA 128 x 256 x 16, cta_group::1), operands staged by TMA into 128B-swizzled
//    shared memory, accumulators double-buffered in TMEM (2 x 256 columns);
//  * channels are the M (TMEM lane) dimension, so each epilogue thread owns one channel and the
//    reduction over points is a private register reduction straight out of tcgen05.ld:
//    running max, min, sum and sum of squares per (batch element, channel);
//  * the activation never leaves the SM.  max and min are kept because
//    max_n relu(s*y_n + t) = relu(s*max_n y_n + t) for s >= 0 and relu(s*min_n y_n + t) for s < 0
//    (s, t = folded BatchNorm scale/shift), and sum / sum^2 are exactly the batch statistics
//    training-mode BatchNorm needs -- so BN (either mode) + ReLU + max-pool finish on a (B,1024)
//    tensor ("Training-mode BatchNorm blocks naive encoder fusion").
//  * warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM allocation),
//    warps 2..9 = epilogue: two warps per TMEM lane quadrant, each reducing half of a tile's 256 columns (with four
//    epilogue warps the 256-value reduction per channel took as long as the tile's MMAs: K is only 128), merged through
//    shared memory at the end; mbarrier pipelines smem<->MMA<->epilogue.
#include <cuda.h>
#include <cuda_bf16.h>

#include "pnae_common.cuh"
#include "pnae_tc.cuh"

namespace {

constexpr int kTileM = 128;        // channels per CTA (UMMA M)
constexpr int kTileN = 256;        // points per MMA tile (UMMA N)
constexpr int kKBox = 64;          // bf16 elements per 128-byte swizzle row
constexpr int kMaxK = 128;
constexpr int kStages = 2;
constexpr int kEncThreads = 320;   // 10 warps: TMA, MMA, 8 epilogue (two per TMEM lane quadrant, half the columns each)
constexpr int kEpiThreads = 256;
// grid = (C/128) x B.  Each CTA: 128 channels of one batch element, looping over its point tiles.
// ARG: also report, per (element, channel), the index of the first point attaining the extremum
// the max-pool will select (the maximum where sign[channel] >= 0, the minimum otherwise): the
// training backward needs it (the pooled gradient flows to that point only).
template <bool ARG>
__global__ void __launch_bounds__(kEncThreads, 1)
encoder_conv_pool_kernel(const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_x,
                         int n, int k, int c, float *__restrict__ omax, float *__restrict__ omin,
                         float *__restrict__ osum, float *__restrict__ osq,
                         const float *__restrict__ sign, int *__restrict__ oarg)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int kboxes = k / kKBox;
    const uint32_t a_bytes = (uint32_t)kboxes * kTileM * 128;           // W block
    const uint32_t b_bytes = (uint32_t)kboxes * kTileN * 128;           // one X stage
    uint8_t *sa = smem;
    uint8_t *sb = smem + a_bytes;
    uint64_t *bars = reinterpret_cast<uint64_t *>(sb + kStages * b_bytes);
    uint64_t *a_full = bars, *b_full = bars + 1, *b_empty = b_full + kStages;
    uint64_t *t_full = b_empty + kStages, *t_empty = t_full + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(t_empty + 2);
    float *merge = reinterpret_cast<float *>(bars) + 64;           // [6][128]: the second half's partial results (256 B past the barriers)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cb = blockIdx.x, e = blockIdx.y;
    const int ntiles = (n + kTileN - 1) / kTileN;

    pnae_pdl_release();
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_w) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_x) : "memory");
        mbar_init(a_full, 1);
        for (int s = 0; s < kStages; s++) { mbar_init(b_full + s, 1); mbar_init(b_empty + s, 1); }
        for (int s = 0; s < 2; s++) { mbar_init(t_full + s, 1); mbar_init(t_empty + s, kEpiThreads); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            mbar_expect_tx(a_full, a_bytes);
            for (int kb = 0; kb < kboxes; kb++) tma_load_2d(sa + (size_t)kb * kTileM * 128, &tm_w, kb * kKBox, cb * kTileM, a_full);
            pnae_pdl_wait();              // the weights may load under the previous kernel's tail (PNAE_OVERLAP_PREVIOUS); X is its output
            for (int t = 0; t < ntiles; t++) {
                const int s = t % kStages;
                if (t >= kStages) mbar_wait(b_empty + s, ((t / kStages) - 1) & 1);
                mbar_expect_tx(b_full + s, b_bytes);
                for (int kb = 0; kb < kboxes; kb++)
                    tma_load_2d(sb + (size_t)s * b_bytes + (size_t)kb * kTileN * 128, &tm_x, kb * kKBox, e * n + t * kTileN, b_full + s);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_bf16(kTileM, kTileN);
            mbar_wait(a_full, 0);
            for (int t = 0; t < ntiles; t++) {
                const int s = t % kStages, buf = t & 1;
                if (t >= 2) mbar_wait(t_empty + buf, ((t >> 1) - 1) & 1);      // epilogue drained this accumulator
                mbar_wait(b_full + s, (t / kStages) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int kk = 0; kk < k / 16; kk++) {
                    const int kb = kk >> 2, kin = kk & 3;                      // 4 UMMA_K=16 steps per 128-byte row
                    const uint64_t ad = umma_desc_sw128(smem_u32(sa + (size_t)kb * kTileM * 128) + kin * 32);
                    const uint64_t bd = umma_desc_sw128(smem_u32(sb + (size_t)s * b_bytes + (size_t)kb * kTileN * 128) + kin * 32);
                    umma_bf16(tmem_base + buf * kTileN, ad, bd, idesc, kk > 0);
                }
                umma_commit(b_empty + s);       // smem stage reusable once these MMAs retire
                umma_commit(t_full + buf);      // accumulator ready for the epilogue
            }
        }
    } else {
        // ===== epilogue: two threads per channel (one per column half), reduction over points in registers =====
        const int q = warp & 3;                                   // TMEM lane quadrant this warp may read
        const int half = (warp - 2) >> 2;                         // columns [half*128, half*128 + 128) of every tile
        constexpr int kHalfN = kTileN / 2;
        float vmax = -__int_as_float(0x7f800000), vmin = __int_as_float(0x7f800000), vsum = 0.f, vsq = 0.f;
        const int ch_mine = cb * kTileM + q * 32 + lane;
        // key = +v (track the maximum) or -v (track the minimum): flipping the sign bit is exact
        pnae_pdl_wait();                  // (`sign` is a parameter, but the outputs below are the previous kernel's to finish with first)
        const unsigned flip = (ARG && ch_mine < c && sign[ch_mine] < 0.f) ? 0x80000000u : 0u;
        float kbest = -__int_as_float(0x7f800000);
        int ibest = 0;
        for (int t = 0; t < ntiles; t++) {
            const int buf = t & 1;
            const int valid = min(kTileN, n - t * kTileN);
            mbar_wait(t_full + buf, (t >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * kTileN + half * kHalfN;
#pragma unroll 1
            for (int ch = 0; ch < kHalfN / 32; ch++) {
                const int col0 = half * kHalfN + ch * 32;         // first column of this 32-column group inside the tile
                if (col0 >= valid) break;                         // (warp-uniform) nothing live in the rest of this half
                float v[32];
                tmem_ld32(taddr + ch * 32, v);
                if (col0 + 32 <= valid) {
#pragma unroll
                    for (int i = 0; i < 32; i++) {
                        vmax = fmaxf(vmax, v[i]); vmin = fminf(vmin, v[i]);
                        vsum += v[i]; vsq = fmaf(v[i], v[i], vsq);
                        if (ARG) {
                            const float key = __uint_as_float(__float_as_uint(v[i]) ^ flip);
                            if (key > kbest) { kbest = key; ibest = t * kTileN + col0 + i; }   // strict: first point wins
                        }
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 32; i++)
                        if (col0 + i < valid) {
                            vmax = fmaxf(vmax, v[i]); vmin = fminf(vmin, v[i]);
                            vsum += v[i]; vsq = fmaf(v[i], v[i], vsq);
                            if (ARG) {
                                const float key = __uint_as_float(__float_as_uint(v[i]) ^ flip);
                                if (key > kbest) { kbest = key; ibest = t * kTileN + col0 + i; }
                            }
                        }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(t_empty + buf);
        }
        // merge the two column halves of every channel: the second half hands its partial results over in shared memory
        const int slot = q * 32 + lane;
        if (half == 1) {
            merge[slot] = vmax; merge[128 + slot] = vmin; merge[256 + slot] = vsum; merge[384 + slot] = vsq;
            merge[512 + slot] = kbest; merge[640 + slot] = __int_as_float(ibest);
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");          // the eight epilogue warps only
        const int ch_out = cb * kTileM + q * 32 + lane;
        if (half == 0 && ch_out < c) {
            vmax = fmaxf(vmax, merge[slot]); vmin = fminf(vmin, merge[128 + slot]);
            vsum += merge[256 + slot]; vsq += merge[384 + slot];
            const size_t o = (size_t)e * c + ch_out;
            omax[o] = vmax; omin[o] = vmin; osum[o] = vsum; osq[o] = vsq;
            if (ARG) {
                const float k1 = merge[512 + slot];
                const int i1 = __float_as_int(merge[640 + slot]);
                if (k1 > kbest || (k1 == kbest && i1 < ibest)) ibest = i1;       // the extremum's FIRST point, whichever half saw it
                oarg[o] = ibest;
            }
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// rows x k bf16, row-major (k contiguous): box = 64 elements (128 B) x box_rows, 128B swizzle
int make_map(CUtensorMap *map, const void *base, uint64_t rows, uint64_t k, uint32_t box_rows)
{
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) { pnae_set_error("cuTensorMapEncodeTiled is not available from this driver"); return PNAE_ERR_CUDA; }
    cuuint64_t dims[2] = {k, rows};
    cuuint64_t strides[1] = {k * sizeof(__nv_bfloat16)};
    cuuint32_t box[2] = {(cuuint32_t)kKBox, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { pnae_set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return PNAE_ERR_CUDA; }
    return PNAE_OK;
}

}  // namespace

extern "C" int pnae_encoder_conv_pool(int b, int n, int k, int c, const void *x_bf16, const void *wt_bf16,
                                      float *out_max, float *out_min, float *out_sum, float *out_sumsq,
                                      const float *sign, int *out_arg, int flags, void *stream)
{
    PNAE_REQUIRE((flags & ~PNAE_OVERLAP_PREVIOUS) == 0, "encoder_conv_pool: unknown flag bits 0x%x", flags);
    PNAE_REQUIRE((sign == nullptr) == (out_arg == nullptr), "encoder_conv_pool: pass both `sign` and `out_arg` or neither");
    PNAE_REQUIRE(b >= 0 && n >= 1, "encoder_conv_pool: need b>=0, n>=1 (got b=%d n=%d)", b, n);
    PNAE_REQUIRE(k >= kKBox && k <= kMaxK && k % kKBox == 0, "encoder_conv_pool: in-channels must be 64 or 128 (got %d)", k);
    PNAE_REQUIRE(c >= kTileM && c % kTileM == 0, "encoder_conv_pool: out-channels must be a multiple of 128 (got %d)", c);
    PNAE_REQUIRE(x_bf16 && wt_bf16 && out_max && out_min && out_sum && out_sumsq, "encoder_conv_pool: NULL pointer");
    PNAE_REQUIRE(pnae_aligned(x_bf16, 16) && pnae_aligned(wt_bf16, 16), "encoder_conv_pool: operands must be 16-byte aligned");
    if (b == 0) return PNAE_OK;
    CUtensorMap tm_w, tm_x;
    int rc = make_map(&tm_w, wt_bf16, (uint64_t)c, (uint64_t)k, kTileM);
    if (rc) return rc;
    rc = make_map(&tm_x, x_bf16, (uint64_t)b * n, (uint64_t)k, kTileN);
    if (rc) return rc;
    const int kboxes = k / kKBox;
    const size_t smem = 1024 + (size_t)kboxes * kTileM * 128 + (size_t)kStages * kboxes * kTileN * 128 + 256 + 6 * 128 * sizeof(float);
    dim3 grid((unsigned)(c / kTileM), (unsigned)b);
    if (out_arg) {
        PNAE_CUDA_OK(cudaFuncSetAttribute(encoder_conv_pool_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PNAE_CUDA_OK(pnae_launch(encoder_conv_pool_kernel<true>, grid, dim3(kEncThreads), smem, (cudaStream_t)stream, (flags & PNAE_OVERLAP_PREVIOUS) != 0,
                                 tm_w, tm_x, n, k, c, out_max, out_min, out_sum, out_sumsq, sign, out_arg));
    } else {
        PNAE_CUDA_OK(cudaFuncSetAttribute(encoder_conv_pool_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PNAE_CUDA_OK(pnae_launch(encoder_conv_pool_kernel<false>, grid, dim3(kEncThreads), smem, (cudaStream_t)stream, (flags & PNAE_OVERLAP_PREVIOUS) != 0,
                                 tm_w, tm_x, n, k, c, out_max, out_min, out_sum, out_sumsq, (const float *)nullptr, (int *)nullptr));
    }
    PNAE_CUDA_OK(cudaGetLastError());
    return PNAE_OK;
}
