// encoder.cu -- the PointNet encoder's dominant layer + pooling as ONE tensor-core kernel (sm_100a).
//
// Reference path: get_model, models/model.py:57-66 -- conv5 (1x1 conv 128 -> 1024, i.e. a per-point
// linear map: utils/tf_util.py:155-185) -> bias -> BatchNorm -> ReLU -> max over the points
// (tf_util.max_pool2d, :368-391).  89% of the encoder's FLOPs are this one GEMM, and the reference
// writes and re-reads the (B, N, 1024) activation (268 MB at B=32) three or more times.
//
// Here:  D[channel, point] = W5^T[channel, :] . X[point, :]   (bf16 operands, fp32 accumulate)
//  * tcgen05.mma.cta_group::2 (UMMA 256 x 256 x 16 across a CTA pair = a cluster of two SMs): each CTA of the pair owns
//    one 128-channel block (its half of M, its own weights, its own TMEM accumulators) and loads only HALF of every
//    256-point X tile; the tensor cores read the other half from the partner's shared memory.  This halves the
//    L2 -> shared-memory traffic of X, which -- not the tensor pipe -- bounded the one-CTA version: eight channel blocks
//    re-reading X asked the L2 for 10 TB/s (tools/enc_trace.py), and it halves the shared memory
//    of a stage, which buys a five-tile-deep TMA pipeline;
//  * accumulators in TMEM, double-buffered (2 x 256 columns per CTA);
//  * channels are the M (TMEM lane) dimension, so each epilogue thread owns one channel and the
//    reduction over points is a private register reduction straight out of tcgen05.ld:
//    running max, min, sum and sum of squares per (batch element, channel);
//  * the activation never leaves the SM.  max and min are kept because
//    max_n relu(s*y_n + t) = relu(s*max_n y_n + t) for s >= 0 and relu(s*min_n y_n + t) for s < 0
//    (s, t = folded BatchNorm scale/shift), and sum / sum^2 are exactly the batch statistics
//    training-mode BatchNorm needs -- so BN (either mode) + ReLU + max-pool finish on a (B,1024)
//    tensor ("Training-mode BatchNorm blocks naive encoder fusion").
//  * warp-specialised: warp 0 = TMA producer, warps 1-2 = MMA issuers (leader CTA only, alternating tiles; warp 1 of
//    both CTAs allocates TMEM), warps 3..10 = epilogue: two warps per TMEM lane quadrant, each reducing half of a tile's
//    256 columns, merged through shared memory when an element ends; mbarrier pipelines smem<->MMA<->epilogue, the
//    MMA-side ones signalled in both CTAs by multicast tcgen05.commit;
//  * persistent: one CTA pair per SM pair takes an equal run of the (element, point tile) stream of its two channel blocks.
#include <cuda.h>
#include <cuda_bf16.h>

#include "pnae_common.cuh"

namespace {

constexpr int kTileM = 128;        // channels per CTA (UMMA M)
constexpr int kTileN = 256;        // points per MMA tile (UMMA N)
constexpr int kHalfTile = 128;     // points of a tile each CTA of the pair loads
constexpr int kTmemCols = 2 * kTileN;     // two accumulators
constexpr int kKBox = 64;          // bf16 elements per 128-byte swizzle row
constexpr int kMaxK = 128;
constexpr int kStages = 5;
constexpr int kEncThreads = 352;   // 11 warps: TMA, 2 x MMA, 8 epilogue (two per TMEM lane quadrant, half the columns each)
constexpr int kFirstEpiWarp = 3;
constexpr int kEpiThreads = 256;
constexpr int kSpinLimit = 1 << 26;

#ifdef PNAE_ENC_TRACE                  // tuning builds only: SM-clock timestamps of two CTAs' pipeline events
__device__ long long g_enc_trace[2][16][8];
#define ENC_TRACE(tile, field) do { if (trace_cta >= 0) g_enc_trace[trace_cta][tile][field] = clock64(); } while (0)
#else
#define ENC_TRACE(tile, field) do { } while (0)
#endif

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// bounded spin: a descriptor mistake must trap, not hang the GPU
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t done = 0;
    int spins = 0;
    while (true) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (done) break;
        if (++spins > kSpinLimit) __trap();
    }
}

__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(smem_dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}

// the CTA-pair forms: loads signal the LEADER's barrier (rank bit of the barrier address cleared), MMAs span both SMs,
// commits arrive on the barrier at the same offset in every CTA of the mask
__device__ __forceinline__ void tma_load_2d_pair(void *smem_dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(smem_dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar) & 0xFEFFFFFFu) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
// arrive on the barrier at the same offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t *bar, uint32_t rank)
{
    asm volatile("{\n\t.reg .b32 rem;\n\tmapa.shared::cluster.u32 rem, %0, %1;\n\tmbarrier.arrive.shared::cluster.b64 _, [rem];\n\t}"
                 ::"r"(smem_u32(bar)), "r"(rank) : "memory");
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// K-major, 128B-swizzled operand tile: rows are 128 bytes, 8-row groups are 1024 bytes apart
// (cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), layout [61,64)=2)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;                 // leading byte offset (unused for swizzled K-major) = 16 B
    d |= (uint64_t)(1024 >> 4) << 32;       // stride byte offset: 8 rows x 128 B
    d |= (uint64_t)1 << 46;                 // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                 // SWIZZLE_128B
    return d;
}

// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, M=128, N=256
__device__ __forceinline__ uint32_t umma_idesc_bf16(int m, int n)
{
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// issue only: the 32 destination registers are written asynchronously and must not be read before tmem_ld_wait()
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32])
{
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r[i]);
}

// two-wide fp32 accumulation (FADD2 / FFMA2): the epilogue is issue-bound, and these halve the sum / sum-of-squares share
__device__ __forceinline__ void add2(float2 &acc, float a, float b)
{
    asm("{\n\t.reg .b64 x, y;\n\tmov.b64 x, {%2, %3};\n\tmov.b64 y, {%0, %1};\n\tadd.rn.f32x2 y, y, x;\n\tmov.b64 {%0, %1}, y;\n\t}"
        : "+f"(acc.x), "+f"(acc.y) : "f"(a), "f"(b));
}
__device__ __forceinline__ void sqacc2(float2 &acc, float a, float b)
{
    asm("{\n\t.reg .b64 x, y;\n\tmov.b64 x, {%2, %3};\n\tmov.b64 y, {%0, %1};\n\tfma.rn.f32x2 y, x, x, y;\n\tmov.b64 {%0, %1}, y;\n\t}"
        : "+f"(acc.x), "+f"(acc.y) : "f"(a), "f"(b));
}

// Work decomposition.  A channel-block pair's (element, point tile) pairs form one stream of B * ceil(N/256) tiles; the
// SM pairs are divided evenly among the channel-block pairs and the clusters of a pair take equal contiguous runs of
// its stream: every SM gets the same number of tiles (+-1) whatever B is, pays the set-up (barriers, TMEM, weights,
// first loads) once, and keeps its weights throughout.  A run boundary cuts an element in parts: each part goes to a
// workspace slot, a ticket counts the parts, and the CTA that delivers the last one merges the slots IN SLOT ORDER (so
// the result does not depend on which CTA came last) and writes the output.
// ARG: also report, per (element, channel), the index of the first point attaining the extremum
// the max-pool will select (the maximum where sign[channel] >= 0, the minimum otherwise): the
// training backward needs it (the pooled gradient flows to that point only).
struct EncTiles {
    int b, ntiles;
    int per_pair;          // clusters per channel-block pair
    int stream;            // tiles per pair = b * ntiles
    int pmax;              // slots per (element, channel)
    // (32-bit: the host checks stream * per_pair < 2^31)
    __device__ int first(int r) const { return (int)((unsigned)r * (unsigned)stream / (unsigned)per_pair); }
    __device__ int owner(int g) const { return (int)(((unsigned)(g + 1) * (unsigned)per_pair - 1u) / (unsigned)stream); }
};

template <bool ARG>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kEncThreads, 1)
encoder_conv_pool_kernel(const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_x,
                         int n, int k, int c, EncTiles tiles, float *__restrict__ omax, float *__restrict__ omin,
                         float *__restrict__ osum, float *__restrict__ osq,
                         const float *__restrict__ sign, int *__restrict__ oarg,
                         float *__restrict__ slots, int *__restrict__ tickets)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int kboxes = k / kKBox;
    const uint32_t a_bytes = (uint32_t)kboxes * kTileM * 128;           // this CTA's W block
    const uint32_t b_bytes = (uint32_t)kboxes * kHalfTile * 128;        // this CTA's half of one X stage
    uint8_t *sa = smem;
    uint8_t *sb = smem + a_bytes;
    uint64_t *bars = reinterpret_cast<uint64_t *>(sb + kStages * b_bytes);
    uint64_t *w_full = bars, *b_full = bars + 1, *b_empty = b_full + kStages;
    uint64_t *t_full = b_empty + kStages, *t_empty = t_full + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(t_empty + 2);
    int *flag = reinterpret_cast<int *>(tmem_slot) + 1;            // [2]: the ticket values, broadcast to the epilogue warps
    float *merge = reinterpret_cast<float *>(bars) + 64;           // [6][128]: the second half's partial results (256 B past the barriers)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t cta_rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
    const bool leader = cta_rank == 0;
    const int ntiles = tiles.ntiles;
    const int cluster = blockIdx.x >> 1;
    const int pair = cluster / tiles.per_pair, rank = cluster % tiles.per_pair;
    const int cb = pair * 2 + (int)cta_rank;
    const bool live = cb * kTileM < c;                             // an odd number of channel blocks leaves the last CTA a bystander
    const int g0 = tiles.first(rank), len = tiles.first(rank + 1) - g0;
    const int t0 = g0 % ntiles, e0 = g0 / ntiles;
#ifdef PNAE_ENC_TRACE
    const int trace_cta = blockIdx.x == 0 ? 0 : blockIdx.x == 1 ? 1 : -1;
    if (threadIdx.x == 0) ENC_TRACE(15, 0);
#endif

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_w) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_x) : "memory");
        mbar_init(w_full, 1);
        for (int s = 0; s < kStages; s++) { mbar_init(b_full + s, 1); mbar_init(b_empty + s, 1); }
        for (int s = 0; s < 2; s++) { mbar_init(t_full + s, 1); mbar_init(t_empty + s, 2 * kEpiThreads / 32); }   // one arrival per epilogue warp of either CTA
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();                                            // barriers and TMEM of BOTH CTAs exist from here on
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) ENC_TRACE(15, 1);

    if (warp == 0) {
        // ===== TMA producer (both CTAs): this CTA's weights once, then its half of every point tile =====
        if (lane == 0) {
            if (leader) mbar_expect_tx(w_full, 2 * a_bytes);
            for (int kb = 0; kb < kboxes; kb++) tma_load_2d_pair(sa + (size_t)kb * kTileM * 128, &tm_w, kb * kKBox, cb * kTileM, w_full);
            int t = t0, e = e0;
            for (int lt = 0; lt < len; lt++) {
                const int s = lt % kStages;
                if (lt >= kStages) mbar_wait(b_empty + s, ((lt / kStages) - 1) & 1);
                if (lt < 15) ENC_TRACE(lt, 0);
                if (leader) mbar_expect_tx(b_full + s, 2 * b_bytes);
                for (int kb = 0; kb < kboxes; kb++)
                    tma_load_2d_pair(sb + (size_t)s * b_bytes + (size_t)kb * kHalfTile * 128, &tm_x, kb * kKBox,
                                     e * n + t * kTileN + (int)cta_rank * kHalfTile, b_full + s);
                if (++t == ntiles) { t = 0; e++; }
            }
        }
    } else if (warp <= 2) {
        // ===== two MMA issuers in the leader CTA, alternating tiles (warp 1: even, warp 2: odd; each owns one
        // accumulator buffer): while one sits in its barrier waits the other's MMAs keep the tensor pipes busy =====
        if (lane == 0 && leader) {
            const uint32_t idesc = umma_idesc_bf16(2 * kTileM, kTileN);
#ifdef PNAE_ENC_ONE_ISSUER
            if (warp == 1)
            for (int lt = 0; lt < len; lt++) {
                const int buf = lt & 1;
                if (lt == 0) mbar_wait(w_full, 0);
#else
            const int buf = warp - 1;
            mbar_wait(w_full, 0);
            for (int lt = buf; lt < len; lt += 2) {
#endif
                const int s = lt % kStages;
                mbar_wait(b_full + s, (lt / kStages) & 1);                       // (long there: the loads run five tiles ahead)
                if (lt < 15) ENC_TRACE(lt, 1);
                if (lt >= 2) mbar_wait(t_empty + buf, ((lt >> 1) - 1) & 1);      // both epilogues drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (lt < 15) ENC_TRACE(lt, 2);
                for (int kk = 0; kk < k / 16; kk++) {
                    const int kb = kk >> 2, kin = kk & 3;                      // 4 UMMA_K=16 steps per 128-byte row
                    const uint64_t ad = umma_desc_sw128(smem_u32(sa + (size_t)kb * kTileM * 128) + kin * 32);
                    const uint64_t bd = umma_desc_sw128(smem_u32(sb + (size_t)s * b_bytes + (size_t)kb * kHalfTile * 128) + kin * 32);
                    umma_bf16_pair(tmem_base + buf * kTileN, ad, bd, idesc, kk > 0);
                }
                umma_commit_pair(b_empty + s);      // smem stage reusable (in both CTAs) once these MMAs retire
                umma_commit_pair(t_full + buf);     // accumulators ready for both epilogues
                if (lt < 15) ENC_TRACE(lt, 3);
            }
        }
    } else {
        // ===== epilogue: two threads per channel (one per column half), reduction over points in registers =====
        const int q = warp & 3;                                   // TMEM lane quadrant this warp may read
        const int half = (warp - kFirstEpiWarp) >> 2;             // columns [half*128, half*128 + 128) of every tile
        const int etid = threadIdx.x - kFirstEpiWarp * 32;
        constexpr int kHalfN = kTileN / 2;
        const float inf = __int_as_float(0x7f800000);
        float vmax = -inf, vmin = inf, kbest = -inf;
        float2 vsum2 = make_float2(0.f, 0.f), vsq2 = make_float2(0.f, 0.f);     // even / odd columns
        int ibest = 0;
        const int slot = q * 32 + lane;
        const int ch_out = cb * kTileM + slot;
        int t = t0, e = e0;
        int t_first = t0;                                         // first tile of the part being accumulated
        int part_e[2] = {0, 0}, part_n[2] = {0, 0}, nparts = 0;                     // the parts this run leaves in workspace slots: element, number of parts
        const size_t plane = (size_t)tiles.b * c;
        // key = +v (track the maximum) or -v (track the minimum): flipping the sign bit is exact
        const unsigned flip = (ARG && live && sign[ch_out] < 0.f) ? 0x80000000u : 0u;
        for (int lt = 0; lt < len; lt++) {
            const int buf = lt & 1;
            const int valid = min(kTileN, n - t * kTileN);
            mbar_wait(t_full + buf, (lt >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lt < 15 && etid == 0) ENC_TRACE(lt, 4);
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * kTileN + half * kHalfN;
            if (live && half * kHalfN < valid) {
                // all four 32-column loads of this half in flight, ONE wait: a load + wait per group left the TMEM read
                // latency (about twice the group's arithmetic while the MMAs of the other accumulator run) exposed four times
                uint32_t r[kHalfN / 32][32];
#pragma unroll
                for (int ch = 0; ch < kHalfN / 32; ch++) tmem_ld32_issue(taddr + ch * 32, r[ch]);
                tmem_ld_wait();
                if (lt < 15 && etid == 0) ENC_TRACE(lt, 6);
#pragma unroll
                for (int ch = 0; ch < kHalfN / 32; ch++) {
                    const int col0 = half * kHalfN + ch * 32;     // first column of this 32-column group inside the tile
                    if (col0 + 32 <= valid) {
#pragma unroll
                        for (int i = 0; i < 32; i += 2) {
                            const float x0 = __uint_as_float(r[ch][i]), x1 = __uint_as_float(r[ch][i + 1]);
                            vmax = fmaxf(vmax, fmaxf(x0, x1)); vmin = fminf(vmin, fminf(x0, x1));
                            add2(vsum2, x0, x1); sqacc2(vsq2, x0, x1);
                        }
                        if (ARG) {
#pragma unroll
                            for (int i = 0; i < 32; i++) {
                                const float key = __uint_as_float(r[ch][i] ^ flip);
                                if (key > kbest) { kbest = key; ibest = t * kTileN + col0 + i; }   // strict: first point wins
                            }
                        }
                    } else if (col0 < valid) {
#pragma unroll
                        for (int i = 0; i < 32; i++)
                            if (col0 + i < valid) {
                                const float x = __uint_as_float(r[ch][i]);
                                vmax = fmaxf(vmax, x); vmin = fminf(vmin, x);
                                vsum2.x += x; vsq2.x = fmaf(x, x, vsq2.x);
                                if (ARG) {
                                    const float key = __uint_as_float(r[ch][i] ^ flip);
                                    if (key > kbest) { kbest = key; ibest = t * kTileN + col0 + i; }
                                }
                            }
                    }
                }
            }
            if (lt < 15 && etid == 0) ENC_TRACE(lt, 7);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(t_empty + buf, 0);          // the leader's barrier: its MMA issuers wait for both CTAs
            if (lt < 15 && etid == 0) ENC_TRACE(lt, 5);

            if (t == ntiles - 1 || lt == len - 1) {
                // ----- flush this element's part -----
                // the two column halves of every channel first: the second half hands its partial results over in shared memory
                float vsum = vsum2.x + vsum2.y, vsq = vsq2.x + vsq2.y;
                if (half == 1) {
                    merge[slot] = vmax; merge[128 + slot] = vmin; merge[256 + slot] = vsum; merge[384 + slot] = vsq;
                    merge[512 + slot] = kbest; merge[640 + slot] = __int_as_float(ibest);
                }
                asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");          // the eight epilogue warps only
                const size_t o = (size_t)e * c + ch_out;
                const bool mine = live && half == 0;
                if (half == 0) {
                    vmax = fmaxf(vmax, merge[slot]); vmin = fminf(vmin, merge[128 + slot]);
                    vsum += merge[256 + slot]; vsq += merge[384 + slot];
                    if (ARG) {
                        const float k1 = merge[512 + slot];
                        const int i1 = __float_as_int(merge[640 + slot]);
                        if (k1 > kbest || (k1 == kbest && i1 < ibest)) { kbest = k1; ibest = i1; }   // the extremum's FIRST point, whichever half saw it
                    }
                }
                if (t_first == 0 && t == ntiles - 1) {            // the whole element was ours
                    if (mine) {
                        omax[o] = vmax; omin[o] = vmin; osum[o] = vsum; osq[o] = vsq;
                        if (ARG) oarg[o] = ibest;
                    }
                } else {
                    // a part (at most two per run: the tail of the element the run starts in, the head of the one it
                    // ends in): write it to its slot now -- slot = this cluster's position among the clusters sharing
                    // the element -- and settle the tickets once, after the run
                    const int gfirst = g0 + lt - t;
                    const int jfirst = tiles.owner(gfirst);
                    if (mine) {
                        float *sl = slots + (size_t)(rank - jfirst) * 6 * plane + o;
                        sl[0] = vmax; sl[plane] = vmin; sl[2 * plane] = vsum; sl[3 * plane] = vsq;
                        sl[4 * plane] = kbest; sl[5 * plane] = __int_as_float(ibest);
                    }
                    const int np = tiles.owner(gfirst + ntiles - 1) - jfirst + 1;
                    if (nparts == 0) { part_e[0] = e; part_n[0] = np; } else { part_e[1] = e; part_n[1] = np; }
                    nparts++;
                }
                asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");          // `merge` may be rewritten
                vmax = -inf; vmin = inf; vsum2 = make_float2(0.f, 0.f); vsq2 = make_float2(0.f, 0.f); kbest = -inf; ibest = 0;
                t_first = 0;
            }
            if (++t == ntiles) { t = 0; e++; }
        }
        if (nparts > 0) {
            // the last barrier ordered every thread's slot writes before this point
            if (etid == 0) ENC_TRACE(14, 1);
            if (etid == 0) {
                __threadfence();        // cumulative: publishes the slot writes the barrier ordered before this thread
                int old[2] = {-1, -1};
#pragma unroll
                for (int i = 0; i < 2; i++) {
                    if (i >= nparts) break;
                    int *ticket = tickets + (size_t)part_e[i] * (c / kTileM) + min(cb, c / kTileM - 1);
                    if (live) old[i] = atomicAdd(ticket, 1);
                    if (old[i] == part_n[i] - 1) *ticket = 0;                    // last part in: leave the ticket ready for the next call
                }
                flag[0] = old[0]; flag[1] = old[1];
                __threadfence();
            }
            asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
            if (etid == 0) ENC_TRACE(14, 2);
#pragma unroll
            for (int i = 0; i < 2; i++)
                if (i < nparts && flag[i] == part_n[i] - 1 && live && half == 0) {
                    const size_t o = (size_t)part_e[i] * c + ch_out;
                    float m1 = -inf, m0 = inf, su = 0.f, sq = 0.f, kb = -inf;
                    int ib = 0;
                    for (int p = 0; p < part_n[i]; p++) {                        // slot order = point order
                        const float *sl = slots + (size_t)p * 6 * plane + o;
                        m1 = fmaxf(m1, __ldcg(sl)); m0 = fminf(m0, __ldcg(sl + plane));
                        su += __ldcg(sl + 2 * plane); sq += __ldcg(sl + 3 * plane);
                        if (ARG) {
                            const float k1 = __ldcg(sl + 4 * plane);
                            if (k1 > kb) { kb = k1; ib = __float_as_int(__ldcg(sl + 5 * plane)); }
                        }
                    }
                    omax[o] = m1; omin[o] = m0; osum[o] = su; osq[o] = sq;
                    if (ARG) oarg[o] = ib;
                }
            if (etid == 0) ENC_TRACE(14, 3);
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();                                            // neither CTA may leave while the other can still signal it or read its operands
    if (threadIdx.x == 96) ENC_TRACE(14, 5);
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTmemCols) : "memory");
    }
    if (threadIdx.x == 32) ENC_TRACE(15, 2);
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

#ifdef PNAE_ENC_TRACE
extern "C" __attribute__((visibility("default"))) int pnae_debug_enc_trace(long long *host)
{
    return (int)cudaMemcpyFromSymbol(host, g_enc_trace, sizeof(g_enc_trace));
}
#endif

namespace {

// the launch geometry of pnae_encoder_conv_pool (see EncTiles)
int enc_tiles(int b, int n, int c, EncTiles *out, int *pairs)
{
    int dev = 0, sms = 0;
    PNAE_CUDA_OK(cudaGetDevice(&dev));
    PNAE_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    EncTiles t;
    t.b = b;
    t.ntiles = (n + kTileN - 1) / kTileN;
    *pairs = (c / kTileM + 1) / 2;
    const long long stream = (long long)b * t.ntiles;
    PNAE_REQUIRE(stream < (1ll << 30), "encoder_conv_pool: batch * points too large");
    t.stream = (int)stream;
    t.per_pair = (sms / 2) / *pairs;                               // one cluster per SM pair, shared evenly by the channel-block pairs
    if (t.per_pair > t.stream) t.per_pair = t.stream;
    if (t.per_pair < 1) t.per_pair = 1;
    PNAE_REQUIRE(stream * t.per_pair < (1ll << 31), "encoder_conv_pool: batch * points too large");
    const int shortest = t.stream / t.per_pair;                    // >= 1
    int pmax = (t.ntiles - 1 + shortest - 1) / shortest + 1;
    if (pmax > t.ntiles) pmax = t.ntiles;
    t.pmax = pmax;
    *out = t;
    return PNAE_OK;
}

}  // namespace

extern "C" int pnae_encoder_conv_pool_workspace_bytes(int b, int n, int c, size_t *bytes)
{
    PNAE_REQUIRE(bytes != nullptr, "encoder_conv_pool_workspace_bytes: NULL pointer");
    PNAE_REQUIRE(b >= 0 && n >= 1 && c >= kTileM && c % kTileM == 0, "encoder_conv_pool_workspace_bytes: need b>=0, n>=1, c a multiple of 128");
    *bytes = 0;
    if (b == 0) return PNAE_OK;
    EncTiles t;
    int pairs = 0;
    int rc = enc_tiles(b, n, c, &t, &pairs);
    if (rc) return rc;
    *bytes = ((size_t)t.pmax * 6 * b * c + (size_t)b * (c / kTileM)) * sizeof(float);
    return PNAE_OK;
}

extern "C" int pnae_encoder_conv_pool(int b, int n, int k, int c, const void *x_bf16, const void *wt_bf16,
                                      float *out_max, float *out_min, float *out_sum, float *out_sumsq,
                                      const float *sign, int *out_arg, void *workspace, size_t workspace_bytes, void *stream)
{
    PNAE_REQUIRE((sign == nullptr) == (out_arg == nullptr), "encoder_conv_pool: pass both `sign` and `out_arg` or neither");
    PNAE_REQUIRE(b >= 0 && n >= 1, "encoder_conv_pool: need b>=0, n>=1 (got b=%d n=%d)", b, n);
    PNAE_REQUIRE(k >= kKBox && k <= kMaxK && k % kKBox == 0, "encoder_conv_pool: in-channels must be 64 or 128 (got %d)", k);
    PNAE_REQUIRE(c >= kTileM && c % kTileM == 0, "encoder_conv_pool: out-channels must be a multiple of 128 (got %d)", c);
    PNAE_REQUIRE(x_bf16 && wt_bf16 && out_max && out_min && out_sum && out_sumsq, "encoder_conv_pool: NULL pointer");
    PNAE_REQUIRE(pnae_aligned(x_bf16, 16) && pnae_aligned(wt_bf16, 16), "encoder_conv_pool: operands must be 16-byte aligned");
    PNAE_REQUIRE((long long)b * n < (1ll << 31), "encoder_conv_pool: batch * points must be below 2^31");
    if (b == 0) return PNAE_OK;
    EncTiles tiles;
    int pairs = 0;
    int rc = enc_tiles(b, n, c, &tiles, &pairs);
    if (rc) return rc;
    const size_t slot_floats = (size_t)tiles.pmax * 6 * b * c;
    const size_t need = (slot_floats + (size_t)b * (c / kTileM)) * sizeof(float);
    PNAE_REQUIRE(workspace != nullptr && workspace_bytes >= need && pnae_aligned(workspace, 4),
                 "encoder_conv_pool: workspace of %zu bytes needed (pnae_encoder_conv_pool_workspace_bytes), got %zu", need, workspace_bytes);
    float *slots = static_cast<float *>(workspace);
    int *tickets = reinterpret_cast<int *>(slots + slot_floats);
    CUtensorMap tm_w, tm_x;
    rc = make_map(&tm_w, wt_bf16, (uint64_t)c, (uint64_t)k, kTileM);
    if (rc) return rc;
    rc = make_map(&tm_x, x_bf16, (uint64_t)b * n, (uint64_t)k, kHalfTile);
    if (rc) return rc;
    const int kboxes = k / kKBox;
    const size_t smem = 1024 + (size_t)kboxes * kTileM * 128 + (size_t)kStages * kboxes * kHalfTile * 128 + 256 + 6 * 128 * sizeof(float);
    const unsigned ctas = 2u * (unsigned)pairs * tiles.per_pair;       // clusters of two
    if (out_arg) {
        PNAE_CUDA_OK(cudaFuncSetAttribute(encoder_conv_pool_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        encoder_conv_pool_kernel<true><<<ctas, kEncThreads, smem, (cudaStream_t)stream>>>(tm_w, tm_x, n, k, c, tiles, out_max, out_min, out_sum, out_sumsq, sign, out_arg, slots, tickets);
    } else {
        PNAE_CUDA_OK(cudaFuncSetAttribute(encoder_conv_pool_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        encoder_conv_pool_kernel<false><<<ctas, kEncThreads, smem, (cudaStream_t)stream>>>(tm_w, tm_x, n, k, c, tiles, out_max, out_min, out_sum, out_sumsq, nullptr, nullptr, slots, tickets);
    }
    PNAE_CUDA_OK(cudaGetLastError());
    return PNAE_OK;
}
