// nn_distance.cu -- Chamfer distance forward + gradient for sm_100a.
//
// Replaces NmDistanceKernel / NmDistanceGradKernel and their launchers
// (reference: tf_ops/nn_distance/tf_nndistance_g.cu:5-157).  Semantics kept:
// squared distances with the FMUL(y)->FFMA(x)->FFMA(z) rounding of the reference
// SASS, strict '<' so the lowest index wins ties, gradient outputs zeroed inside.
//
// Forward design (DESIGN.md "Chamfer forward"):
//  * every UNORDERED pair (xyz1[j], xyz2[k]) is evaluated once and feeds both the
//    row minimum (dist1) and the column minimum (dist2): (a-b)^2 == (b-a)^2 bit for
//    bit, so this is exact and halves the FP32 work of the two reference launches;
//  * the inner loop tracks minima only (FMNMX); indices are recovered afterwards
//    from a coarse tag: the 32-column chunk in which a row's minimum first appeared,
//    and the R-row group (one lane's rows) that produced a column's minimum;
//  * sweep launch: the (element, 256-row block, 32-column chunk) units are split into
//    equal contiguous spans, one per resident warp (stream-K style: no queue, no
//    CTA barrier, balance to one unit); a warp keeps its rows in registers while it
//    walks the chunks of a row block, the next chunk's columns arrive by cp.async;
//    partial (min, tag) keys go to an L2-resident workspace with coalesced stores;
//  * finalize launch (programmatic dependent launch): four lanes per output point
//    reduce its partial keys and re-evaluate the <=32 (rows) / R (columns) tagged
//    candidates with the same arithmetic to emit the exact first argmin.
#include <cooperative_groups.h>

#include <cstdlib>

#include <vector>

#include "pnae_common.cuh"

namespace cg = cooperative_groups;

namespace {

typedef unsigned long long u64;

#ifndef PNAE_NN_ROWS
#define PNAE_NN_ROWS 8
#endif
constexpr int kR = PNAE_NN_ROWS;      // rows per lane (contiguous: lane l owns rows l*kR .. l*kR+kR-1 of the block)
constexpr int kRowsPerBlock = 32 * kR; // 256 rows per warp
constexpr int kChunk = 32;            // columns per unit = columns per row-argmin tag
constexpr int kWarps = 4;             // warps per CTA (independent; no CTA-wide barrier anywhere)
constexpr int kGroup = 4;             // columns whose cross-lane reductions are batched
#ifndef PNAE_NN_CTAS
#define PNAE_NN_CTAS 4
#endif
constexpr int kCtasPerSm = PNAE_NN_CTAS;
#ifndef PNAE_NN_UNROLL
#define PNAE_NN_UNROLL 2
#endif
constexpr int kUnroll = PNAE_NN_UNROLL;   // column groups per trip of the sweep's inner loop
constexpr size_t kWsBudget = 256ull << 20;

struct FwdParams {
    int be;            // elements in this launch
    int n, m;
    int nrb, nch;      // row blocks / column chunks per element
    int nslot;         // row-partial slots per row block in the workspace layout (bound for any launch of the call)
    int nsl;           // slots in use in THIS launch: spans touching one row block <= nsl <= nslot; slots the sweep
                       // does not reach are filled with +inf keys, so the finalize reads nsl slots unconditionally
    long long units;   // be * nrb * nch
    long long warps;   // sweep grid size in warps: the span of warp w is [w*units/warps, (w+1)*units/warps)
    const float *xyz1, *xyz2;
    float *dist1, *dist2;
    int *idx1, *idx2;
    u64 *rowkeys;      // [be][nrb][nslot][256]  (min bits << 32 | chunk index)
    u64 *colkeys;      // [be][nrb][m]           (min bits << 32 | ballot of lanes holding the min)
    // fused loss + gradient (pnae_chamfer_loss_grad); all NULL/0 on the plain forward path
    float *loss;       // [1]      += w1*sum(dist1) + w2*sum(dist2)
    float *gxyz1, *gxyz2;   // (be,n,3), (be,m,3): d loss / d xyz, zeroed by the sweep, accumulated by the finalize
    float w1, w2;
    const float *gd1, *gd2;   // optional per-point upstream gradients (b,n), (b,m): replace w1 / w2 (pnae_nn_distance_fwd_grad)
    int zero_loss;     // this launch is the first chunk of the call: it also zeroes *loss
    int small;         // (units + 1) * warps and the point count fit 32 bits: cheap unsigned index arithmetic
#ifdef PNAE_NN_TRACE
    unsigned long long *trace;   // tools/trace_nn.py: [warp][4] globaltimer stamps (entry, after wait, first data, exit)
#endif
};

// three-input minimum (FMNMX3); NaN operands are ignored like fminf
__device__ __forceinline__ float min3f(float a, float b, float c)
{
    float r;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

// warp that owns unit u under the span formula above
template <typename T>
__device__ __forceinline__ T owner_of(T u, T warps, T units)
{
    return ((u + 1) * warps - 1) / units;
}

// explicit 32-bit shared-memory addressing for the sweep's inner loop (one base register, immediate offsets)
__device__ __forceinline__ unsigned smem_u32(const void *ptr) { return (unsigned)__cvta_generic_to_shared(ptr); }
__device__ __forceinline__ float4 lds128(unsigned addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ u64 lds64(unsigned addr)
{
    u64 v;
    asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128(unsigned addr, unsigned x, unsigned y, unsigned z, unsigned w)
{
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ void sts128_if(int pred, unsigned addr, unsigned x, unsigned y, unsigned z, unsigned w)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %5, 0;\n\t@p st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n\t}"
                 ::"r"(addr), "r"(x), "r"(y), "r"(z), "r"(w), "r"(pred) : "memory");
}

// cp.async (LDGSTS) helpers: 4-byte granularity because points are 12-byte xyz triples
__device__ __forceinline__ void cp_async4(unsigned smem_dst, const void *gmem_src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_dst), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Asynchronous copy of one chunk's 32 columns into float4 slots (w unused) at shared address scol_s.  An
// interior chunk is 96 contiguous floats: one base address, immediate offsets.  Only the last, partial chunk
// of an element clamps its indices (duplicates of the last point never change a minimum).
__device__ __forceinline__ void prefetch_cols(const float *p2, int m, int ch, unsigned scol_s, int lane)
{
    const int col0 = ch * kChunk;
    if (col0 + kChunk <= m) {
        const float *src = p2 + (size_t)col0 * 3 + lane;
#pragma unroll
        for (int i = 0; i < 3; i++) {
            const int f = lane + 32 * i, c = f / 3;
            cp_async4(scol_s + c * 16 + (f - c * 3) * 4, src + 32 * i);
        }
    } else {
#pragma unroll
        for (int i = 0; i < 3; i++) {
            const int f = lane + 32 * i, c = f / 3;
            const int k = min(col0 + c, m - 1);
            cp_async4(scol_s + c * 16 + (f - c * 3) * 4, p2 + (size_t)k * 3 + (f - c * 3));
        }
    }
}

// Same for the 256 rows of row block rb, as flat xyz floats (768 contiguous floats when the block is interior).
__device__ __forceinline__ void prefetch_rows(const float *p1, int n, int rb, unsigned srow_s, int lane)
{
    const int row0 = rb * kRowsPerBlock;
    if (row0 + kRowsPerBlock <= n) {
        const float *src = p1 + (size_t)row0 * 3 + lane;
#pragma unroll
        for (int i = 0; i < kRowsPerBlock * 3 / 32; i++) cp_async4(srow_s + (lane + 32 * i) * 4, src + 32 * i);
    } else {
#pragma unroll
        for (int i = 0; i < kRowsPerBlock * 3 / 32; i++) {
            const int f = lane + 32 * i, r = f / 3;
            const int j = min(row0 + r, n - 1);
            cp_async4(srow_s + f * 4, p1 + (size_t)j * 3 + (f - r * 3));
        }
    }
}

// Sweep launch.  Each warp is an independent worker walking its span of units: an outer loop over the row
// blocks the span touches (rows loaded to registers, partial row keys flushed on leaving), an inner loop over
// the span's chunks inside that row block with no index arithmetic beyond a few running pointers.
struct __align__(16) SweepSmem {                 // per warp
    float4 col[2][kChunk + 1];                   // +1: the loop's look-ahead load of the last group lands there
    float row[kRowsPerBlock * 3];
    u64 key[kChunk];
    float4 snap[kR / 4][32];                     // each lane's row minima as of the previous chunk boundary
    int4 tag[kR / 4][32];                        // chunk in which each row's minimum last strictly decreased
};

__global__ void __launch_bounds__(kWarps * 32, kCtasPerSm)
nn_fwd_kernel(const FwdParams p)
{
    constexpr unsigned kColBuf = (kChunk + 1) * (unsigned)sizeof(float4);
    constexpr unsigned kRowOff = 2 * kColBuf, kKeyOff = kRowOff + kRowsPerBlock * 12;
    constexpr unsigned kSnapOff = kKeyOff + kChunk * 8, kTagOff = kSnapOff + kR * 32 * 4;
    static_assert(kKeyOff % 16 == 0 && sizeof(SweepSmem) == kTagOff + kR * 32 * 4, "SweepSmem layout");
    __shared__ SweepSmem smem_all[kWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // let the finalize launch become resident as SMs drain (it blocks in cudaGridDependencySynchronize
    // until every CTA of this grid has finished and flushed): its launch latency hides under the sweep's tail
    asm volatile("griddepcontrol.launch_dependents;");
#ifdef PNAE_NN_TRACE
    auto stamp = [&](int i) {
        if (p.trace != nullptr && lane == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            p.trace[((size_t)blockIdx.x * kWarps + warp) * 4 + i] = t;
        }
    };
    stamp(0);
    bool traced = false;
#endif
    const unsigned sm_s = smem_u32(&smem_all[warp]);       // the one shared-memory base register of this warp
    const int wid = blockIdx.x * kWarps + warp;
    const int per_e = p.nrb * p.nch;
    int rem;                      // units left in this warp's span
    int e, rb, ch;
    // span [u, uend) of this warp and the (element, row block, chunk) of its first unit; chunk fastest, so a
    // span stays inside one row block as long as possible
    if (p.small) {
        const unsigned w = (unsigned)wid, W = (unsigned)p.warps, U = (unsigned)p.units;
        const unsigned u0 = w * U / W;
        rem = (int)((w + 1) * U / W - u0);
        e = (int)(u0 / (unsigned)per_e);
        const unsigned r = u0 - (unsigned)e * (unsigned)per_e;
        rb = (int)(r / (unsigned)p.nch);
        ch = (int)(r - (unsigned)rb * (unsigned)p.nch);
    } else {
        const long long u = wid * p.units / p.warps;
        rem = (int)((wid + 1) * p.units / p.warps - u);     // <= ceil(units / warps): the workspace budget keeps it far below 2^31
        e = (int)(u / per_e);
        const int r = (int)(u - (long long)e * per_e);
        rb = r / p.nch;
        ch = r - rb * p.nch;
    }
    // this grid is itself launched with programmatic stream serialization: everything above (parameters only)
    // may run while the kernel before it drains; global memory may only be touched from here on
    asm volatile("griddepcontrol.wait;" ::: "memory");
#ifdef PNAE_NN_TRACE
    stamp(1);
#endif
    if (p.gxyz1 != nullptr) {
        // fused loss+gradient: the finalize accumulates into these with atomics, so clear them here (it cannot
        // start touching them before this whole grid has finished)
        const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nt = (long long)gridDim.x * blockDim.x;
        for (long long i = tid; i < (long long)p.be * p.n * 3; i += nt) p.gxyz1[i] = 0.f;
        for (long long i = tid; i < (long long)p.be * p.m * 3; i += nt) p.gxyz2[i] = 0.f;
        if (tid == 0 && p.zero_loss) *p.loss = 0.f;
    }
    if (rem <= 0) return;

    // rows and their running minima live in registers; the per-chunk bookkeeping (snapshot of the minima at the
    // last chunk boundary, tags) lives in shared memory so the inner loop has the whole register file
    float rx[kR], ry[kR], rz[kR], best[kR];
    unsigned buf = 0;             // byte offset of the column buffer in use: 0 or kColBuf
    prefetch_rows(p.xyz1 + (size_t)e * p.n * 3, p.n, rb, sm_s + kRowOff, lane);
    prefetch_cols(p.xyz2 + (size_t)e * p.m * 3, p.m, ch, sm_s, lane);
    cp_async_commit();

    for (;;) {
        // ---- one row block: the chunks [ch0, ch_end) of (e, rb) belong to this warp
        const int ch0 = ch;
        const int ch_end = ch + min(p.nch - ch, rem);
        rem -= ch_end - ch;        // > 0: the span goes on into the next row block (then ch_end == nch)
        const float *p2 = p.xyz2 + (size_t)e * p.m * 3;
        int kcol = ch * kChunk + lane;
        u64 *ck = p.colkeys + ((size_t)e * p.nrb + rb) * p.m + kcol;

        cp_async_wait_all();
        __syncwarp();
#ifdef PNAE_NN_TRACE
        if (!traced) { stamp(2); traced = true; }
#endif
        {
            float tmp[kR * 3];
#pragma unroll
            for (int i = 0; i < kR * 3 / 4; i++) {
                const float4 v = lds128(sm_s + kRowOff + lane * (kR * 12) + i * 16);
                tmp[4 * i] = v.x; tmp[4 * i + 1] = v.y; tmp[4 * i + 2] = v.z; tmp[4 * i + 3] = v.w;
            }
#pragma unroll
            for (int r = 0; r < kR; r++) {
                rx[r] = tmp[3 * r]; ry[r] = tmp[3 * r + 1]; rz[r] = tmp[3 * r + 2];
                best[r] = __int_as_float(0x7f800000);
            }
#pragma unroll
            for (int h = 0; h < kR / 4; h++) {
                sts128(sm_s + kSnapOff + h * 512 + lane * 16, 0x7f800000u, 0x7f800000u, 0x7f800000u, 0x7f800000u);
                sts128(sm_s + kTagOff + h * 512 + lane * 16, 0u, 0u, 0u, 0u);
            }
        }
        __syncwarp();      // every lane has its rows in registers: the row buffer may be refilled

        for (;;) {
            // successor unit's data arrives while this chunk is swept
            if (ch + 1 < ch_end) {
                prefetch_cols(p2, p.m, ch + 1, sm_s + (buf ^ kColBuf), lane);
            } else if (rem > 0) {
                int e2 = e, rb2 = rb + 1;
                if (rb2 == p.nrb) { rb2 = 0; e2++; }
                prefetch_rows(p.xyz1 + (size_t)e2 * p.n * 3, p.n, rb2, sm_s + kRowOff, lane);
                prefetch_cols(p.xyz2 + (size_t)e2 * p.m * 3, p.m, 0, sm_s + (buf ^ kColBuf), lane);
            }
            cp_async_commit();

            // ---- 256 rows x 32 columns
            unsigned ca = sm_s + buf;                      // column records; this group's keys go to ca-relative ka
            unsigned ka = sm_s + kKeyOff;
            const unsigned cend = ca + kChunk * (unsigned)sizeof(float4);
            float4 qn = lds128(ca);
#pragma unroll kUnroll
            do {
                // kGroup columns at a time: their cross-lane reductions (REDUX -> compare -> ballot) are
                // independent chains, issued back to back so their fixed latencies overlap
                unsigned bits[kGroup];
#pragma unroll
                for (int g = 0; g < kGroup; g += 2) {
                    // two columns at a time so the minima can use the three-input FMNMX3
                    const float4 q0 = qn;
                    const float4 q1 = lds128(ca + (g + 1) * (unsigned)sizeof(float4));
                    qn = lds128(ca + (g + 2) * (unsigned)sizeof(float4));   // next pair's first record is in flight during this pair's math
                    float d0[kR], d1[kR];
#pragma unroll
                    for (int r = 0; r < kR; r++) {
                        d0[r] = pnae_sqdist(q0.x - rx[r], q0.y - ry[r], q0.z - rz[r]);
                        d1[r] = pnae_sqdist(q1.x - rx[r], q1.y - ry[r], q1.z - rz[r]);
                        best[r] = min3f(best[r], d0[r], d1[r]);
                    }
                    // column minima over this lane's rows; d >= 0: unsigned order == float order
                    static_assert(kR % 2 == 0 && kR >= 4, "column tree below takes the rows two at a time");
                    float c0 = min3f(d0[0], d0[1], d0[2]), c1 = min3f(d1[0], d1[1], d1[2]);
#pragma unroll
                    for (int r = 3; r + 1 < kR; r += 2) { c0 = min3f(c0, d0[r], d0[r + 1]); c1 = min3f(c1, d1[r], d1[r + 1]); }
                    bits[g] = __float_as_uint(fminf(c0, d0[kR - 1]));
                    bits[g + 1] = __float_as_uint(fminf(c1, d1[kR - 1]));
                }
                unsigned mn[kGroup], who[kGroup];
#pragma unroll
                for (int g = 0; g < kGroup; g++) mn[g] = __reduce_min_sync(0xffffffffu, bits[g]);
#pragma unroll
                for (int g = 0; g < kGroup; g++) who[g] = __ballot_sync(0xffffffffu, bits[g] == mn[g]);
                static_assert(kGroup == 4, "two 16-byte key stores per group");
                sts128_if(lane == 0, ka, who[0], mn[0], who[1], mn[1]);          // u64 key = min bits << 32 | ballot
                sts128_if(lane == 0, ka + 16, who[2], mn[2], who[3], mn[3]);
                ca += kGroup * (unsigned)sizeof(float4);
                ka += kGroup * (unsigned)sizeof(u64);
            } while (ca != cend);
            __syncwarp();
            const u64 mykey = lds64(sm_s + kKeyOff + lane * 8);     // lane c carries the key of column c of this chunk
            __syncwarp();
            // a strict decrease during this chunk => the row's running minimum first appears here
#pragma unroll
            for (int h = 0; h < kR / 4; h++) {
                const float4 sn = lds128(sm_s + kSnapOff + h * 512 + lane * 16);
                const float4 tg = lds128(sm_s + kTagOff + h * 512 + lane * 16);
                const unsigned uch = (unsigned)ch;
                sts128(sm_s + kTagOff + h * 512 + lane * 16,
                       best[4 * h] < sn.x ? uch : __float_as_uint(tg.x), best[4 * h + 1] < sn.y ? uch : __float_as_uint(tg.y),
                       best[4 * h + 2] < sn.z ? uch : __float_as_uint(tg.z), best[4 * h + 3] < sn.w ? uch : __float_as_uint(tg.w));
                sts128(sm_s + kSnapOff + h * 512 + lane * 16, __float_as_uint(best[4 * h]), __float_as_uint(best[4 * h + 1]),
                       __float_as_uint(best[4 * h + 2]), __float_as_uint(best[4 * h + 3]));
            }
            if (kcol < p.m) *ck = mykey;
            ck += kChunk; kcol += kChunk;
            buf ^= kColBuf;
            if (++ch == ch_end) break;
            cp_async_wait_all();
            __syncwarp();
        }

        // ---- leaving the row block: store this warp's partial row keys
        {
            const long long first = ((long long)e * p.nrb + rb) * p.nch;      // first unit of the row block
            auto owner = [&](long long t) {
                return p.small ? (long long)owner_of<unsigned>((unsigned)t, (unsigned)p.warps, (unsigned)p.units)
                               : owner_of<long long>(t, p.warps, p.units);
            };
            // rank of this warp among the warps whose spans touch the row block: consecutive warps when every
            // warp has work (units >= warps), one warp per unit otherwise -- the smaller of the two counts
            const long long own = owner(first);
            const int slot = (int)min((long long)wid - own, (long long)ch0);
            u64 *rk0 = p.rowkeys + (((size_t)e * p.nrb + rb) * p.nslot) * kRowsPerBlock + lane * kR;
            u64 *rk = rk0 + (size_t)slot * kRowsPerBlock;
#pragma unroll
            for (int h = 0; h < kR / 4; h++) {
                const float4 tg = lds128(sm_s + kTagOff + h * 512 + lane * 16);
                rk[4 * h] = ((u64)__float_as_uint(best[4 * h]) << 32) | __float_as_uint(tg.x);
                rk[4 * h + 1] = ((u64)__float_as_uint(best[4 * h + 1]) << 32) | __float_as_uint(tg.y);
                rk[4 * h + 2] = ((u64)__float_as_uint(best[4 * h + 2]) << 32) | __float_as_uint(tg.z);
                rk[4 * h + 3] = ((u64)__float_as_uint(best[4 * h + 3]) << 32) | __float_as_uint(tg.w);
            }
            if (ch0 == 0) {
                // the warp that swept the block's first unit also pads the slots no span reaches
                const int used = (int)min(owner(first + p.nch - 1) - own + 1, (long long)p.nch);
                for (int sl = used; sl < p.nsl; sl++) {
#pragma unroll
                    for (int r = 0; r < kR; r++) rk0[(size_t)sl * kRowsPerBlock + r] = ~0ull;
                }
            }
        }
        if (rem <= 0) break;
        ch = 0;
        if (++rb == p.nrb) { rb = 0; e++; }
    }
#ifdef PNAE_NN_TRACE
    stamp(3);
#endif
}

// Finalize launch: kFinLanes lanes per output point.  Each group reduces the point's partial
// keys, then re-evaluates the tagged candidates (32 columns for a point of xyz1, kR rows for a
// point of xyz2) with the same arithmetic as the sweep; the lowest index whose distance equals
// the minimum is the reference's first argmin.  All loads of a phase are independent, so a
// point costs two dependent L2 round trips.
#ifndef PNAE_NN_FINLANES
#define PNAE_NN_FINLANES 4
#endif
constexpr int kFinLanes = PNAE_NN_FINLANES;
#ifndef PNAE_NN_FINTHREADS
#define PNAE_NN_FINTHREADS 128
#endif
#ifndef PNAE_NN_FINOCC
#define PNAE_NN_FINOCC 8
#endif
constexpr int kFinThreads = PNAE_NN_FINTHREADS;
#ifndef PNAE_NN_FINWAVES
#define PNAE_NN_FINWAVES 1            // finalize grid size in units of "one resident wave"
#endif
constexpr int kFinOcc = PNAE_NN_FINOCC;       // resident CTAs per SM the finalize is compiled for

// FUSED: additionally accumulate loss = sum(g1*dist1) + sum(g2*dist2) and its gradient
//   d/d a_j = 2 g (a_j - c_nn(j)),   d/d c_nn(j) = -2 g (a_j - c_nn(j))        (tf_nndistance_g.cu:142-148)
// with g the constant w1 / w2 (pnae_chamfer_loss_grad) or the caller's grad_dist arrays (pnae_nn_distance_fwd_grad),
// so a Chamfer step needs no separate gradient launch and no dist/idx round trip.  dist/idx outputs and the loss
// value are optional on this path.
// Grid: y = element, x = CTAs sharing that element's n + m points (its points of xyz1, then those of xyz2) in
// blocks of kFinThreads/kFinLanes: no divisions anywhere.  The grid is sized to be resident at once; each lane
// group walks its points with the NEXT point's coordinates and partial keys already in flight while the current
// point's candidates are fetched and compared, so a point costs one exposed L2 round trip instead of two.
struct FinPoint {
    float x, y, z;
    u64 v[4];          // first four partial keys of this lane (slot / row block  sub + t * kFinLanes)
};

__device__ __forceinline__ void fin_issue(const FwdParams &p, int e, int pt, int sub, FinPoint &a)
{
    const int per_e = p.n + p.m;
    const int r = min(pt, per_e - 1);
    const bool row = r < p.n;
    const int j = row ? r : r - p.n;
    const float *src = (row ? p.xyz1 + (size_t)e * p.n * 3 : p.xyz2 + (size_t)e * p.m * 3) + (size_t)j * 3;
    a.x = __ldg(src); a.y = __ldg(src + 1); a.z = __ldg(src + 2);
    const int rb = j / kRowsPerBlock;
    const u64 *base = row ? p.rowkeys + (((size_t)e * p.nrb + rb) * p.nslot + sub) * kRowsPerBlock + (j - rb * kRowsPerBlock)
                          : p.colkeys + ((size_t)e * p.nrb + sub) * p.m + j;
    const size_t stride = (size_t)kFinLanes * (row ? kRowsPerBlock : p.m);
    const int count = row ? p.nsl : p.nrb;
#pragma unroll
    for (int t = 0; t < 4; t++) a.v[t] = (sub + t * kFinLanes < count) ? __ldcg(base + t * stride) : ~0ull;
}

template <bool FUSED>
__global__ void __launch_bounds__(kFinThreads, kFinOcc)
nn_finalize_kernel(const FwdParams p)
{
    float loss_acc = 0.f;
    const int sub = threadIdx.x & (kFinLanes - 1);
    // shuffles stay inside one point's lane group: a warp whose points straddle the xyz1/xyz2 boundary of an
    // element (n not a multiple of 32/kFinLanes) takes both branches below, so a full-warp mask would be divergent
    const unsigned gmask = (unsigned)((1ull << kFinLanes) - 1ull) << ((threadIdx.x & 31) & ~(kFinLanes - 1));
    constexpr int kPtsPerCta = kFinThreads / kFinLanes;
    const int per_e = p.n + p.m;
    const int e = blockIdx.y;
    const float *p1 = p.xyz1 + (size_t)e * p.n * 3;
    const float *p2 = p.xyz2 + (size_t)e * p.m * 3;
    // wide candidate loads need the element bases 16-byte (xyz2) / 8-byte (xyz1) aligned
    const bool vec2 = (reinterpret_cast<size_t>(p.xyz2) & 15) == 0 && (p.m & 3) == 0;
    const bool vec1 = (reinterpret_cast<size_t>(p.xyz1) & 7) == 0 && (p.n & 1) == 0;
    asm volatile("griddepcontrol.wait;" ::: "memory");        // launched with programmatic stream serialization
    asm volatile("griddepcontrol.launch_dependents;");        // the gradient kernel may queue up behind us the same way
    int pt = (int)(blockIdx.x * (unsigned)kPtsPerCta + threadIdx.x / kFinLanes);
    const int step = (int)(gridDim.x * (unsigned)kPtsPerCta);
    const int pt_end = (per_e + kPtsPerCta - 1) / kPtsPerCta * kPtsPerCta;      // warp-uniform trip count
    FinPoint nxt;
    if (pt < pt_end) fin_issue(p, e, pt, sub, nxt);
    for (; pt < pt_end; pt += step) {
        const FinPoint cur = nxt;
        const bool live = pt < per_e;
        const int r = live ? pt : per_e - 1;
        const float x = cur.x, y = cur.y, z = cur.z;
        if (r < p.n) {
            // point j of xyz1 -> dist1 / idx1
            const int j = r;
            u64 key = min(min(cur.v[0], cur.v[1]), min(cur.v[2], cur.v[3]));   // (min bits, chunk): lower distance, then lower chunk
            if (p.nsl > 4 * kFinLanes) {
                const int rb = j / kRowsPerBlock;
                const u64 *rk = p.rowkeys + (((size_t)e * p.nrb + rb) * p.nslot) * kRowsPerBlock + (j - rb * kRowsPerBlock);
                for (int sl = sub + 4 * kFinLanes; sl < p.nsl; sl += kFinLanes) key = min(key, __ldcg(rk + (size_t)sl * kRowsPerBlock));
            }
#pragma unroll
            for (int o = kFinLanes / 2; o > 0; o >>= 1) key = min(key, (u64)__shfl_xor_sync(gmask, key, o));
            const float want = __uint_as_float((unsigned)(key >> 32));
            const int k0 = (int)(unsigned)key * kChunk;
            constexpr int kPer = kChunk / kFinLanes;          // candidates per lane, contiguous
            float cf[kPer * 3];
            if (kPer % 4 == 0 && vec2 && k0 + kChunk <= p.m) {
                const float4 *src = reinterpret_cast<const float4 *>(p2 + (size_t)(k0 + sub * kPer) * 3);
#pragma unroll
                for (int c = 0; c < kPer * 3 / 4; c++) {
                    const float4 v = __ldg(src + c);
                    cf[4 * c] = v.x; cf[4 * c + 1] = v.y; cf[4 * c + 2] = v.z; cf[4 * c + 3] = v.w;
                }
            } else {
#pragma unroll
                for (int c = 0; c < kPer; c++) {
                    const int k = min(k0 + sub * kPer + c, p.m - 1);
                    cf[3 * c] = __ldg(p2 + k * 3); cf[3 * c + 1] = __ldg(p2 + k * 3 + 1); cf[3 * c + 2] = __ldg(p2 + k * 3 + 2);
                }
            }
            if (pt + step < pt_end) fin_issue(p, e, pt + step, sub, nxt);   // in flight while the candidates arrive
            int found = 0x7fffffff;
#pragma unroll
            for (int c = kPer - 1; c >= 0; c--)
                if (pnae_sqdist(cf[3 * c] - x, cf[3 * c + 1] - y, cf[3 * c + 2] - z) == want) found = min(k0 + sub * kPer + c, p.m - 1);
#pragma unroll
            for (int o = kFinLanes / 2; o > 0; o >>= 1) found = min(found, __shfl_xor_sync(gmask, found, o));
            if (live && sub == 0) {
                const int nn = found == 0x7fffffff ? min(k0, p.m - 1) : found;
                if (p.dist1 != nullptr) {
                    p.dist1[(size_t)e * p.n + j] = want;
                    p.idx1[(size_t)e * p.n + j] = nn;
                }
                if (FUSED) {
                    const float w = p.gd1 != nullptr ? __ldg(p.gd1 + (size_t)e * p.n + j) : p.w1;
                    loss_acc = fmaf(w, want, loss_acc);
                    const float g = __fmul_rn(w, 2.0f);
                    float *ga = p.gxyz1 + ((size_t)e * p.n + j) * 3, *gc = p.gxyz2 + ((size_t)e * p.m + nn) * 3;
                    const float vx = __fmul_rn(g, __fsub_rn(x, __ldg(p2 + nn * 3))), vy = __fmul_rn(g, __fsub_rn(y, __ldg(p2 + nn * 3 + 1)));
                    const float vz = __fmul_rn(g, __fsub_rn(z, __ldg(p2 + nn * 3 + 2)));
                    atomicAdd(ga, vx); atomicAdd(ga + 1, vy); atomicAdd(ga + 2, vz);
                    atomicAdd(gc, -vx); atomicAdd(gc + 1, -vy); atomicAdd(gc + 2, -vz);
                }
            }
        } else {
            // point k of xyz2 -> dist2 / idx2
            const int k = r - p.n;
            u64 key = ~0ull;       // (min bits, row block): the lowest row block wins ties
            unsigned who = 1;      // ballot of the lanes that held the minimum in that row block
#pragma unroll
            for (int t = 0; t < 4; t++) {
                const int rb = sub + t * kFinLanes;
                const u64 cand = (cur.v[t] & 0xffffffff00000000ull) | (unsigned)rb;
                if (rb < p.nrb && cand < key) { key = cand; who = (unsigned)cur.v[t]; }
            }
            if (p.nrb > 4 * kFinLanes) {
                const u64 *ck = p.colkeys + (size_t)e * p.nrb * p.m + k;
                for (int rb = sub + 4 * kFinLanes; rb < p.nrb; rb += kFinLanes) {
                    const u64 v = __ldcg(ck + (size_t)rb * p.m);
                    const u64 cand = (v & 0xffffffff00000000ull) | (unsigned)rb;
                    if (cand < key) { key = cand; who = (unsigned)v; }
                }
            }
#pragma unroll
            for (int o = kFinLanes / 2; o > 0; o >>= 1) {
                const u64 k2 = __shfl_xor_sync(gmask, key, o);
                const unsigned w2 = __shfl_xor_sync(gmask, who, o);
                if (k2 < key) { key = k2; who = w2; }
            }
            const int rbw = (int)(unsigned)key;
            const float want = __uint_as_float((unsigned)(key >> 32));
            const int j0 = rbw * kRowsPerBlock + (__ffs(who) - 1) * kR;     // lowest lane holding the min
            constexpr int kPer = kR / kFinLanes;
            float cf[kPer * 3];
            if (kPer % 2 == 0 && vec1 && j0 + kR <= p.n) {
                const float2 *src = reinterpret_cast<const float2 *>(p1 + (size_t)(j0 + sub * kPer) * 3);
#pragma unroll
                for (int c = 0; c < kPer * 3 / 2; c++) {
                    const float2 v = __ldg(src + c);
                    cf[2 * c] = v.x; cf[2 * c + 1] = v.y;
                }
            } else {
#pragma unroll
                for (int c = 0; c < kPer; c++) {
                    const int j = min(j0 + sub * kPer + c, p.n - 1);
                    cf[3 * c] = __ldg(p1 + j * 3); cf[3 * c + 1] = __ldg(p1 + j * 3 + 1); cf[3 * c + 2] = __ldg(p1 + j * 3 + 2);
                }
            }
            if (pt + step < pt_end) fin_issue(p, e, pt + step, sub, nxt);   // in flight while the candidates arrive
            int found = 0x7fffffff;
#pragma unroll
            for (int c = kPer - 1; c >= 0; c--)
                if (pnae_sqdist(x - cf[3 * c], y - cf[3 * c + 1], z - cf[3 * c + 2]) == want) found = min(j0 + sub * kPer + c, p.n - 1);
#pragma unroll
            for (int o = kFinLanes / 2; o > 0; o >>= 1) found = min(found, __shfl_xor_sync(gmask, found, o));
            if (live && sub == 0) {
                const int nn = found == 0x7fffffff ? min(j0, p.n - 1) : found;
                if (p.dist2 != nullptr) {
                    p.dist2[(size_t)e * p.m + k] = want;
                    p.idx2[(size_t)e * p.m + k] = nn;
                }
                if (FUSED) {
                    const float w = p.gd2 != nullptr ? __ldg(p.gd2 + (size_t)e * p.m + k) : p.w2;
                    loss_acc = fmaf(w, want, loss_acc);
                    const float g = __fmul_rn(w, 2.0f);
                    float *ga = p.gxyz2 + ((size_t)e * p.m + k) * 3, *gc = p.gxyz1 + ((size_t)e * p.n + nn) * 3;
                    const float vx = __fmul_rn(g, __fsub_rn(x, __ldg(p1 + nn * 3))), vy = __fmul_rn(g, __fsub_rn(y, __ldg(p1 + nn * 3 + 1)));
                    const float vz = __fmul_rn(g, __fsub_rn(z, __ldg(p1 + nn * 3 + 2)));
                    atomicAdd(ga, vx); atomicAdd(ga + 1, vy); atomicAdd(ga + 2, vz);
                    atomicAdd(gc, -vx); atomicAdd(gc + 1, -vy); atomicAdd(gc + 2, -vz);
                }
            }
        }
    }
    if (FUSED && p.loss != nullptr) {
        __shared__ float s_loss[kFinThreads / 32];
        loss_acc = warp_sum(loss_acc);
        if ((threadIdx.x & 31) == 0) s_loss[threadIdx.x >> 5] = loss_acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            float t = 0.f;
#pragma unroll
            for (int w = 0; w < kFinThreads / 32; w++) t += s_loss[w];
            if (t != 0.f) atomicAdd(p.loss, t);
        }
    }
}

struct FwdPlan {
    int nrb, nch, nslot, be;
    long long warps;
    size_t row_bytes, col_bytes;   // per launch chunk of `be` elements
    size_t total;
};

// ctas: sweep CTAs per SM the grid is sized for (the kernel is compiled for kCtasPerSm)
FwdPlan make_plan(int b, int n, int m, int sms, int ctas = kCtasPerSm)
{
    FwdPlan pl;
    pl.nrb = (n + kRowsPerBlock - 1) / kRowsPerBlock;
    pl.nch = (m + kChunk - 1) / kChunk;
    pl.warps = (long long)sms * ctas * kWarps;
    // a row block's nch units are touched by at most ceil(nch / floor(units/warps)) + 1 spans; bound it
    // independently of `be` (units >= nrb*nch): spans are never shorter than floor(nrb*nch/warps)
    const long long min_span = max(1ll, (long long)pl.nrb * pl.nch / pl.warps);
    pl.nslot = (int)min((long long)pl.nch, (pl.nch + min_span - 1) / min_span + 1);
    const size_t per_e = sizeof(u64) * ((size_t)pl.nrb * pl.nslot * kRowsPerBlock + (size_t)pl.nrb * m);
    const long long be = (long long)(kWsBudget / (per_e ? per_e : 1));
    pl.be = (int)max(1ll, min(min((long long)b, be), 65535ll));   // the finalize grid carries the element in blockIdx.y
    pl.row_bytes = sizeof(u64) * (size_t)pl.be * pl.nrb * pl.nslot * kRowsPerBlock;
    pl.col_bytes = sizeof(u64) * (size_t)pl.be * pl.nrb * m;
    pl.total = pl.row_bytes + pl.col_bytes;
    return pl;
}

// slots one row block can need in a launch of `be` elements: its nch units meet at most
// ceil(nch / shortest span) + 1 spans (every unit its own span when there are fewer units than warps)
int launch_slots(const FwdPlan &pl, int be)
{
    const long long units = (long long)be * pl.nrb * pl.nch;
    const long long span = units / pl.warps;
    return span >= 1 ? (int)min((long long)pl.nslot, (pl.nch + span - 1) / span + 1) : pl.nslot;
}

// ---------------------------------------------------------------------------
// Gradient: one cluster per batch element, one launch.
//   phase 1 (plain stores)  grad_a[j]      = 2 g (a_j - c_idx[j])      for both clouds
//   cluster barrier         (a scatter only ever targets its own element)
//   phase 2 (float atomics) grad_c[idx[j]] -= 2 g (a_j - c_idx[j])
// No memset: phase 1 overwrites every output.  Summation order of the scattered
// half is unspecified, as in the reference (tf_nndistance_g.cu:143-148).
// ---------------------------------------------------------------------------
constexpr int kBwdThreads = 512;
constexpr int kBwdCluster = 8;

__global__ void __launch_bounds__(kBwdThreads)
nn_bwd_kernel(int b, int n, const float *__restrict__ xyz1, int m, const float *__restrict__ xyz2,
              const float *__restrict__ grad_dist1, const int *__restrict__ idx1,
              const float *__restrict__ grad_dist2, const int *__restrict__ idx2,
              float *grad_xyz1, float *grad_xyz2)
{
    cg::cluster_group cluster = cg::this_cluster();
    const int cs = (int)cluster.num_blocks();
    const int nclusters = gridDim.x / cs;
    const int tid = (int)cluster.block_rank() * kBwdThreads + threadIdx.x;
    asm volatile("griddepcontrol.launch_dependents;");        // a following sweep may set itself up while this grid runs
    asm volatile("griddepcontrol.wait;" ::: "memory");        // programmatic dependent launch: idx comes from the kernel before
    const int stride = cs * kBwdThreads;
    constexpr int kKeep = 4;      // points per thread whose scatter term stays in registers across the barrier
    const bool keep = (long long)n + m <= (long long)stride * kKeep;
    for (int e = blockIdx.x / cs; e < b; e += nclusters) {
        const float *p1 = xyz1 + (size_t)e * n * 3, *p2 = xyz2 + (size_t)e * m * 3;
        float *g1 = grad_xyz1 + (size_t)e * n * 3, *g2 = grad_xyz2 + (size_t)e * m * 3;
        // v = 2 g (a_j - c_idx[j]) of point t (t < n: a point of xyz1, else of xyz2) and where its scatter half goes
        auto term = [&](int t, float &vx, float &vy, float &vz, float *&ga, float *&gc) {
            const bool second = t >= n;
            const int j = second ? t - n : t;
            const float *a = (second ? p2 : p1) + j * 3;
            const int j2 = second ? idx2[(size_t)e * m + j] : idx1[(size_t)e * n + j];
            const float *c = (second ? p1 : p2) + j2 * 3;
            const float g = __fmul_rn(second ? grad_dist2[(size_t)e * m + j] : grad_dist1[(size_t)e * n + j], 2.0f);
            vx = __fmul_rn(g, __fsub_rn(__ldg(a), __ldg(c)));
            vy = __fmul_rn(g, __fsub_rn(__ldg(a + 1), __ldg(c + 1)));
            vz = __fmul_rn(g, __fsub_rn(__ldg(a + 2), __ldg(c + 2)));
            ga = (second ? g2 : g1) + j * 3;
            gc = (second ? g1 : g2) + j2 * 3;
        };
        if (keep) {
            // the usual case (n + m <= 4 * cluster threads): every term is evaluated once
            float kx[kKeep], ky[kKeep], kz[kKeep];
            float *kc[kKeep];
#pragma unroll
            for (int i = 0; i < kKeep; i++) {
                const int t = tid + i * stride;
                kc[i] = nullptr;
                if (t < n + m) {
                    float *ga;
                    term(t, kx[i], ky[i], kz[i], ga, kc[i]);
                    ga[0] = kx[i]; ga[1] = ky[i]; ga[2] = kz[i];
                }
            }
            // the barrier orders phase 1's plain stores before phase 2's atomics on the same addresses
            // (release/acquire at cluster scope; all of an element's traffic stays inside its cluster)
            cluster.sync();
#pragma unroll
            for (int i = 0; i < kKeep; i++)
                if (kc[i] != nullptr) { atomicAdd(kc[i], -kx[i]); atomicAdd(kc[i] + 1, -ky[i]); atomicAdd(kc[i] + 2, -kz[i]); }
        } else {
            for (int phase = 0; phase < 2; phase++) {
                for (int t = tid; t < n + m; t += stride) {
                    float vx, vy, vz, *ga, *gc;
                    term(t, vx, vy, vz, ga, gc);
                    if (phase == 0) { ga[0] = vx; ga[1] = vy; ga[2] = vz; }
                    else { atomicAdd(gc, -vx); atomicAdd(gc + 1, -vy); atomicAdd(gc + 2, -vz); }
                }
                if (phase == 0) cluster.sync();
            }
        }
        // the next round works on another element's buffers: no barrier needed between rounds
    }
}

}  // namespace

extern "C" size_t pnae_nn_distance_workspace_bytes(int b, int n, int m)
{
    if (b <= 0 || n <= 0 || m <= 0) return 0;
    return make_plan(b, n, m, pnae_sm_count()).total;
}

extern "C" int pnae_nn_distance_plan(int b, int n, int m, int sm_count, int *plan)
{
    PNAE_REQUIRE(b >= 1 && n >= 1 && m >= 1 && sm_count >= 1 && plan != nullptr, "nn_distance_plan: invalid argument");
    const FwdPlan pl = make_plan(b, n, m, sm_count);
    plan[0] = pl.nrb; plan[1] = pl.nch; plan[2] = pl.nslot; plan[3] = pl.be;
    plan[4] = (int)pl.warps;
    plan[5] = launch_slots(pl, pl.be);                                     // full launches
    plan[6] = launch_slots(pl, b % pl.be ? b % pl.be : pl.be);             // the last, possibly partial one
    plan[7] = kRowsPerBlock; plan[8] = kChunk;
    return PNAE_OK;
}

namespace {
// shared launcher: plain forward (loss == NULL) or fused loss + gradient
int launch_fwd(const char *op, int b, int n, const float *xyz1, int m, const float *xyz2,
               float *dist1, int *idx1, float *dist2, int *idx2,
               float *loss, float *gxyz1, float *gxyz2, float w1, float w2, const float *gd1, const float *gd2,
               void *workspace, size_t workspace_bytes, void *stream,
               int ctas = kCtasPerSm, cudaStream_t fin_stream = nullptr, cudaEvent_t swept = nullptr)
{
    // fin_stream: the finalize goes to that stream, behind the event `swept` recorded after the sweep (the pipelined graph:
    // the next step's sweep then follows this one on `stream` without waiting for this step's finalize)
    const int sms = pnae_sm_count();
    const FwdPlan pl = make_plan(b, n, m, sms, ctas);
    PNAE_REQUIRE(fin_stream == nullptr || pl.be >= b, "%s: the pipelined form needs the whole batch in one launch", op);
    if (workspace == nullptr || workspace_bytes < pl.total) {
        pnae_set_error("%s: workspace too small (%zu < %zu bytes)", op, workspace_bytes, pl.total);
        return PNAE_ERR_WORKSPACE;
    }
    PNAE_REQUIRE(pnae_aligned(workspace, 8), "%s: workspace must be 8-byte aligned", op);
    PNAE_REQUIRE((long long)n + m < (1ll << 31) - 4096, "%s: n + m must stay below 2^31 - 4096 (got %d + %d)", op, n, m);
    cudaStream_t st = (cudaStream_t)stream;
    char *ws = (char *)workspace;
    for (int e0 = 0; e0 < b; e0 += pl.be) {
        FwdParams p;
        p.be = min(pl.be, b - e0);
        p.n = n; p.m = m; p.nrb = pl.nrb; p.nch = pl.nch; p.nslot = pl.nslot;
        p.units = (long long)p.be * pl.nrb * pl.nch;
        p.warps = pl.warps;
        p.xyz1 = xyz1 + (size_t)e0 * n * 3; p.xyz2 = xyz2 + (size_t)e0 * m * 3;
        p.dist1 = dist1 ? dist1 + (size_t)e0 * n : nullptr; p.idx1 = idx1 ? idx1 + (size_t)e0 * n : nullptr;
        p.dist2 = dist2 ? dist2 + (size_t)e0 * m : nullptr; p.idx2 = idx2 ? idx2 + (size_t)e0 * m : nullptr;
        p.rowkeys = (u64 *)ws;
        p.colkeys = (u64 *)(ws + pl.row_bytes);
        p.loss = loss;
        p.gxyz1 = gxyz1 ? gxyz1 + (size_t)e0 * n * 3 : nullptr;
        p.gxyz2 = gxyz2 ? gxyz2 + (size_t)e0 * m * 3 : nullptr;
        p.w1 = w1; p.w2 = w2; p.zero_loss = (e0 == 0 && loss != nullptr);
        p.gd1 = gd1 ? gd1 + (size_t)e0 * n : nullptr;
        p.gd2 = gd2 ? gd2 + (size_t)e0 * m : nullptr;
        // 32-bit index arithmetic whenever the point count and the span formula's (u+1)*warps fit
        const bool force64 = getenv("PNAE_NN_INDEX64") != nullptr;            // test hook for the wide path, read on every call
        const long long groups = (long long)p.be * ((long long)n + m);
        const bool small = !force64 && groups < (1ll << 31) && (p.units + 1) * p.warps < (1ll << 32);
        p.small = small;
#ifdef PNAE_NN_TRACE
        { const char *tp = getenv("PNAE_NN_TRACE_PTR"); p.trace = tp ? (unsigned long long *)strtoull(tp, nullptr, 16) : nullptr; }
#endif
        p.nsl = launch_slots(pl, p.be);
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // launch latency overlaps the previous kernel's tail
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(pl.warps / kWarps));
        cfg.blockDim = dim3(kWarps * 32);
        cfg.stream = st;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        PNAE_CUDA_OK(cudaLaunchKernelEx(&cfg, nn_fwd_kernel, p));
        // resident at once: sms * kFinOcc CTAs shared out over the elements (at least one, at most all its blocks)
        const int fin_blocks = (n + m + kFinThreads / kFinLanes - 1) / (kFinThreads / kFinLanes);
        const int fin_x = max(1, min(fin_blocks, (sms * kFinOcc * PNAE_NN_FINWAVES + p.be - 1) / p.be));
        cfg.gridDim = dim3((unsigned)fin_x, (unsigned)p.be);
        cfg.blockDim = dim3(kFinThreads);
        if (fin_stream != nullptr) {
            PNAE_CUDA_OK(cudaEventRecord(swept, st));
            PNAE_CUDA_OK(cudaStreamWaitEvent(fin_stream, swept, 0));
            cfg.stream = fin_stream;
            cfg.numAttrs = 0;             // a full dependency on the sweep (the event), no programmatic edge to the previous finalize
        }
        if (gxyz1 != nullptr) PNAE_CUDA_OK(cudaLaunchKernelEx(&cfg, nn_finalize_kernel<true>, p));
        else PNAE_CUDA_OK(cudaLaunchKernelEx(&cfg, nn_finalize_kernel<false>, p));
    }
    return PNAE_OK;
}
}  // namespace

extern "C" int pnae_nn_distance_fwd(int b, int n, const float *xyz1, int m, const float *xyz2,
                                    float *dist1, int *idx1, float *dist2, int *idx2,
                                    void *workspace, size_t workspace_bytes, void *stream)
{
    PNAE_REQUIRE(b >= 0 && n >= 1 && m >= 1, "nn_distance: need b>=0, n>=1, m>=1 (got b=%d n=%d m=%d)", b, n, m);
    PNAE_REQUIRE(xyz1 && xyz2 && dist1 && idx1 && dist2 && idx2, "nn_distance: NULL pointer");
    if (b == 0) return PNAE_OK;
    return launch_fwd("nn_distance", b, n, xyz1, m, xyz2, dist1, idx1, dist2, idx2, nullptr, nullptr, nullptr, 0.f, 0.f,
                      nullptr, nullptr, workspace, workspace_bytes, stream);
}

extern "C" int pnae_chamfer_loss_grad(int b, int n, const float *xyz1, int m, const float *xyz2, float w1, float w2,
                                      float *loss, float *grad_xyz1, float *grad_xyz2,
                                      float *dist1, int *idx1, float *dist2, int *idx2,
                                      void *workspace, size_t workspace_bytes, void *stream)
{
    PNAE_REQUIRE(b >= 1 && n >= 1 && m >= 1, "chamfer_loss_grad: need b>=1, n>=1, m>=1 (got b=%d n=%d m=%d)", b, n, m);
    PNAE_REQUIRE(xyz1 && xyz2 && loss && grad_xyz1 && grad_xyz2, "chamfer_loss_grad: NULL pointer");
    PNAE_REQUIRE((dist1 != nullptr) == (idx1 != nullptr) && (dist2 != nullptr) == (idx2 != nullptr),
                 "chamfer_loss_grad: dist/idx outputs go in pairs");
    return launch_fwd("chamfer_loss_grad", b, n, xyz1, m, xyz2, dist1, idx1, dist2, idx2, loss, grad_xyz1, grad_xyz2, w1, w2,
                      nullptr, nullptr, workspace, workspace_bytes, stream);
}

extern "C" int pnae_nn_distance_fwd_grad(int b, int n, const float *xyz1, int m, const float *xyz2,
                                         const float *grad_dist1, const float *grad_dist2,
                                         float *dist1, int *idx1, float *dist2, int *idx2,
                                         float *grad_xyz1, float *grad_xyz2,
                                         void *workspace, size_t workspace_bytes, void *stream)
{
    PNAE_REQUIRE(b >= 0 && n >= 1 && m >= 1, "nn_distance_fwd_grad: need b>=0, n>=1, m>=1 (got b=%d n=%d m=%d)", b, n, m);
    PNAE_REQUIRE(xyz1 && xyz2 && grad_dist1 && grad_dist2 && dist1 && idx1 && dist2 && idx2 && grad_xyz1 && grad_xyz2,
                 "nn_distance_fwd_grad: NULL pointer");
    if (b == 0) return PNAE_OK;
    return launch_fwd("nn_distance_fwd_grad", b, n, xyz1, m, xyz2, dist1, idx1, dist2, idx2, nullptr, grad_xyz1, grad_xyz2, 0.f, 0.f,
                      grad_dist1, grad_dist2, workspace, workspace_bytes, stream);
}

extern "C" int pnae_nn_distance_bwd(int b, int n, const float *xyz1, int m, const float *xyz2,
                                    const float *grad_dist1, const int *idx1,
                                    const float *grad_dist2, const int *idx2,
                                    float *grad_xyz1, float *grad_xyz2, void *stream)
{
    PNAE_REQUIRE(b >= 0 && n >= 1 && m >= 1, "nn_distance_grad: need b>=0, n>=1, m>=1 (got b=%d n=%d m=%d)", b, n, m);
    PNAE_REQUIRE(xyz1 && xyz2 && grad_dist1 && idx1 && grad_dist2 && idx2 && grad_xyz1 && grad_xyz2,
                 "nn_distance_grad: NULL pointer");
    if (b == 0) return PNAE_OK;
    int cs = kBwdCluster;
    while (cs > 1 && (long long)(cs / 2) * kBwdThreads >= (long long)n + m) cs >>= 1;   // small clouds: smaller clusters
    const int nclusters = (int)min((long long)b, (long long)max(1, pnae_sm_count() * 2 / cs));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(nclusters * cs));
    cfg.blockDim = dim3(kBwdThreads);
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    PNAE_CUDA_OK(cudaLaunchKernelEx(&cfg, nn_bwd_kernel, b, n, xyz1, m, xyz2, grad_dist1, idx1, grad_dist2, idx2,
                                    grad_xyz1, grad_xyz2));
    return PNAE_OK;
}

// ---------------------------------------------------------------------------
// CUDA-graph form of one Chamfer step (forward + gradient): the three kernels are captured once
// over fixed buffers and replayed with a single launch.  At B=32, N=M=2048 the step is ~56 us of
// GPU time, less than launching it kernel by kernel costs on the host.
// ---------------------------------------------------------------------------
struct PnaeGraph {
    cudaGraph_t graph;
    cudaGraphExec_t exec;
};

static int graph_create(bool fused, int steps, int b, int n, const float *const *xyz1, int m, const float *const *xyz2,
                        float *dist1, int *idx1, float *dist2, int *idx2,
                        const float *grad_dist1, const float *grad_dist2,
                        float *grad_xyz1, float *grad_xyz2,
                        void *workspace, size_t workspace_bytes, void **handle)
{
    PNAE_REQUIRE(handle != nullptr, "chamfer_graph_create: NULL handle");
    PNAE_REQUIRE(steps >= 1 && xyz1 && xyz2, "chamfer_graph_create: need steps >= 1 and input pointer lists");
    *handle = nullptr;
    cudaStream_t st;
    PNAE_CUDA_OK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    cudaGraph_t graph = nullptr;
    int rc = PNAE_OK;
    cudaError_t ce = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
    if (ce != cudaSuccess) {
        cudaStreamDestroy(st);
        pnae_set_error("cudaStreamBeginCapture failed: %s", cudaGetErrorString(ce));
        return PNAE_ERR_CUDA;
    }
    for (int s = 0; s < steps && rc == PNAE_OK; s++) {
        if (fused) {
            rc = pnae_nn_distance_fwd_grad(b, n, xyz1[s], m, xyz2[s], grad_dist1, grad_dist2, dist1, idx1, dist2, idx2,
                                           grad_xyz1, grad_xyz2, workspace, workspace_bytes, st);
            continue;
        }
        rc = pnae_nn_distance_fwd(b, n, xyz1[s], m, xyz2[s], dist1, idx1, dist2, idx2, workspace, workspace_bytes, st);
        if (rc == PNAE_OK && grad_xyz1 != nullptr && grad_xyz2 != nullptr)     // NULL gradients: forward only
            rc = pnae_nn_distance_bwd(b, n, xyz1[s], m, xyz2[s], grad_dist1, idx1, grad_dist2, idx2, grad_xyz1, grad_xyz2, st);
    }
    ce = cudaStreamEndCapture(st, &graph);
    cudaStreamDestroy(st);
    if (rc != PNAE_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (ce != cudaSuccess || graph == nullptr) {
        pnae_set_error("cudaStreamEndCapture failed: %s", cudaGetErrorString(ce));
        return PNAE_ERR_CUDA;
    }
    cudaGraphExec_t exec = nullptr;
    ce = cudaGraphInstantiate(&exec, graph, 0);
    if (ce != cudaSuccess) {
        cudaGraphDestroy(graph);
        pnae_set_error("cudaGraphInstantiate failed: %s", cudaGetErrorString(ce));
        return PNAE_ERR_CUDA;
    }
    PnaeGraph *g = new PnaeGraph{graph, exec};
    *handle = g;
    return PNAE_OK;
}

extern "C" int pnae_chamfer_graph_create_multi(int steps, int b, int n, const float *const *xyz1, int m, const float *const *xyz2,
                                               float *dist1, int *idx1, float *dist2, int *idx2,
                                               const float *grad_dist1, const float *grad_dist2,
                                               float *grad_xyz1, float *grad_xyz2,
                                               void *workspace, size_t workspace_bytes, void **handle)
{
    return graph_create(false, steps, b, n, xyz1, m, xyz2, dist1, idx1, dist2, idx2, grad_dist1, grad_dist2, grad_xyz1, grad_xyz2,
                        workspace, workspace_bytes, handle);
}

extern "C" int pnae_chamfer_graph_create_fused_multi(int steps, int b, int n, const float *const *xyz1, int m, const float *const *xyz2,
                                                     float *dist1, int *idx1, float *dist2, int *idx2,
                                                     const float *grad_dist1, const float *grad_dist2,
                                                     float *grad_xyz1, float *grad_xyz2,
                                                     void *workspace, size_t workspace_bytes, void **handle)
{
    PNAE_REQUIRE(grad_xyz1 && grad_xyz2, "chamfer_graph_create_fused: gradient outputs are required");
    return graph_create(true, steps, b, n, xyz1, m, xyz2, dist1, idx1, dist2, idx2, grad_dist1, grad_dist2, grad_xyz1, grad_xyz2,
                        workspace, workspace_bytes, handle);
}

extern "C" int pnae_chamfer_graph_create(int b, int n, const float *xyz1, int m, const float *xyz2,
                                         float *dist1, int *idx1, float *dist2, int *idx2,
                                         const float *grad_dist1, const float *grad_dist2,
                                         float *grad_xyz1, float *grad_xyz2,
                                         void *workspace, size_t workspace_bytes, void **handle)
{
    return pnae_chamfer_graph_create_multi(1, b, n, &xyz1, m, &xyz2, dist1, idx1, dist2, idx2, grad_dist1, grad_dist2,
                                           grad_xyz1, grad_xyz2, workspace, workspace_bytes, handle);
}

// Software-pipelined form of the multi-step graph: step s+1's sweep follows step s's sweep directly and runs WHILE step
// s's finalize (and gradient) resolve on a second captured stream.  The sweep is compute-bound and the finalize
// latency-bound (two dependent L2 round trips per point), so together they cost little more than the sweep alone.  Steps
// cycle through `nsets` >= 2 output sets and `nws` >= 2 workspaces (nws <= nsets); step s+nws's sweep waits for step s's
// finalize (it reuses that workspace and, with nsets == nws, zeroes those gradients).  With two workspaces a sweep still
// waits for the finalize of the step before the previous one, which cannot get SM slots before the previous sweep retires:
// 48.2 us per fused step at B=32, N=M=2048; with three the sweeps follow each other without a gap.
extern "C" int pnae_chamfer_graph_create_pipelined(int fused, int steps, int nsets, int nws, int b, int n, const float *const *xyz1, int m,
                                                   const float *const *xyz2, float *const *dist1, int *const *idx1,
                                                   float *const *dist2, int *const *idx2,
                                                   const float *grad_dist1, const float *grad_dist2,
                                                   float *const *grad_xyz1, float *const *grad_xyz2,
                                                   void *const *workspace, size_t workspace_bytes, void **handle)
{
    PNAE_REQUIRE(handle != nullptr, "chamfer_graph_create_pipelined: NULL handle");
    *handle = nullptr;
    PNAE_REQUIRE(steps >= 1 && nsets >= 2 && nws >= 2 && nws <= nsets && b >= 1 && n >= 1 && m >= 1 && xyz1 && xyz2 && dist1 && idx1 && dist2 && idx2 && workspace,
                 "chamfer_graph_create_pipelined: invalid argument (at least two output sets and two workspaces, no more workspaces than sets)");
    const bool grads = grad_xyz1 != nullptr && grad_xyz2 != nullptr;
    PNAE_REQUIRE(!fused || grads, "chamfer_graph_create_pipelined: the fused form needs gradient outputs");
    PNAE_REQUIRE(!grads || (grad_dist1 && grad_dist2), "chamfer_graph_create_pipelined: upstream gradients are required with gradient outputs");
    for (int k = 0; k < nsets; k++)
        PNAE_REQUIRE(dist1[k] && idx1[k] && dist2[k] && idx2[k] && (!grads || (grad_xyz1[k] && grad_xyz2[k])),
                     "chamfer_graph_create_pipelined: every output set must be complete");
    for (int k = 0; k < nws; k++)
        PNAE_REQUIRE(workspace[k] != nullptr && (k == 0 || workspace[k] != workspace[0]), "chamfer_graph_create_pipelined: %d distinct workspaces are required", nws);
    cudaStream_t st, fin;
    PNAE_CUDA_OK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    if (cudaStreamCreateWithFlags(&fin, cudaStreamNonBlocking) != cudaSuccess) { cudaStreamDestroy(st); pnae_set_error("cudaStreamCreate failed"); return PNAE_ERR_CUDA; }
    std::vector<cudaEvent_t> ev(2 * (size_t)steps, nullptr);
    int rc = PNAE_OK;
    for (auto &e : ev)
        if (rc == PNAE_OK && cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { pnae_set_error("cudaEventCreate failed"); rc = PNAE_ERR_CUDA; }
    cudaGraph_t graph = nullptr;
    cudaError_t ce = cudaSuccess;
    if (rc == PNAE_OK) {
        ce = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
        if (ce != cudaSuccess) { pnae_set_error("cudaStreamBeginCapture failed: %s", cudaGetErrorString(ce)); rc = PNAE_ERR_CUDA; }
    }
    if (rc == PNAE_OK) {
        // Sweep CTAs per SM in this form: two.  The sweep's inner loop is bound by the FP32 pipe, which eight warps per SM
        // already fill, and half the register file stays free for the previous steps' finalize CTAs to be resident
        // beside it.  Measured per fused step (B=32, N=M=2048; 52.8 us sequential) with three workspaces: two CTAs 46.4 us
        // in every window of every run length; four 48.4 us (the finalize only finds slots as sweep CTAs retire).  With
        // two workspaces two CTAs measured anything from 46.6 to 62 us from window to window, four a steady 48.2 us.
#ifndef PNAE_NN_PIPE_CTAS
#define PNAE_NN_PIPE_CTAS 2
#endif
        const int ctas = PNAE_NN_PIPE_CTAS;
        for (int s = 0; s < steps && rc == PNAE_OK; s++) {
            const int k = s % nsets, wk = s % nws;    // output set, workspace
            cudaEvent_t swept = ev[2 * s], done = ev[2 * s + 1];
            // this sweep reuses the workspace of step s - nws (nws <= nsets: and no output set younger than that)
            if (s >= nws && cudaStreamWaitEvent(st, ev[2 * (s - nws) + 1], 0) != cudaSuccess) { pnae_set_error("cudaStreamWaitEvent failed"); rc = PNAE_ERR_CUDA; break; }
            if (fused) {
                rc = launch_fwd("chamfer_graph_create_pipelined", b, n, xyz1[s], m, xyz2[s], dist1[k], idx1[k], dist2[k], idx2[k], nullptr,
                                grad_xyz1[k], grad_xyz2[k], 0.f, 0.f, grad_dist1, grad_dist2, workspace[wk], workspace_bytes, st, ctas, fin, swept);
            } else {
                rc = launch_fwd("chamfer_graph_create_pipelined", b, n, xyz1[s], m, xyz2[s], dist1[k], idx1[k], dist2[k], idx2[k], nullptr,
                                nullptr, nullptr, 0.f, 0.f, nullptr, nullptr, workspace[wk], workspace_bytes, st, ctas, fin, swept);
                if (rc == PNAE_OK && grads)
                    rc = pnae_nn_distance_bwd(b, n, xyz1[s], m, xyz2[s], grad_dist1, idx1[k], grad_dist2, idx2[k], grad_xyz1[k], grad_xyz2[k], fin);
            }
            if (rc == PNAE_OK && cudaEventRecord(done, fin) != cudaSuccess) { pnae_set_error("cudaEventRecord failed"); rc = PNAE_ERR_CUDA; }
        }
        // join: the origin stream waits for the last finalize (the finalize stream is ordered, so for all of them)
        if (rc == PNAE_OK && cudaStreamWaitEvent(st, ev[2 * (steps - 1) + 1], 0) != cudaSuccess) { pnae_set_error("cudaStreamWaitEvent failed"); rc = PNAE_ERR_CUDA; }
        ce = cudaStreamEndCapture(st, &graph);
    }
    for (auto &e : ev) if (e) cudaEventDestroy(e);
    cudaStreamDestroy(fin);
    cudaStreamDestroy(st);
    if (rc != PNAE_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (ce != cudaSuccess || graph == nullptr) {
        pnae_set_error("cudaStreamEndCapture failed: %s", cudaGetErrorString(ce));
        return PNAE_ERR_CUDA;
    }
    cudaGraphExec_t exec = nullptr;
    ce = cudaGraphInstantiate(&exec, graph, 0);
    if (ce != cudaSuccess) {
        cudaGraphDestroy(graph);
        pnae_set_error("cudaGraphInstantiate failed: %s", cudaGetErrorString(ce));
        return PNAE_ERR_CUDA;
    }
    *handle = new PnaeGraph{graph, exec};
    return PNAE_OK;
}

extern "C" int pnae_graph_launch(void *handle, void *stream)
{
    PNAE_REQUIRE(handle != nullptr, "graph_launch: NULL handle");
    PNAE_CUDA_OK(cudaGraphLaunch(static_cast<PnaeGraph *>(handle)->exec, (cudaStream_t)stream));
    return PNAE_OK;
}

extern "C" int pnae_graph_destroy(void *handle)
{
    if (handle == nullptr) return PNAE_OK;
    PnaeGraph *g = static_cast<PnaeGraph *>(handle);
    cudaGraphExecDestroy(g->exec);
    cudaGraphDestroy(g->graph);
    delete g;
    return PNAE_OK;
}
