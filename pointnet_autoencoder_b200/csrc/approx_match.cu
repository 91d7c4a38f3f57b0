// approx_match.cu -- approximate EMD soft assignment for sm_100a.
//
// Replaces approxmatch / approxmatchLauncher (reference:
// tf_ops/approxmatch/tf_approxmatch_g.cu:1-182).  Same algorithm: 10 levels j=7..-2
// (level=-4^j, 0 at the end), per level  A: ratioL = remainL / (1e-9 + sum_l E remainR),
// B: sumr = remainR * sum_k E ratioL -> ratioR, remainR,  C: remainL -= sum_l E ratioL ratioR
// with E = exp(level * |x1_k - x2_l|^2), fp32 throughout, the reference's FMA contractions.
//
// B200 design (DESIGN.md "approx_match"):
//  * ONE persistent cooperative launch; every sweep of every batch element is spread over
//    all SMs (stream-K style contiguous spans of (element, own-block, streamed-chunk) tasks),
//    with a grid barrier between dependent sweeps.  The reference runs one CTA per element.
//  * sweep C of level j and sweep A of level j+1 own the same points and stream the same
//    points, so they share one pass over the pairs (one distance, two exponentials);
//    the last level (level 0 => E == 1) degenerates to O(n+m) sums and is done in closed
//    form; sweep C of the last level only feeds the dense `match`, which is never formed.
//    27 exponentials per pair instead of 30.
//  * the per-point epilogues (the divisions / clamps that produce ratioL, ratioR, remainL,
//    remainR) are evaluated on the fly while a task stages its streamed points, from the
//    previous sweep's partial sums -- no extra pass, no extra barrier.  Staging is the job of
//    dedicated producer warps working one task ahead into a double-buffered tile (named
//    barriers): its dependent L2 round trips never stall the warps that feed the SFU.
//  * the inner loop is packed FP32 (FFMA2/FADD2/FMUL2: two own points per thread share every
//    instruction) feeding MUFU.EX2; on B200 the packed form lets the SFU run concurrently
//    with the FMA pipe (profiles/r1_microbench_b200.txt: 9.5 vs 14.1 cycles per pair).
//  * the per-level factors ratioL_j / ratioR_j are the output; the dense (b,m,n) tensor is
//    only materialised (match_cost.cu) when the caller asks for it.
#include <cooperative_groups.h>

#include "pnae_common.cuh"

namespace cg = cooperative_groups;

namespace {

#ifndef PNAE_AM_THREADS
#define PNAE_AM_THREADS 512
#endif
#ifndef PNAE_AM_TS
#define PNAE_AM_TS 128
#endif
#ifndef PNAE_AM_CTAS
#define PNAE_AM_CTAS 1
#endif
// One CTA per SM: 16 compute warps that advance together (a barrier pair per task) plus 4 producer warps.  With two
// CTAs of 8 warps per SM -- round 1's shape -- the two did not advance together: the hardware favoured one, which finished
// a sweep after 39 us and then idled at the grid barrier while the other, alone and latency-bound, took until 56 us
// (tools/am_trace.py).  Measured at B=32, N=M=2048 (ms): 2 x 256 threads 1.253; 512 + 64 producers 1.193; 512 + 128: 1.145,
// with the pair loop unrolled 8x 1.135; 512 + 256: 1.151; streamed chunks of 64 / 256 points 1.213 / 1.283.
#ifndef PNAE_AM_UNROLL
#define PNAE_AM_UNROLL 8
#endif
#ifndef PNAE_AM_PRODUCERS
#define PNAE_AM_PRODUCERS 128
#endif
constexpr int kInnerUnroll = PNAE_AM_UNROLL;
constexpr int kThreads = PNAE_AM_THREADS;     // compute threads per CTA
constexpr int kProd = PNAE_AM_PRODUCERS;      // producer threads per CTA: they stage the streamed points of the NEXT task
constexpr int kCtaThreads = kThreads + kProd;
constexpr int kOwn = 2 * kThreads;     // own points per task: one packed pair per thread
constexpr int kTs = PNAE_AM_TS;        // streamed points per task
constexpr int kCtasPerSm = PNAE_AM_CTAS;
constexpr int kLevels = PNAE_NUM_LEVELS;
// OFF: measured on B200 it saves 7 % (approx_match) / 13 % (match_cost), but the SFU's 2^-22 relative error becomes
// 2^-20 on the derived exponentials and the ill-conditioned schedule amplifies that past the parity bar (match_cost
// up to 3.4e-5 relative against the oracle, gradients 3x farther from the fp64 truth than the reference kernels).
#ifndef PNAE_EMD_POW4
#define PNAE_EMD_POW4 0
#endif
constexpr bool kPow4 = PNAE_EMD_POW4 != 0;   // derive E_j from E_{j+1} by two squarings where both are needed in one pass

struct EmdParams {
    int b, n, m;
    const float *xyz1, *xyz2;
    float *factors;        // (b, 10, n+m)
    float *remL, *remR;    // [2][b][n], [2][b][m]   remaining mass, double-buffered by level parity
    float *ps0, *ps1;      // [2][b][nslot][maxnm]   partial sums, double-buffered by stage parity
    int nslot, maxnm;
    float multiL, multiR;
};

// geometry of one sweep: `own` points are owned (accumulated) by threads, `str` points stream past
struct Geo {
    int nown, nstr;        // points per element on either side
    int nob, nch;          // own blocks / streamed chunks per element
    long long tasks;       // b * nob * nch, chunk fastest
};

__device__ __forceinline__ Geo make_geo(int b, int nown, int nstr)
{
    Geo g;
    g.nown = nown; g.nstr = nstr;
    g.nob = (nown + kOwn - 1) / kOwn;
    g.nch = (nstr + kTs - 1) / kTs;
    g.tasks = (long long)b * g.nob * g.nch;
    return g;
}

// CTA that owns task t when `tasks` tasks are split into contiguous spans over `ctas` CTAs
__device__ __forceinline__ long long owner_of(long long t, long long ctas, long long tasks)
{
    return ((t + 1) * ctas - 1) / tasks;
}

// number of partial-sum slots the sweep with geometry g left for own block (e, ob)
__device__ __forceinline__ int slots_of(const Geo &g, int e, int ob, long long ctas)
{
    const long long first = ((long long)e * g.nob + ob) * g.nch;
    return (int)min(owner_of(first + g.nch - 1, ctas, g.tasks) - owner_of(first, ctas, g.tasks) + 1, (long long)g.nch);
}

// sum, in slot order, of the partial sums a previous sweep left for own point `i` of element e
__device__ __forceinline__ float gather(const float *ps, const EmdParams &p, const Geo &g, int e, int i, long long ctas,
                                        float init)
{
    const int ns = slots_of(g, e, i / kOwn, ctas);
    const float *q = ps + ((size_t)e * p.nslot) * p.maxnm + i;
    float s = init;
    for (int t = 0; t < ns; t++) s = __fadd_rn(s, __ldcg(q + (size_t)t * p.maxnm));
    return s;
}

enum Kind { kA0 = 0, kB = 1, kCA = 2, kC = 3 };

struct __align__(16) Rec {   // one streamed point, duplicated for the packed (two-own-points) math
    float x0, x1, y0, y1;    // (x,x,y,y)
    float z0, z1, w0, w1;    // (z,z,w,w)   w = the sweep's first weight
};

__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }

// One sweep over all elements.  lev = level index (0..9).
//   kA0 : own k, stream l, w = remainR_0 = multiR                     -> ps0 = suml_0 (without the 1e-9)
//   kB  : own l, stream k, w = ratioL_lev[k]                          -> ps0 = sumr (before * remainR)
//   kCA : own k, stream l, w = ratioR_lev[l], v = remainR_{lev+1}[l]  -> ps0 = sumc_lev, ps1 = suml_{lev+1}
//   kC  : as kCA without the A half (last real level)                 -> ps0 = sumc_lev
// named barriers of the producer / consumer hand-off (0 is __syncthreads): one FULL and one EMPTY per tile buffer
constexpr int kBarFull = 1, kBarEmpty = 3, kBarCompute = 5;
__device__ __forceinline__ void bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

template <int KIND>
__device__ void sweep(const EmdParams &p, int lev, int stage, Rec (*tiles)[kTs], float2 (*tilevs)[kTs])
{
    const long long ctas = gridDim.x;
    const bool rowside = (KIND != kB);
    const Geo g = make_geo(p.b, rowside ? p.n : p.m, rowside ? p.m : p.n);
    const Geo gprev = make_geo(p.b, rowside ? p.m : p.n, rowside ? p.n : p.m);   // geometry of the previous sweep
    const float *own_xyz = rowside ? p.xyz1 : p.xyz2;
    const float *str_xyz = rowside ? p.xyz2 : p.xyz1;
    const int par = stage & 1;
    float *out0 = p.ps0 + (size_t)par * p.b * p.nslot * p.maxnm;
    float *out1 = p.ps1 + (size_t)par * p.b * p.nslot * p.maxnm;
    const float *in0 = p.ps0 + (size_t)(par ^ 1) * p.b * p.nslot * p.maxnm;
    const float *in1 = p.ps1 + (size_t)(par ^ 1) * p.b * p.nslot * p.maxnm;
    const float c0 = pnae_level_scale(lev), c1 = pnae_level_scale(lev + 1);
    const float2 sc0 = f2(c0, c0), sc1 = f2(c1, c1);

    const long long t0 = (long long)blockIdx.x * g.tasks / ctas;
    const long long tend = ((long long)blockIdx.x + 1) * g.tasks / ctas;
    const int per_e = g.nob * g.nch;

    if (threadIdx.x >= kThreads) {
        // ===== producers: stage task t's streamed points (and evaluate the previous sweep's epilogue for them) into tile
        // buffer (t - t0) & 1 while the compute warps are still on task t - 1 =====
        for (long long t = t0; t < tend; t++) {
            const int it = (int)(t - t0), bufi = it & 1;
            Rec *tile = tiles[bufi];
            float2 *tilev = tilevs[bufi];
            const int e = (int)(t / per_e);
            const int r = (int)(t - (long long)e * per_e);
            const int ob = r / g.nch, ch = r - ob * g.nch;
            const float *sxyz = str_xyz + (size_t)e * g.nstr * 3;
            if (it >= 2) bar_sync(kBarEmpty + bufi, kCtaThreads);            // the compute warps are through with this buffer
        // ---- stage the streamed chunk; evaluate the previous sweep's epilogue for its points
        const int s0 = ch * kTs;
        for (int i = threadIdx.x - kThreads; i < kTs; i += kProd) {
            const int s = s0 + i;
            const bool valid = s < g.nstr;
            const int sc = valid ? s : g.nstr - 1;
            const float x = __ldg(sxyz + sc * 3), y = __ldg(sxyz + sc * 3 + 1), z = __ldg(sxyz + sc * 3 + 2);
            float w = 0.f, v = 0.f;
            const bool writer = valid && ob == 0;      // exactly one task per (element, streamed point) stores state
            if (KIND == kA0) {
                w = p.multiR;
            } else if (KIND == kB) {
                // streamed k: remainL_lev, then ratioL_lev = remainL_lev / (1e-9 + suml_lev)   (:58, :159)
                float rem;
                if (lev == 0) rem = p.multiL;
                else {
                    const float prev = __ldcg(p.remL + ((size_t)((lev - 1) & 1) * p.b + e) * p.n + sc);
                    rem = fmaxf(0.f, __fsub_rn(prev, gather(in0, p, gprev, e, sc, ctas, 0.f)));
                }
                const float suml = gather(lev == 0 ? in0 : in1, p, gprev, e, sc, ctas, 1e-9f);
                w = __fdiv_rn(rem, suml);
                if (writer) {
                    p.remL[((size_t)(lev & 1) * p.b + e) * p.n + s] = rem;
                    p.factors[((size_t)e * kLevels + lev) * (p.n + p.m) + s] = w;
                }
            } else {
                // streamed l: sweep B's epilogue (:102-106)
                const float rem = lev == 0 ? p.multiR : __ldcg(p.remR + ((size_t)(lev & 1) * p.b + e) * p.m + sc);
                const float sumr = __fmul_rn(gather(in0, p, gprev, e, sc, ctas, 0.f), rem);
                const float cons = fminf(__fdiv_rn(rem, __fadd_rn(sumr, 1e-9f)), 1.0f);
                w = __fmul_rn(cons, rem);                        // ratioR_lev
                v = fmaxf(0.f, __fsub_rn(rem, sumr));            // remainR_{lev+1}
                if (writer) {
                    p.remR[((size_t)((lev + 1) & 1) * p.b + e) * p.m + s] = v;
                    p.factors[((size_t)e * kLevels + lev) * (p.n + p.m) + p.n + s] = w;
                }
            }
            if (!valid) { w = 0.f; v = 0.f; }                   // padding contributes exactly nothing
            Rec rec;
            rec.x0 = rec.x1 = x; rec.y0 = rec.y1 = y; rec.z0 = rec.z1 = z; rec.w0 = rec.w1 = w;
            tile[i] = rec;
            if (KIND == kCA) tilev[i] = f2(v, v);
        }

            bar_arrive(kBarFull + bufi, kCtaThreads);
        }
        // the compute warps' last (at most two) EMPTY arrivals have no refill to wait for them: take them here, so every
        // barrier is back at zero when the sweep ends
        const int nt = (int)(tend - t0);
        for (int it = nt > 2 ? nt - 2 : 0; it < nt; it++) bar_sync(kBarEmpty + (it & 1), kCtaThreads);
        return;
    }

    // ===== compute warps =====
    int held_e = -1, held_ob = -1;
    long long held_t0 = 0;
    float2 nox = f2(0, 0), noy = f2(0, 0), noz = f2(0, 0), rl = f2(0, 0), acc0 = f2(0, 0), acc1 = f2(0, 0);
    int own_i0 = 0, own_i1 = 0;

    auto flush = [&]() {
        const long long first = ((long long)held_e * g.nob + held_ob) * g.nch;
        const int slot = (int)min((long long)blockIdx.x - owner_of(first, ctas, g.tasks), held_t0 - first);
        const size_t base = ((size_t)held_e * p.nslot + slot) * p.maxnm;
        if (own_i0 < g.nown) { out0[base + own_i0] = acc0.x; if (KIND == kCA) out1[base + own_i0] = acc1.x; }
        if (own_i1 < g.nown) { out0[base + own_i1] = acc0.y; if (KIND == kCA) out1[base + own_i1] = acc1.y; }
    };

    for (long long t = t0; t < tend; t++) {
        const int bufi = (int)(t - t0) & 1;
        const Rec *tile = tiles[bufi];
        const float2 *tilev = tilevs[bufi];
        const int e = (int)(t / per_e);
        const int r = (int)(t - (long long)e * per_e);
        const int ob = r / g.nch;
        const float *oxyz = own_xyz + (size_t)e * g.nown * 3;
        if (e != held_e || ob != held_ob) {
            if (held_e >= 0) flush();
            held_e = e; held_ob = ob; held_t0 = t;
            own_i0 = ob * kOwn + 2 * threadIdx.x; own_i1 = own_i0 + 1;
            const int a = min(own_i0, g.nown - 1), c = min(own_i1, g.nown - 1);
            nox = f2(-__ldg(oxyz + a * 3), -__ldg(oxyz + c * 3));
            noy = f2(-__ldg(oxyz + a * 3 + 1), -__ldg(oxyz + c * 3 + 1));
            noz = f2(-__ldg(oxyz + a * 3 + 2), -__ldg(oxyz + c * 3 + 2));
            if (KIND == kCA || KIND == kC) {
                // ratioL_lev[k] of the own points, written by the B sweep of this level
                const float *fl = p.factors + ((size_t)e * kLevels + lev) * (p.n + p.m);
                rl = f2(__ldcg(fl + a), __ldcg(fl + c));
            }
            acc0 = f2(0, 0); acc1 = f2(0, 0);
        }

        bar_sync(kBarFull + bufi, kCtaThreads);                              // the producers have staged this task

        // ---- own pair x kTs streamed points, packed FP32
#pragma unroll kInnerUnroll
        for (int i = 0; i < kTs; i++) {
            const float4 r0 = *reinterpret_cast<const float4 *>(&tile[i].x0);
            const float4 r1 = *reinterpret_cast<const float4 *>(&tile[i].z0);
            const float2 dx = __fadd2_rn(f2(r0.x, r0.y), nox);
            const float2 dy = __fadd2_rn(f2(r0.z, r0.w), noy);
            const float2 dz = __fadd2_rn(f2(r1.x, r1.y), noz);
            const float2 d = __ffma2_rn(dz, dz, __ffma2_rn(dx, dx, __fmul2_rn(dy, dy)));
            float2 ex;
            if (KIND == kCA && kPow4) {
                // level_j = 4 level_{j+1}  =>  E_j = E_{j+1}^4: one MUFU.EX2 and two multiplies on the (idle) FMA pipe
                // instead of two MUFU.EX2 on the (saturated) SFU.  The SFU's relative error (2^-22) becomes
                // 2^-20 on E_j; 19 exponentials per pair instead of 27.
                const float2 u1 = __fmul2_rn(d, sc1);
                const float2 e1 = f2(pnae_ex2(u1.x), pnae_ex2(u1.y));
                acc1 = __ffma2_rn(e1, tilev[i], acc1);                        // sweep A of the next level
                const float2 e2 = __fmul2_rn(e1, e1);
                ex = __fmul2_rn(__fmul2_rn(e2, e2), rl);                      // (E * rl) * rr   (:151)
                acc0 = __ffma2_rn(ex, f2(r1.z, r1.w), acc0);
            } else {
                const float2 u = __fmul2_rn(d, sc0);
                ex = f2(pnae_ex2(u.x), pnae_ex2(u.y));
                if (KIND == kCA || KIND == kC) ex = __fmul2_rn(ex, rl);       // (E * rl) * rr   (:151)
                acc0 = __ffma2_rn(ex, f2(r1.z, r1.w), acc0);
                if (KIND == kCA) {
                    const float2 u1 = __fmul2_rn(d, sc1);
                    const float2 e1 = f2(pnae_ex2(u1.x), pnae_ex2(u1.y));
                    acc1 = __ffma2_rn(e1, tilev[i], acc1);                    // sweep A of the next level
                }
            }
        }
        bar_arrive(kBarEmpty + bufi, kCtaThreads);
    }
    if (held_e >= 0) flush();
}

// Last level (level == 0 => E == 1 for every pair): the three sweeps collapse to sums over points.
// One CTA per element.  Needs remainR_9 and ratioR_8 (written while sweep C of level 8 staged its
// points) and sumc_8 (ps0 of the previous stage).
__device__ void last_level(const EmdParams &p, int stage, float *red)
{
    const long long ctas = gridDim.x;
    const Geo grow = make_geo(p.b, p.n, p.m);
    const float *in0 = p.ps0 + (size_t)((stage & 1) ^ 1) * p.b * p.nslot * p.maxnm;
    const int lev = kLevels - 1;
    auto block_sum = [&](float v) -> float {      // (the compute warps only: the producers are not in here)
        v = warp_sum(v);
        bar_sync(kBarCompute, kThreads);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
        bar_sync(kBarCompute, kThreads);
        float s = 0.f;
        for (int w = 0; w < kThreads / 32; w++) s += red[w];
        return s;
    };
    for (int e = blockIdx.x; e < p.b; e += gridDim.x) {
        const float *remR9 = p.remR + ((size_t)(lev & 1) * p.b + e) * p.m;
        float *fac = p.factors + ((size_t)e * kLevels + lev) * (p.n + p.m);
        float s = 0.f;
        for (int l = threadIdx.x; l < p.m; l += kThreads) s += __ldcg(remR9 + l);
        const float suml = __fadd_rn(1e-9f, block_sum(s));                 // 1e-9 + sum_l 1 * remainR[l]
        s = 0.f;
        for (int k = threadIdx.x; k < p.n; k += kThreads) {
            const float prev = __ldcg(p.remL + ((size_t)((lev - 1) & 1) * p.b + e) * p.n + k);
            const float rem = fmaxf(0.f, __fsub_rn(prev, gather(in0, p, grow, e, k, ctas, 0.f)));
            const float rl = __fdiv_rn(rem, suml);
            fac[k] = rl;
            s += rl;
        }
        const float sumk = block_sum(s);                                   // sum_k 1 * ratioL[k]
        for (int l = threadIdx.x; l < p.m; l += kThreads) {
            const float rem = __ldcg(remR9 + l);
            const float sumr = __fmul_rn(sumk, rem);
            const float cons = fminf(__fdiv_rn(rem, __fadd_rn(sumr, 1e-9f)), 1.0f);
            fac[p.n + l] = __fmul_rn(cons, rem);
        }
    }
}

#ifdef PNAE_AM_TRACE                   // tuning builds only (tools/am_trace.py): SM-clock stamps of one CTA's sweeps and barriers
__device__ long long g_am_trace[3][64];
__device__ long long g_am_all[1024][4];        // per CTA: smid, duration of level 3's B sweep, of its C/A sweep, tasks
extern "C" __attribute__((visibility("default"))) int pnae_debug_am_trace(long long *host)
{
    return (int)cudaMemcpyFromSymbol(host, g_am_trace, sizeof(g_am_trace));
}
extern "C" __attribute__((visibility("default"))) int pnae_debug_am_all(long long *host)
{
    return (int)cudaMemcpyFromSymbol(host, g_am_all, sizeof(g_am_all));
}
#define AM_TRACE(i) do { if (trace_cta >= 0 && threadIdx.x == 0 && (i) < 64) g_am_trace[trace_cta][i] = clock64(); } while (0)
#else
#define AM_TRACE(i) do { } while (0)
#endif

__global__ void __launch_bounds__(kCtaThreads, kCtasPerSm)
approx_match_kernel(const EmdParams p)
{
#ifdef PNAE_AM_TRACE
    const int trace_cta = blockIdx.x == 0 ? 0 : blockIdx.x == 1 ? 1 : blockIdx.x == gridDim.x - 1 ? 2 : -1;
    int ti = 0;
#endif
    __shared__ Rec tile[2][kTs];                 // double-buffered: the producers stage one task ahead
    __shared__ float2 tilev[2][kTs];
    __shared__ float red[kThreads / 32];
    cg::grid_group grid = cg::this_grid();
    int stage = 0;
#ifdef PNAE_AM_TRACE
#define AM_STAMP() do { AM_TRACE(ti); ti++; } while (0)
#else
#define AM_STAMP() do { } while (0)
#endif
    AM_STAMP();
    sweep<kA0>(p, 0, stage++, tile, tilev);
    AM_STAMP();
    __threadfence();
    grid.sync();
    for (int lev = 0; lev < kLevels - 1; lev++) {
        AM_STAMP();
#ifdef PNAE_AM_TRACE
        long long tb0 = clock64();
#endif
        sweep<kB>(p, lev, stage++, tile, tilev);
#ifdef PNAE_AM_TRACE
        if (lev == 3 && threadIdx.x == 0 && blockIdx.x < 1024) {
            unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            g_am_all[blockIdx.x][0] = smid; g_am_all[blockIdx.x][1] = clock64() - tb0;
        }
#endif
        AM_STAMP();
        __threadfence();
        grid.sync();
        AM_STAMP();
#ifdef PNAE_AM_TRACE
        long long tc0 = clock64();
#endif
        if (lev < kLevels - 2) sweep<kCA>(p, lev, stage++, tile, tilev);
        else sweep<kC>(p, lev, stage++, tile, tilev);
#ifdef PNAE_AM_TRACE
        if (lev == 3 && threadIdx.x == 0 && blockIdx.x < 1024) g_am_all[blockIdx.x][2] = clock64() - tc0;
#endif
        AM_STAMP();
        __threadfence();
        grid.sync();
    }
    AM_STAMP();
    if (threadIdx.x < kThreads) last_level(p, stage, red);
    AM_STAMP();
}

struct EmdPlan {
    int grid, nslot, maxnm;
    size_t rem_floats, ps_floats, total;
};

EmdPlan make_plan(int b, int n, int m, int sms)
{
    EmdPlan pl;
    pl.grid = sms * kCtasPerSm;
    pl.maxnm = n > m ? n : m;
    auto slots = [&](int nown, int nstr) -> int {
        const long long nob = (nown + kOwn - 1) / kOwn, nch = (nstr + kTs - 1) / kTs;
        const long long span = max(1ll, (long long)b * nob * nch / pl.grid);
        return (int)min(nch, (nch + span - 1) / span + 1);
    };
    pl.nslot = max(slots(n, m), slots(m, n));
    pl.rem_floats = 2 * (size_t)b * ((size_t)n + m);
    pl.ps_floats = 2 * (size_t)b * pl.nslot * pl.maxnm;
    pl.total = sizeof(float) * (pl.rem_floats + 2 * pl.ps_floats);
    return pl;
}

}  // namespace

int pnae_match_from_factors_impl(int b, int n, int m, const float *xyz1, const float *xyz2,
                                 const float *factors, float *match, cudaStream_t st);

extern "C" int pnae_approx_match_plan(int b, int n, int m, int sm_count, int *plan)
{
    PNAE_REQUIRE(b >= 1 && n >= 1 && m >= 1 && sm_count >= 1 && plan != nullptr, "approx_match_plan: invalid argument");
    const EmdPlan pl = make_plan(b, n, m, sm_count);
    plan[0] = pl.grid; plan[1] = pl.nslot; plan[2] = pl.maxnm; plan[3] = kOwn; plan[4] = kTs; plan[5] = kThreads;
    return PNAE_OK;
}

extern "C" size_t pnae_approx_match_workspace_bytes(int b, int n, int m)
{
    if (b <= 0 || n <= 0 || m <= 0) return 0;
    return make_plan(b, n, m, pnae_sm_count()).total;
}

extern "C" int pnae_approx_match(int b, int n, int m, const float *xyz1, const float *xyz2,
                                 float *factors, float *match,
                                 void *workspace, size_t workspace_bytes, void *stream)
{
    PNAE_REQUIRE(b >= 0 && n >= 1 && m >= 1, "approx_match: need b>=0, n>=1, m>=1 (got b=%d n=%d m=%d)", b, n, m);
    PNAE_REQUIRE(xyz1 && xyz2 && factors, "approx_match: NULL pointer (factors is required)");
    if (b == 0) return PNAE_OK;
    const EmdPlan pl = make_plan(b, n, m, pnae_sm_count());
    if (workspace == nullptr || workspace_bytes < pl.total) {
        pnae_set_error("approx_match: workspace too small (%zu < %zu bytes)", workspace_bytes, pl.total);
        return PNAE_ERR_WORKSPACE;
    }
    PNAE_REQUIRE(pnae_aligned(workspace, 4), "approx_match: workspace must be 4-byte aligned");
    EmdParams p;
    p.b = b; p.n = n; p.m = m;
    p.xyz1 = xyz1; p.xyz2 = xyz2; p.factors = factors;
    float *ws = (float *)workspace;
    p.remL = ws;
    p.remR = ws + 2 * (size_t)b * n;
    p.ps0 = ws + pl.rem_floats;
    p.ps1 = p.ps0 + pl.ps_floats;
    p.nslot = pl.nslot; p.maxnm = pl.maxnm;
    p.multiL = (n >= m) ? 1.0f : (float)(m / n);      // integer division, tf_approxmatch_g.cu:4-10
    p.multiR = (n >= m) ? (float)(n / m) : 1.0f;
    cudaStream_t st = (cudaStream_t)stream;
    void *args[] = {(void *)&p};
    PNAE_CUDA_OK(cudaLaunchCooperativeKernel((const void *)approx_match_kernel, dim3(pl.grid), dim3(kCtaThreads), args, 0, st));
    if (match != nullptr) return pnae_match_from_factors_impl(b, n, m, xyz1, xyz2, factors, match, st);
    return PNAE_OK;
}
