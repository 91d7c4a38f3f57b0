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
//  * a persistent sweep launch: warps pull (element, 32*R rows, kCols columns) tasks from
//    an atomic queue; partial (min, tag) keys go to an L2-resident workspace with plain
//    coalesced stores; a second, fully parallel launch (programmatic dependent launch)
//    reduces the partials of every point and re-evaluates the <=32 (rows) / R (columns)
//    tagged candidates to emit the exact first argmin.
#include <cooperative_groups.h>

#include "pnae_common.cuh"

namespace cg = cooperative_groups;

namespace {

typedef unsigned long long u64;

constexpr int kR = 8;                 // rows per lane (contiguous: lane l owns rows l*kR .. l*kR+kR-1 of the block)
constexpr int kRowsPerTask = 32 * kR; // 256
#ifndef PNAE_NN_COLS
#define PNAE_NN_COLS 64
#endif
constexpr int kCols = PNAE_NN_COLS;   // columns per task
constexpr int kChunk = 32;            // columns per row-argmin tag
constexpr int kWarps = 4;             // warps per CTA (independent; no CTA-wide barrier anywhere)
#ifndef PNAE_NN_CTAS
#define PNAE_NN_CTAS 4
#endif
constexpr int kCtasPerSm = PNAE_NN_CTAS;
constexpr size_t kWsBudget = 256ull << 20;

struct FwdParams {
    int be;            // elements in this launch
    int n, m;
    int nrb, ncr;      // row blocks / column ranges per element
    const float *xyz1, *xyz2;
    float *dist1, *dist2;
    int *idx1, *idx2;
    u64 *rowkeys;      // [be][ncr][n]  (min bits << 32 | global chunk index)
    u64 *colkeys;      // [be][nrb][m]  (min bits << 32 | ballot of lanes holding the min)
    int *ctr;          // [0] task queue head
};

// cp.async (LDGSTS) helpers: 4-byte granularity because points are 12-byte xyz triples
__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gmem_src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Start the asynchronous copy of a main task's operands into this warp's staging buffers:
// its 32*kR rows as flat xyz floats, its kCols columns as float4 slots (w unused).
__device__ __forceinline__ void prefetch_task(const FwdParams &p, int e, int rb, int cr, float *srow, float4 *scol)
{
    const int lane = threadIdx.x & 31;
    const float *p1 = p.xyz1 + (size_t)e * p.n * 3;
    const float *p2 = p.xyz2 + (size_t)e * p.m * 3;
    const int row0 = rb * kRowsPerTask, col0 = cr * kCols;
#pragma unroll
    for (int i = 0; i < kRowsPerTask * 3 / 32; i++) {
        const int f = lane + 32 * i, r = f / 3;
        const int j = min(row0 + r, p.n - 1);              // clamped duplicates never change a minimum
        cp_async4(srow + f, p1 + (size_t)j * 3 + (f - r * 3));
    }
#pragma unroll
    for (int i = 0; i < kCols * 3 / 32; i++) {
        const int f = lane + 32 * i, c = f / 3;
        const int k = min(col0 + c, p.m - 1);
        cp_async4(reinterpret_cast<float *>(scol + c) + (f - c * 3), p2 + (size_t)k * 3 + (f - c * 3));
    }
    cp_async_commit();
}

// Sweep one task: 32*kR rows (kR per lane, in registers) against kCols staged columns.
// Leaves the partial keys in the workspace; completion is signalled later (see the kernel).
__device__ __forceinline__ void main_task(const FwdParams &p, int e, int rb, int cr,
                                          const float4 *scol, u64 *skey,
                                          const float (&rx)[kR], const float (&ry)[kR], const float (&rz)[kR])
{
    const int lane = threadIdx.x & 31;
    float best[kR], snap[kR];
    int tag[kR];
    const int row0 = rb * kRowsPerTask + lane * kR;
#pragma unroll
    for (int r = 0; r < kR; r++) {
        best[r] = snap[r] = __int_as_float(0x7f800000);
        tag[r] = 0;
    }
    const int col0 = cr * kCols;

    for (int ch = 0; ch < kCols / kChunk; ch++) {
#pragma unroll 4
        for (int cc = 0; cc < kChunk; cc++) {
            const int c = ch * kChunk + cc;
            const float4 q = scol[c];
            float d[kR];
#pragma unroll
            for (int r = 0; r < kR; r++) {
                d[r] = pnae_sqdist(q.x - rx[r], q.y - ry[r], q.z - rz[r]);
                best[r] = fminf(best[r], d[r]);
            }
            // column minimum over this lane's rows (tree), then over the warp
#pragma unroll
            for (int s = kR / 2; s > 0; s >>= 1)
#pragma unroll
                for (int r = 0; r < s; r++) d[r] = fminf(d[r], d[r + s]);
            const unsigned bits = __float_as_uint(d[0]);      // d >= 0: unsigned order == float order
            const unsigned mn = __reduce_min_sync(0xffffffffu, bits);
            const unsigned who = __ballot_sync(0xffffffffu, bits == mn);
            if (lane == 0) skey[c] = ((u64)mn << 32) | who;
        }
        // a strict decrease during this chunk => the row's running minimum first appears here
#pragma unroll
        for (int r = 0; r < kR; r++) {
            if (best[r] < snap[r]) tag[r] = ch;
            snap[r] = best[r];
        }
    }
    __syncwarp();
    // partial keys: [cr][row] and [rb][col], coalesced
    u64 *rk = p.rowkeys + ((size_t)e * p.ncr + cr) * p.n;
#pragma unroll
    for (int r = 0; r < kR; r++) {
        const int j = row0 + r;
        if (j < p.n) rk[j] = ((u64)__float_as_uint(best[r]) << 32) | (unsigned)(cr * (kCols / kChunk) + tag[r]);
    }
    u64 *ck = p.colkeys + ((size_t)e * p.nrb + rb) * p.m;
#pragma unroll
    for (int c = lane; c < kCols; c += 32)
        if (col0 + c < p.m) ck[col0 + c] = skey[c];
}

// One persistent launch for the sweep.  Each warp is an independent worker (no CTA-wide barrier
// anywhere): tasks (element, row block, column range) come from one atomic queue, claimed TWO
// ahead, so the atomic's round trip and the cp.async operand copy of the next task hide under
// the current task's FP32 work.
__global__ void __launch_bounds__(kWarps * 32, kCtasPerSm)
nn_fwd_kernel(const FwdParams p)
{
    __shared__ __align__(16) float4 scol_all[kWarps][2][kCols];
    __shared__ __align__(16) float srow_all[kWarps][kRowsPerTask * 3];
    __shared__ u64 skey_all[kWarps][kCols];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int main_per_e = p.nrb * p.ncr;
    const long long n_main = (long long)p.be * main_per_e;
    float *srow = srow_all[warp];

    auto decode = [&](long long t, int &e, int &rb, int &cr) {
        e = (int)(t / main_per_e);
        const int r = (int)(t - (long long)e * main_per_e);
        rb = r / p.ncr;          // column range fastest: neighbouring tasks share the row block
        cr = r - rb * p.ncr;
    };
    auto claim = [&]() -> unsigned { return lane == 0 ? atomicAdd((unsigned *)p.ctr, 1u) : 0u; };
    auto bcast = [&](unsigned v) -> long long { return (long long)__shfl_sync(0xffffffffu, v, 0); };

    long long cur = bcast(claim());
    long long nxt = bcast(claim());
    int buf = 0;
    if (cur < n_main) {
        int e, rb, cr;
        decode(cur, e, rb, cr);
        prefetch_task(p, e, rb, cr, srow, scol_all[warp][buf]);
    }
    while (cur < n_main) {
        const unsigned pend = claim();            // task after next; consumed at the bottom of the loop
        int e, rb, cr;
        decode(cur, e, rb, cr);
        cp_async_wait_all();
        __syncwarp();
        float rx[kR], ry[kR], rz[kR];
        {
            const float4 *src = reinterpret_cast<const float4 *>(srow + lane * kR * 3);
            float tmp[kR * 3];
#pragma unroll
            for (int i = 0; i < kR * 3 / 4; i++) {
                const float4 v = src[i];
                tmp[4 * i] = v.x; tmp[4 * i + 1] = v.y; tmp[4 * i + 2] = v.z; tmp[4 * i + 3] = v.w;
            }
#pragma unroll
            for (int r = 0; r < kR; r++) { rx[r] = tmp[3 * r]; ry[r] = tmp[3 * r + 1]; rz[r] = tmp[3 * r + 2]; }
        }
        __syncwarp();      // every lane has its rows in registers: the row buffer may be refilled
        if (nxt < n_main) {
            int e2, rb2, cr2;
            decode(nxt, e2, rb2, cr2);
            prefetch_task(p, e2, rb2, cr2, srow, scol_all[warp][buf ^ 1]);
        }
        main_task(p, e, rb, cr, scol_all[warp][buf], skey_all[warp], rx, ry, rz);
        buf ^= 1;
        __syncwarp();
        cur = nxt;
        nxt = bcast(pend);
    }
}

// Second (tiny, fully parallel) launch: kFinLanes lanes per output point.  Each group reduces the
// point's partial keys, then re-evaluates the tagged candidates (32 columns for a point of xyz1,
// kR rows for a point of xyz2) with the same arithmetic as the sweep; the lowest index whose
// distance equals the minimum is the reference's first argmin.  All loads of a phase are
// independent, so a point costs three dependent L2 round trips.
constexpr int kFinLanes = 4;
constexpr int kFinThreads = 256;

__device__ __forceinline__ u64 group_min_u64(u64 v)
{
#pragma unroll
    for (int o = kFinLanes / 2; o > 0; o >>= 1) {
        const u64 w = __shfl_xor_sync(0xffffffffu, v, o);
        v = min(v, w);
    }
    return v;
}

__global__ void __launch_bounds__(kFinThreads)
nn_finalize_kernel(const FwdParams p)
{
#if __CUDA_ARCH__ >= 900
    cudaGridDependencySynchronize();      // launched with programmatic stream serialization
#endif
    const int sub = threadIdx.x & (kFinLanes - 1);
    const long long per_e = (long long)p.n + p.m;
    const long long total = (long long)p.be * per_e;
    constexpr int kPtsPerWarp = 32 / kFinLanes;
    const long long warp_id = ((long long)blockIdx.x * kFinThreads + threadIdx.x) >> 5;
    const long long n_warps = (long long)gridDim.x * kFinThreads >> 5;
    for (long long base = warp_id * kPtsPerWarp; base < total; base += n_warps * kPtsPerWarp) {   // warp-uniform trip count
        const long long pt = base + (threadIdx.x & 31) / kFinLanes;
        const bool live = pt < total;
        const long long q = live ? pt : total - 1;
        const int e = (int)(q / per_e);
        const int r = (int)(q - (long long)e * per_e);
        const float *p1 = p.xyz1 + (size_t)e * p.n * 3;
        const float *p2 = p.xyz2 + (size_t)e * p.m * 3;
        if (r < p.n) {
            // point j of xyz1 -> dist1 / idx1
            const int j = r;
            const u64 *rk = p.rowkeys + (size_t)e * p.ncr * p.n + j;
            u64 key = ~0ull;
            for (int cr = sub; cr < p.ncr; cr += 4 * kFinLanes) {
                u64 v[4];
#pragma unroll
                for (int u = 0; u < 4; u++) v[u] = (cr + u * kFinLanes < p.ncr) ? __ldcg(rk + (size_t)(cr + u * kFinLanes) * p.n) : ~0ull;
#pragma unroll
                for (int u = 0; u < 4; u++) key = min(key, v[u]);
            }
            key = group_min_u64(key);
            const float want = __uint_as_float((unsigned)(key >> 32));
            const int k0 = (int)(unsigned)key * kChunk;
            const float x = __ldg(p1 + j * 3), y = __ldg(p1 + j * 3 + 1), z = __ldg(p1 + j * 3 + 2);
            constexpr int kPer = kChunk / kFinLanes;          // candidates per lane, contiguous
            int found = 0x7fffffff;
#pragma unroll
            for (int c = kPer - 1; c >= 0; c--) {
                const int k = min(k0 + sub * kPer + c, p.m - 1);
                const float d = pnae_sqdist(__ldg(p2 + k * 3) - x, __ldg(p2 + k * 3 + 1) - y, __ldg(p2 + k * 3 + 2) - z);
                if (d == want) found = k;
            }
#pragma unroll
            for (int o = kFinLanes / 2; o > 0; o >>= 1) found = min(found, __shfl_xor_sync(0xffffffffu, found, o));
            if (live && sub == 0) {
                p.dist1[(size_t)e * p.n + j] = want;
                p.idx1[(size_t)e * p.n + j] = found == 0x7fffffff ? min(k0, p.m - 1) : found;
            }
        } else {
            // point k of xyz2 -> dist2 / idx2
            const int k = r - p.n;
            const u64 *ck = p.colkeys + (size_t)e * p.nrb * p.m + k;
            u64 key = ~0ull;       // (min bits, row block) first: the lowest row block wins ties
            for (int rb = sub; rb < p.nrb; rb += 4 * kFinLanes) {
                u64 v[4];
#pragma unroll
                for (int u = 0; u < 4; u++) v[u] = (rb + u * kFinLanes < p.nrb) ? __ldcg(ck + (size_t)(rb + u * kFinLanes) * p.m) : ~0ull;
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const u64 cand = (v[u] & 0xffffffff00000000ull) | (unsigned)(rb + u * kFinLanes);
                    if (v[u] != ~0ull) key = min(key, cand);
                }
            }
            key = group_min_u64(key);
            const int rbw = (int)(unsigned)key;
            const unsigned who = (unsigned)__ldcg(ck + (size_t)rbw * p.m);      // ballot of lanes that held the min
            const float want = __uint_as_float((unsigned)(key >> 32));
            const int j0 = rbw * kRowsPerTask + (__ffs(who) - 1) * kR;
            const float x = __ldg(p2 + k * 3), y = __ldg(p2 + k * 3 + 1), z = __ldg(p2 + k * 3 + 2);
            constexpr int kPer = kR / kFinLanes;
            int found = 0x7fffffff;
#pragma unroll
            for (int c = kPer - 1; c >= 0; c--) {
                const int j = min(j0 + sub * kPer + c, p.n - 1);
                const float d = pnae_sqdist(x - __ldg(p1 + j * 3), y - __ldg(p1 + j * 3 + 1), z - __ldg(p1 + j * 3 + 2));
                if (d == want) found = j;
            }
#pragma unroll
            for (int o = kFinLanes / 2; o > 0; o >>= 1) found = min(found, __shfl_xor_sync(0xffffffffu, found, o));
            if (live && sub == 0) {
                p.dist2[(size_t)e * p.m + k] = want;
                p.idx2[(size_t)e * p.m + k] = found == 0x7fffffff ? min(j0, p.n - 1) : found;
            }
        }
    }
}

struct FwdPlan {
    int nrb, ncr, be;
    size_t row_bytes, col_bytes, ctr_bytes;   // per launch chunk of `be` elements
    size_t total;
};

FwdPlan make_plan(int b, int n, int m)
{
    FwdPlan pl;
    pl.nrb = (n + kRowsPerTask - 1) / kRowsPerTask;
    pl.ncr = (m + kCols - 1) / kCols;
    const size_t per_e = sizeof(u64) * ((size_t)pl.ncr * n + (size_t)pl.nrb * m);
    long long be = (long long)(kWsBudget / (per_e ? per_e : 1));
    // keep the task / counter arithmetic inside 32 bits
    const long long tasks_per_e = (long long)pl.nrb * pl.ncr;
    be = min(be, (long long)(0x7ff00000 / tasks_per_e));   // queue head (+ one overshoot per warp) stays inside 31 bits
    pl.be = (int)max(1ll, min((long long)b, be));
    pl.row_bytes = sizeof(u64) * (size_t)pl.be * pl.ncr * n;
    pl.col_bytes = sizeof(u64) * (size_t)pl.be * pl.nrb * m;
    pl.ctr_bytes = 256;
    pl.total = pl.row_bytes + pl.col_bytes + pl.ctr_bytes;
    return pl;
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
    const int stride = cs * kBwdThreads;
    for (int e = blockIdx.x / cs; e < b; e += nclusters) {
        const float *p1 = xyz1 + (size_t)e * n * 3, *p2 = xyz2 + (size_t)e * m * 3;
        float *g1 = grad_xyz1 + (size_t)e * n * 3, *g2 = grad_xyz2 + (size_t)e * m * 3;
        for (int phase = 0; phase < 2; phase++) {
            for (int t = tid; t < n + m; t += stride) {
                const bool second = t >= n;
                const int j = second ? t - n : t;
                const float *a = (second ? p2 : p1) + j * 3;
                const int j2 = second ? idx2[(size_t)e * m + j] : idx1[(size_t)e * n + j];
                const float *c = (second ? p1 : p2) + j2 * 3;
                const float g = __fmul_rn(second ? grad_dist2[(size_t)e * m + j] : grad_dist1[(size_t)e * n + j], 2.0f);
                const float vx = __fmul_rn(g, __fsub_rn(__ldg(a), __ldg(c)));
                const float vy = __fmul_rn(g, __fsub_rn(__ldg(a + 1), __ldg(c + 1)));
                const float vz = __fmul_rn(g, __fsub_rn(__ldg(a + 2), __ldg(c + 2)));
                if (phase == 0) {
                    float *ga = (second ? g2 : g1) + j * 3;
                    ga[0] = vx; ga[1] = vy; ga[2] = vz;
                } else {
                    float *gc = (second ? g1 : g2) + j2 * 3;
                    atomicAdd(gc, -vx); atomicAdd(gc + 1, -vy); atomicAdd(gc + 2, -vz);
                }
            }
            __threadfence();
            cluster.sync();
        }
    }
}

}  // namespace

extern "C" size_t pnae_nn_distance_workspace_bytes(int b, int n, int m)
{
    if (b <= 0 || n <= 0 || m <= 0) return 0;
    return make_plan(b, n, m).total;
}

extern "C" int pnae_nn_distance_fwd(int b, int n, const float *xyz1, int m, const float *xyz2,
                                    float *dist1, int *idx1, float *dist2, int *idx2,
                                    void *workspace, size_t workspace_bytes, void *stream)
{
    PNAE_REQUIRE(b >= 0 && n >= 1 && m >= 1, "nn_distance: need b>=0, n>=1, m>=1 (got b=%d n=%d m=%d)", b, n, m);
    PNAE_REQUIRE(xyz1 && xyz2 && dist1 && idx1 && dist2 && idx2, "nn_distance: NULL pointer");
    if (b == 0) return PNAE_OK;
    const FwdPlan pl = make_plan(b, n, m);
    if (workspace == nullptr || workspace_bytes < pl.total) {
        pnae_set_error("nn_distance: workspace too small (%zu < %zu bytes)", workspace_bytes, pl.total);
        return PNAE_ERR_WORKSPACE;
    }
    PNAE_REQUIRE(pnae_aligned(workspace, 8), "nn_distance: workspace must be 8-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    char *ws = (char *)workspace;
    const int grid = pnae_sm_count() * kCtasPerSm;
    for (int e0 = 0; e0 < b; e0 += pl.be) {
        FwdParams p;
        p.be = min(pl.be, b - e0);
        p.n = n; p.m = m; p.nrb = pl.nrb; p.ncr = pl.ncr;
        p.xyz1 = xyz1 + (size_t)e0 * n * 3; p.xyz2 = xyz2 + (size_t)e0 * m * 3;
        p.dist1 = dist1 + (size_t)e0 * n; p.idx1 = idx1 + (size_t)e0 * n;
        p.dist2 = dist2 + (size_t)e0 * m; p.idx2 = idx2 + (size_t)e0 * m;
        p.rowkeys = (u64 *)ws;
        p.colkeys = (u64 *)(ws + pl.row_bytes);
        p.ctr = (int *)(ws + pl.row_bytes + pl.col_bytes);
        PNAE_CUDA_OK(cudaMemsetAsync(p.ctr, 0, sizeof(int), st));
        nn_fwd_kernel<<<grid, kWarps * 32, 0, st>>>(p);
        PNAE_CUDA_OK(cudaGetLastError());
        {
            const long long groups = (long long)p.be * ((long long)n + m);
            const long long want_blocks = (groups * kFinLanes + kFinThreads - 1) / kFinThreads;
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned)min(want_blocks, (long long)pnae_sm_count() * 8));
            cfg.blockDim = dim3(kFinThreads);
            cfg.stream = st;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // launch latency overlaps the sweep's tail
            attr[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            PNAE_CUDA_OK(cudaLaunchKernelEx(&cfg, nn_finalize_kernel, p));
        }
    }
    return PNAE_OK;
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
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    PNAE_CUDA_OK(cudaLaunchKernelEx(&cfg, nn_bwd_kernel, b, n, xyz1, m, xyz2, grad_dist1, idx1, grad_dist2, idx2,
                                    grad_xyz1, grad_xyz2));
    return PNAE_OK;
}
