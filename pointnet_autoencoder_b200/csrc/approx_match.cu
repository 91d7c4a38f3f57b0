// approx_match.cu -- approximate EMD soft assignment for sm_100a.
//
// Replaces approxmatch / approxmatchLauncher (reference:
// tf_ops/approxmatch/tf_approxmatch_g.cu:1-182).  Same schedule: 10 levels
// j=7..-2, three sweeps per level (A: ratioL, B: ratioR/remainR, C: match/remainL),
// fp32 accumulators, the same FMA contractions the reference compiles to.
//
// What is different from the reference kernel:
//  * one thread-block CLUSTER per batch element instead of one CTA: dataset rows
//    (sweeps A, C) and query columns (sweep B) are split over the CTAs of the
//    cluster, which meet at a hardware cluster barrier between sweeps; only the
//    n- or m-long state vectors cross CTAs (through L2);
//  * the per-level factors ratioL_j / ratioR_j are the primary output; the dense
//    (b,m,n) tensor is only touched when the caller asks for it;
//  * 2^t on the SFU with flush-to-zero (see pnae_ex2), one multiply for
//    level*log2e, R owned points per thread so one LDS.128 feeds R pairs.
#include <cooperative_groups.h>

#include "pnae_common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int kThreads = 256;
constexpr int kTile = 1024;   // streamed points per shared-memory tile (float4: x,y,z,weight)
constexpr int kR = 2;         // owned points per thread per pass

enum SweepKind { kSweepA = 0, kSweepB = 1, kSweepC = 2 };

// Stream `ns` points (coordinates `sp`, weights `sw`) past the owned points
// [lo,hi) of `op`; acc_r = sum over streamed points, in index order, of
//   A,B: fma(E, w, acc)        C: fma(E*rl_r, w, acc)   [+ match RMW]
// then apply the sweep's epilogue to each owned point.
template <int KIND>
__device__ __forceinline__ void sweep(float4 *tile, float scale,
                                      const float *__restrict__ op, int lo, int hi,
                                      const float *__restrict__ sp, const float *sw, int ns,
                                      float *remainL, float *remainR, float *ratioL, float *ratioR,
                                      float *match_i, int n)
{
    for (int base = lo; base < hi; base += kThreads * kR) {
        float ox[kR], oy[kR], oz[kR], acc[kR], rl[kR];
        int own[kR];
#pragma unroll
        for (int r = 0; r < kR; r++) {
            own[r] = base + r * kThreads + (int)threadIdx.x;
            const int j = min(own[r], hi - 1);
            ox[r] = __ldg(op + j * 3 + 0);
            oy[r] = __ldg(op + j * 3 + 1);
            oz[r] = __ldg(op + j * 3 + 2);
            acc[r] = (KIND == kSweepA) ? 1e-9f : 0.0f;
            rl[r] = (KIND == kSweepC) ? __ldcg(ratioL + j) : 0.0f;
        }
        for (int s0 = 0; s0 < ns; s0 += kTile) {
            const int cnt = min(kTile, ns - s0);
            __syncthreads();
            for (int t = threadIdx.x; t < cnt; t += kThreads) {
                const float *p = sp + (size_t)(s0 + t) * 3;
                tile[t] = make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), __ldcg(sw + s0 + t));
            }
            __syncthreads();
#pragma unroll 4
            for (int t = 0; t < cnt; t++) {
                const float4 p = tile[t];
#pragma unroll
                for (int r = 0; r < kR; r++) {
                    const float d = pnae_sqdist(p.x - ox[r], p.y - oy[r], p.z - oz[r]);
                    float e = pnae_ex2(__fmul_rn(d, scale));
                    if (KIND == kSweepC) {
                        e = __fmul_rn(e, rl[r]);
                        if (match_i != nullptr && own[r] < hi) {
                            float *mp = match_i + (size_t)(s0 + t) * n + own[r];
                            *mp = __fmaf_rn(e, p.w, *mp);
                        }
                    }
                    acc[r] = __fmaf_rn(e, p.w, acc[r]);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < kR; r++) {
            const int j = own[r];
            if (j >= hi) continue;
            if (KIND == kSweepA) {
                ratioL[j] = __fdiv_rn(remainL[j], acc[r]);                       // :58
            } else if (KIND == kSweepB) {
                const float rr = remainR[j];
                const float sumr = __fmul_rn(acc[r], rr);                        // :102
                const float consumption = fminf(__fdiv_rn(rr, __fadd_rn(sumr, 1e-9f)), 1.0f);
                ratioR[j] = __fmul_rn(consumption, rr);                          // :104
                remainR[j] = fmaxf(0.0f, __fsub_rn(rr, sumr));                   // :105
            } else {
                remainL[j] = fmaxf(0.0f, __fsub_rn(remainL[j], acc[r]));         // :159
            }
        }
    }
}

__global__ void __launch_bounds__(kThreads, 3)
approx_match_kernel(int b, int n, int m, const float *__restrict__ xyz1, const float *__restrict__ xyz2,
                    float *__restrict__ factors, float *match, float *ws)
{
    __shared__ float4 tile[kTile];
    cg::cluster_group cluster = cg::this_cluster();
    const int cs = (int)cluster.num_blocks();
    const int crank = (int)cluster.block_rank();
    const int ncluster = gridDim.x / cs;
    const int cid = blockIdx.x / cs;

    // integer division, tf_approxmatch_g.cu:4-10
    const float multiL = (n >= m) ? 1.0f : (float)(m / n);
    const float multiR = (n >= m) ? (float)(n / m) : 1.0f;

    const int nper = (n + cs - 1) / cs, mper = (m + cs - 1) / cs;
    const int klo = min(n, crank * nper), khi = min(n, klo + nper);
    const int llo = min(m, crank * mper), lhi = min(m, llo + mper);

    for (int i = cid; i < b; i += ncluster) {
        const float *p1 = xyz1 + (size_t)i * n * 3;
        const float *p2 = xyz2 + (size_t)i * m * 3;
        float *remainL = ws + (size_t)i * (n + m);
        float *remainR = remainL + n;
        float *match_i = match ? match + (size_t)i * n * m : nullptr;

        for (int k = klo + threadIdx.x; k < khi; k += kThreads) remainL[k] = multiL;
        for (int l = llo + threadIdx.x; l < lhi; l += kThreads) remainR[l] = multiR;
        if (match_i) {
            // this CTA zeroes the rows l of its column slice (contiguous (lhi-llo)*n floats)
            float *z = match_i + (size_t)llo * n;
            const size_t cnt = (size_t)(lhi - llo) * n;
            for (size_t t = threadIdx.x; t < cnt; t += kThreads) z[t] = 0.0f;
        }
        __threadfence();
        cluster.sync();

        for (int lev = 0; lev < PNAE_NUM_LEVELS; lev++) {
            const float scale = pnae_level_scale(lev);
            float *ratioL = factors + ((size_t)i * PNAE_NUM_LEVELS + lev) * (n + m);
            float *ratioR = ratioL + n;
            sweep<kSweepA>(tile, scale, p1, klo, khi, p2, remainR, m, remainL, remainR, ratioL, ratioR, nullptr, n);
            __threadfence();
            cluster.sync();
            sweep<kSweepB>(tile, scale, p2, llo, lhi, p1, ratioL, n, remainL, remainR, ratioL, ratioR, nullptr, n);
            __threadfence();
            cluster.sync();
            sweep<kSweepC>(tile, scale, p1, klo, khi, p2, ratioR, m, remainL, remainR, ratioL, ratioR, match_i, n);
            // next sweep A reads remainR (written in B, already synchronised) and this
            // CTA's own remainL; the match RMW of a thread only touches its own k.
            __syncthreads();
        }
        __threadfence();
        cluster.sync();   // workspace of this cluster is reused by its next element
    }
}

}  // namespace

extern "C" size_t pnae_approx_match_workspace_bytes(int b, int n, int m)
{
    if (b <= 0 || n <= 0 || m <= 0) return 0;
    return sizeof(float) * (size_t)b * ((size_t)n + (size_t)m);
}

extern "C" int pnae_approx_match(int b, int n, int m, const float *xyz1, const float *xyz2,
                                 float *factors, float *match,
                                 void *workspace, size_t workspace_bytes, void *stream)
{
    PNAE_REQUIRE(b >= 0 && n >= 1 && m >= 1, "approx_match: need b>=0, n>=1, m>=1 (got b=%d n=%d m=%d)", b, n, m);
    PNAE_REQUIRE(xyz1 && xyz2 && factors, "approx_match: NULL pointer (factors is required)");
    if (b == 0) return PNAE_OK;
    const size_t need = pnae_approx_match_workspace_bytes(b, n, m);
    if (workspace == nullptr || workspace_bytes < need) {
        pnae_set_error("approx_match: workspace too small (%zu < %zu bytes)", workspace_bytes, need);
        return PNAE_ERR_WORKSPACE;
    }
    const int sms = pnae_sm_count();
    int cs = 8;
    while (cs > 1 && (long long)b * cs > 2ll * sms) cs >>= 1;
    while (cs > 1 && (n + cs - 1) / cs < 32 && (m + cs - 1) / cs < 32) cs >>= 1;   // tiny clouds: fewer, fuller CTAs
    const int nclusters = (int)min((long long)b, (long long)(4 * sms / cs > 0 ? 4 * sms / cs : 1));

    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(nclusters * cs));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    PNAE_CUDA_OK(cudaLaunchKernelEx(&cfg, approx_match_kernel, b, n, m, xyz1, xyz2, factors, match, (float *)workspace));
    return PNAE_OK;
}
