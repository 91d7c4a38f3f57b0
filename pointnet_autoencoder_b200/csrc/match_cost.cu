// match_cost.cu -- EMD transport cost and its gradients for sm_100a.
//
// Replaces matchcost / matchcostgrad1 / matchcostgrad2 and their launchers
// (reference: tf_ops/approxmatch/tf_approxmatch_g.cu:183-295).
//
// Two families:
//  * factor path (pnae_match_cost_factors): cost, grad1 and grad2 in ONE pass over
//    the pairs, re-evaluating match[l,k] = sum_j E_j(k,l) ratioL_j[k] ratioR_j[l]
//    on the fly from the per-level factors, so the (b,m,n) tensor never exists;
//  * dense path (pnae_match_cost_fwd / _bwd, pnae_match_from_factors): the
//    reference's own signatures over a materialised (b,m,n) match, for callers
//    that hold one.
#include "pnae_common.cuh"

namespace {

constexpr int kL = PNAE_NUM_LEVELS;

// ---------------------------------------------------------------------------
// factor path
// ---------------------------------------------------------------------------
constexpr int kFThreads = 128;
constexpr int kFR = 2;        // dataset rows per thread
constexpr int kFTile = 128;   // query columns per shared-memory tile

struct __align__(16) ColRec {   // one streamed query point: coordinates + its 10 ratioR_j
    float x, y, z, pad;
    float rr[12];
};

template <bool WITH_GRAD>
__global__ void __launch_bounds__(kFThreads)
match_cost_factors_kernel(int n, int m, const float *__restrict__ xyz1, const float *__restrict__ xyz2,
                          const float *__restrict__ factors, float *__restrict__ cost,
                          float *__restrict__ grad1, float *__restrict__ grad2, int nblk)
{
    __shared__ ColRec tile[kFTile];
    __shared__ float red[kFThreads / 32];

    const int i = blockIdx.x / nblk;
    const int rb = blockIdx.x - i * nblk;
    const float *p1 = xyz1 + (size_t)i * n * 3;
    const float *p2 = xyz2 + (size_t)i * m * 3;
    const float *fac = factors + (size_t)i * kL * (n + m);

    float x1[kFR], y1[kFR], z1[kFR], rl[kFR][kL];
    float gx[kFR], gy[kFR], gz[kFR];
    bool live[kFR];
    int row[kFR];
    float csum = 0.f;
#pragma unroll
    for (int r = 0; r < kFR; r++) {
        row[r] = rb * (kFThreads * kFR) + r * kFThreads + (int)threadIdx.x;
        live[r] = row[r] < n;
        const int k = min(row[r], n - 1);
        x1[r] = __ldg(p1 + k * 3); y1[r] = __ldg(p1 + k * 3 + 1); z1[r] = __ldg(p1 + k * 3 + 2);
#pragma unroll
        for (int j = 0; j < kL; j++) rl[r][j] = live[r] ? __ldg(fac + (size_t)j * (n + m) + k) : 0.f;
        gx[r] = gy[r] = gz[r] = 0.f;
    }

    const int lane = threadIdx.x & 31;
    for (int l0 = 0; l0 < m; l0 += kFTile) {
        const int cnt = min(kFTile, m - l0);
        __syncthreads();
        for (int t = threadIdx.x; t < cnt; t += kFThreads) {
            const int l = l0 + t;
            ColRec c;
            c.x = __ldg(p2 + l * 3); c.y = __ldg(p2 + l * 3 + 1); c.z = __ldg(p2 + l * 3 + 2); c.pad = 0.f;
#pragma unroll
            for (int j = 0; j < kL; j++) c.rr[j] = __ldg(fac + (size_t)j * (n + m) + n + l);
            c.rr[10] = c.rr[11] = 0.f;
            tile[t] = c;
        }
        __syncthreads();
        for (int t = 0; t < cnt; t++) {
            const float4 p = *reinterpret_cast<const float4 *>(&tile[t].x);
            const float4 ra = *reinterpret_cast<const float4 *>(&tile[t].rr[0]);
            const float4 rb4 = *reinterpret_cast<const float4 *>(&tile[t].rr[4]);
            const float4 rc = *reinterpret_cast<const float4 *>(&tile[t].rr[8]);
            const float rr[kL] = {ra.x, ra.y, ra.z, ra.w, rb4.x, rb4.y, rb4.z, rb4.w, rc.x, rc.y};
            float g2x = 0.f, g2y = 0.f, g2z = 0.f;
#pragma unroll
            for (int r = 0; r < kFR; r++) {
                const float dx = x1[r] - p.x, dy = y1[r] - p.y, dz = z1[r] - p.z;
                const float d = pnae_sqdist(dx, dy, dz);
                float mv = 0.f;    // match[l,k], accumulated in level order like `match+=w` (:152)
#pragma unroll
                for (int j = 0; j < kL; j++) {
                    const float e = pnae_ex2(__fmul_rn(d, pnae_level_scale(j)));
                    mv = __fmaf_rn(__fmul_rn(e, rl[r][j]), rr[j], mv);
                }
                csum = __fmaf_rn(__fsqrt_rn(d), mv, csum);                      // :207-208
                if (WITH_GRAD) {
                    const float w = __fmul_rn(mv, pnae_rsqrt(fmaxf(d, 1e-20f)));   // :243, :281
                    gx[r] = __fmaf_rn(dx, w, gx[r]); gy[r] = __fmaf_rn(dy, w, gy[r]); gz[r] = __fmaf_rn(dz, w, gz[r]);
                    g2x = __fmaf_rn(-dx, w, g2x); g2y = __fmaf_rn(-dy, w, g2y); g2z = __fmaf_rn(-dz, w, g2z);
                }
            }
            if (WITH_GRAD) {
                g2x = warp_sum(g2x); g2y = warp_sum(g2y); g2z = warp_sum(g2z);
                if (lane == 0) {
                    float *g = grad2 + ((size_t)i * m + l0 + t) * 3;
                    atomicAdd(g, g2x); atomicAdd(g + 1, g2y); atomicAdd(g + 2, g2z);
                }
            }
        }
    }
    if (WITH_GRAD) {
#pragma unroll
        for (int r = 0; r < kFR; r++)
            if (live[r]) {
                float *g = grad1 + ((size_t)i * n + row[r]) * 3;
                g[0] = gx[r]; g[1] = gy[r]; g[2] = gz[r];
            }
    }
    csum = warp_sum(csum);
    if (lane == 0) red[threadIdx.x >> 5] = csum;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kFThreads / 32; w++) s += red[w];
        atomicAdd(cost + i, s);
    }
}

// ---------------------------------------------------------------------------
// dense path
// ---------------------------------------------------------------------------
constexpr int kDThreads = 256;
constexpr int kDRows = 64;   // query rows l per CTA in the k-parallel kernels

// match[i,l,k] from the factors; thread owns k (coalesced stores), CTA owns kDRows rows l.
__global__ void __launch_bounds__(kDThreads)
match_from_factors_kernel(int n, int m, const float *__restrict__ xyz1, const float *__restrict__ xyz2,
                          const float *__restrict__ factors, float *__restrict__ match, int nkb, int nlb)
{
    __shared__ ColRec tile[kDRows];
    int bid = blockIdx.x;
    const int kb = bid % nkb; bid /= nkb;
    const int lb = bid % nlb;
    const int i = bid / nlb;
    const float *p1 = xyz1 + (size_t)i * n * 3;
    const float *p2 = xyz2 + (size_t)i * m * 3;
    const float *fac = factors + (size_t)i * kL * (n + m);
    const int l0 = lb * kDRows, cnt = min(kDRows, m - l0);
    for (int t = threadIdx.x; t < cnt; t += kDThreads) {
        const int l = l0 + t;
        ColRec c;
        c.x = __ldg(p2 + l * 3); c.y = __ldg(p2 + l * 3 + 1); c.z = __ldg(p2 + l * 3 + 2); c.pad = 0.f;
#pragma unroll
        for (int j = 0; j < kL; j++) c.rr[j] = __ldg(fac + (size_t)j * (n + m) + n + l);
        c.rr[10] = c.rr[11] = 0.f;
        tile[t] = c;
    }
    __syncthreads();
    const int k = kb * kDThreads + threadIdx.x;
    if (k >= n) return;
    const float x1 = __ldg(p1 + k * 3), y1 = __ldg(p1 + k * 3 + 1), z1 = __ldg(p1 + k * 3 + 2);
    float rl[kL];
#pragma unroll
    for (int j = 0; j < kL; j++) rl[j] = __ldg(fac + (size_t)j * (n + m) + k);
    float *out = match + ((size_t)i * m + l0) * n + k;
    for (int t = 0; t < cnt; t++) {
        const float d = pnae_sqdist(tile[t].x - x1, tile[t].y - y1, tile[t].z - z1);
        float mv = 0.f;
#pragma unroll
        for (int j = 0; j < kL; j++)
            mv = __fmaf_rn(__fmul_rn(pnae_ex2(__fmul_rn(d, pnae_level_scale(j))), rl[j]), tile[t].rr[j], mv);
        __stcs(out + (size_t)t * n, mv);
    }
}

// cost[i] += sum over this CTA's (l-block, k-block) of sqrtf(d) * match ; also grad1 partials.
// Thread owns k: the match reads are coalesced along k.
template <bool GRAD>
__global__ void __launch_bounds__(kDThreads)
dense_rows_kernel(int n, int m, const float *__restrict__ xyz1, const float *__restrict__ xyz2,
                  const float *__restrict__ match, float *__restrict__ cost, float *__restrict__ grad1,
                  int nkb, int nlb)
{
    __shared__ float4 tile[kDRows];
    __shared__ float red[kDThreads / 32];
    int bid = blockIdx.x;
    const int kb = bid % nkb; bid /= nkb;
    const int lb = bid % nlb;
    const int i = bid / nlb;
    const float *p1 = xyz1 + (size_t)i * n * 3;
    const float *p2 = xyz2 + (size_t)i * m * 3;
    const int l0 = lb * kDRows, cnt = min(kDRows, m - l0);
    for (int t = threadIdx.x; t < cnt; t += kDThreads)
        tile[t] = make_float4(__ldg(p2 + (l0 + t) * 3), __ldg(p2 + (l0 + t) * 3 + 1), __ldg(p2 + (l0 + t) * 3 + 2), 0.f);
    __syncthreads();
    const int k = kb * kDThreads + threadIdx.x;
    float s = 0.f, gx = 0.f, gy = 0.f, gz = 0.f;
    if (k < n) {
        const float x1 = __ldg(p1 + k * 3), y1 = __ldg(p1 + k * 3 + 1), z1 = __ldg(p1 + k * 3 + 2);
        const float *mp = match + ((size_t)i * m + l0) * n + k;
#pragma unroll 4
        for (int t = 0; t < cnt; t++) {
            const float mv = __ldcs(mp + (size_t)t * n);
            const float dx = x1 - tile[t].x, dy = y1 - tile[t].y, dz = z1 - tile[t].z;
            const float d = pnae_sqdist(dx, dy, dz);
            if (!GRAD) {
                s = __fmaf_rn(__fsqrt_rn(d), mv, s);
            } else {
                const float w = __fmul_rn(mv, pnae_rsqrt(fmaxf(d, 1e-20f)));
                gx = __fmaf_rn(dx, w, gx); gy = __fmaf_rn(dy, w, gy); gz = __fmaf_rn(dz, w, gz);
            }
        }
    }
    if (!GRAD) {
        s = warp_sum(s);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x == 0) {
            float tsum = 0.f;
#pragma unroll
            for (int w = 0; w < kDThreads / 32; w++) tsum += red[w];
            atomicAdd(cost + i, tsum);
        }
    } else if (k < n) {
        float *g = grad1 + ((size_t)i * n + k) * 3;
        atomicAdd(g, gx); atomicAdd(g + 1, gy); atomicAdd(g + 2, gz);
    }
}

// grad2[i,l] = sum_k match[i,l,k] (x2_l - x1_k) rsqrt(max(d,1e-20)); one warp per row l,
// lanes stride along k (the contiguous axis of match).
__global__ void __launch_bounds__(kDThreads)
dense_grad2_kernel(int b, int n, int m, const float *__restrict__ xyz1, const float *__restrict__ xyz2,
                   const float *__restrict__ match, float *__restrict__ grad2)
{
    const int lane = threadIdx.x & 31;
    const long long rows = (long long)b * m;
    for (long long row = (long long)blockIdx.x * (kDThreads / 32) + (threadIdx.x >> 5); row < rows;
         row += (long long)gridDim.x * (kDThreads / 32)) {
        const int i = (int)(row / m);
        const float *p1 = xyz1 + (size_t)i * n * 3;
        const float *q = xyz2 + (size_t)row * 3;
        const float x2 = __ldg(q), y2 = __ldg(q + 1), z2 = __ldg(q + 2);
        const float *mp = match + (size_t)row * n;
        float gx = 0.f, gy = 0.f, gz = 0.f;
        for (int k = lane; k < n; k += 32) {
            const float dx = x2 - __ldg(p1 + k * 3), dy = y2 - __ldg(p1 + k * 3 + 1), dz = z2 - __ldg(p1 + k * 3 + 2);
            const float w = __fmul_rn(__ldcs(mp + k), pnae_rsqrt(fmaxf(pnae_sqdist(dx, dy, dz), 1e-20f)));
            gx = __fmaf_rn(dx, w, gx); gy = __fmaf_rn(dy, w, gy); gz = __fmaf_rn(dz, w, gz);
        }
        gx = warp_sum(gx); gy = warp_sum(gy); gz = warp_sum(gz);
        if (lane == 0) {
            float *g = grad2 + (size_t)row * 3;
            g[0] = gx; g[1] = gy; g[2] = gz;
        }
    }
}

int check_common(const char *op, int b, int n, int m, const void *xyz1, const void *xyz2)
{
    PNAE_REQUIRE(b >= 0 && n >= 1 && m >= 1, "%s: need b>=0, n>=1, m>=1 (got b=%d n=%d m=%d)", op, b, n, m);
    PNAE_REQUIRE(xyz1 && xyz2, "%s: NULL pointer", op);
    return PNAE_OK;
}

}  // namespace

extern "C" int pnae_match_cost_factors(int b, int n, int m, const float *xyz1, const float *xyz2,
                                       const float *factors, float *cost, float *grad1, float *grad2,
                                       void *stream)
{
    int rc = check_common("match_cost_factors", b, n, m, xyz1, xyz2);
    if (rc) return rc;
    PNAE_REQUIRE(factors && cost, "match_cost_factors: NULL pointer");
    PNAE_REQUIRE((grad1 == nullptr) == (grad2 == nullptr), "match_cost_factors: pass both gradients or neither");
    if (b == 0) return PNAE_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int nblk = (n + kFThreads * kFR - 1) / (kFThreads * kFR);
    PNAE_CUDA_OK(cudaMemsetAsync(cost, 0, sizeof(float) * (size_t)b, st));
    if (grad1) {
        PNAE_CUDA_OK(cudaMemsetAsync(grad2, 0, sizeof(float) * (size_t)b * m * 3, st));
        match_cost_factors_kernel<true><<<(unsigned)(b * nblk), kFThreads, 0, st>>>(n, m, xyz1, xyz2, factors, cost, grad1, grad2, nblk);
    } else {
        match_cost_factors_kernel<false><<<(unsigned)(b * nblk), kFThreads, 0, st>>>(n, m, xyz1, xyz2, factors, cost, nullptr, nullptr, nblk);
    }
    PNAE_CUDA_OK(cudaGetLastError());
    return PNAE_OK;
}

int pnae_match_from_factors_impl(int b, int n, int m, const float *xyz1, const float *xyz2,
                                 const float *factors, float *match, cudaStream_t st)
{
    const int nkb = (n + kDThreads - 1) / kDThreads, nlb = (m + kDRows - 1) / kDRows;
    const long long grid = (long long)b * nkb * nlb;
    PNAE_REQUIRE(grid < (1ll << 31), "match_from_factors: problem too large for one launch");
    match_from_factors_kernel<<<(unsigned)grid, kDThreads, 0, st>>>(n, m, xyz1, xyz2, factors, match, nkb, nlb);
    PNAE_CUDA_OK(cudaGetLastError());
    return PNAE_OK;
}

extern "C" int pnae_match_from_factors(int b, int n, int m, const float *xyz1, const float *xyz2,
                                       const float *factors, float *match, void *stream)
{
    int rc = check_common("match_from_factors", b, n, m, xyz1, xyz2);
    if (rc) return rc;
    PNAE_REQUIRE(factors && match, "match_from_factors: NULL pointer");
    if (b == 0) return PNAE_OK;
    return pnae_match_from_factors_impl(b, n, m, xyz1, xyz2, factors, match, (cudaStream_t)stream);
}

extern "C" int pnae_match_cost_fwd(int b, int n, int m, const float *xyz1, const float *xyz2,
                                   const float *match, float *cost, void *stream)
{
    int rc = check_common("match_cost", b, n, m, xyz1, xyz2);
    if (rc) return rc;
    PNAE_REQUIRE(match && cost, "match_cost: NULL pointer");
    if (b == 0) return PNAE_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int nkb = (n + kDThreads - 1) / kDThreads, nlb = (m + kDRows - 1) / kDRows;
    const long long grid = (long long)b * nkb * nlb;
    PNAE_REQUIRE(grid < (1ll << 31), "match_cost: problem too large for one launch");
    PNAE_CUDA_OK(cudaMemsetAsync(cost, 0, sizeof(float) * (size_t)b, st));
    dense_rows_kernel<false><<<(unsigned)grid, kDThreads, 0, st>>>(n, m, xyz1, xyz2, match, cost, nullptr, nkb, nlb);
    PNAE_CUDA_OK(cudaGetLastError());
    return PNAE_OK;
}

extern "C" int pnae_match_cost_bwd(int b, int n, int m, const float *xyz1, const float *xyz2,
                                   const float *match, float *grad1, float *grad2, void *stream)
{
    int rc = check_common("match_cost_grad", b, n, m, xyz1, xyz2);
    if (rc) return rc;
    PNAE_REQUIRE(match && grad1 && grad2, "match_cost_grad: NULL pointer");
    if (b == 0) return PNAE_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int nkb = (n + kDThreads - 1) / kDThreads, nlb = (m + kDRows - 1) / kDRows;
    const long long grid = (long long)b * nkb * nlb;
    PNAE_REQUIRE(grid < (1ll << 31), "match_cost_grad: problem too large for one launch");
    PNAE_CUDA_OK(cudaMemsetAsync(grad1, 0, sizeof(float) * (size_t)b * n * 3, st));
    dense_rows_kernel<true><<<(unsigned)grid, kDThreads, 0, st>>>(n, m, xyz1, xyz2, match, nullptr, grad1, nkb, nlb);
    PNAE_CUDA_OK(cudaGetLastError());
    const long long rows = (long long)b * m;
    const int g2 = (int)min((rows + kDThreads / 32 - 1) / (kDThreads / 32), (long long)pnae_sm_count() * 16);
    dense_grad2_kernel<<<g2, kDThreads, 0, st>>>(b, n, m, xyz1, xyz2, match, grad2);
    PNAE_CUDA_OK(cudaGetLastError());
    return PNAE_OK;
}
