// nn_distance.cu -- Chamfer distance forward + gradient for sm_100a.
//
// Replaces NmDistanceKernel / NmDistanceGradKernel and their launchers
// (reference: tf_ops/nn_distance/tf_nndistance_g.cu:5-157).  Semantics kept:
// squared distances with the FMUL(y)->FFMA(x)->FFMA(z) rounding of the reference
// SASS, strict '<' so the lowest index wins ties, gradient outputs zeroed inside.
#include "pnae_common.cuh"

namespace {

constexpr int kFwdThreads = 128;
constexpr int kFwdTile = 1024;          // candidate points per shared-memory tile (16 KB as float4)

// One CTA = one (batch element, direction, block of kFwdThreads*R query points).
// Both directions run in the same launch; candidates are staged once per tile as
// float4 so the inner loop is one broadcast LDS.128 per candidate for R queries.
template <int R>
__global__ void __launch_bounds__(kFwdThreads)
nn_fwd_kernel(int n, const float *__restrict__ xyz1, int m, const float *__restrict__ xyz2,
              float *__restrict__ dist1, int *__restrict__ idx1,
              float *__restrict__ dist2, int *__restrict__ idx2, int nb1, int nb2)
{
    __shared__ float4 tile[kFwdTile];

    const int per_batch = nb1 + nb2;
    const int i = blockIdx.x / per_batch;
    int rb = blockIdx.x - i * per_batch;
    // direction 1: queries = xyz1, candidates = xyz2 ; direction 2 swapped
    const bool dir2 = rb >= nb1;
    if (dir2) rb -= nb1;
    const int nq = dir2 ? m : n;
    const int nc = dir2 ? n : m;
    const float *q = (dir2 ? xyz2 : xyz1) + (size_t)i * nq * 3;
    const float *c = (dir2 ? xyz1 : xyz2) + (size_t)i * nc * 3;
    float *dist = (dir2 ? dist2 : dist1) + (size_t)i * nq;
    int *idx = (dir2 ? idx2 : idx1) + (size_t)i * nq;

    float qx[R], qy[R], qz[R], best[R];
    int besti[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
        int j = min(rb * (kFwdThreads * R) + r * kFwdThreads + (int)threadIdx.x, nq - 1);
        qx[r] = __ldg(q + j * 3 + 0);
        qy[r] = __ldg(q + j * 3 + 1);
        qz[r] = __ldg(q + j * 3 + 2);
        best[r] = __int_as_float(0x7f800000);   // +inf: the first candidate always wins, like `k==0 ||`
        besti[r] = 0;
    }

    for (int k0 = 0; k0 < nc; k0 += kFwdTile) {
        const int cnt = min(kFwdTile, nc - k0);
        __syncthreads();
        for (int t = threadIdx.x; t < cnt; t += kFwdThreads) {
            const float *p = c + (size_t)(k0 + t) * 3;
            tile[t] = make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), 0.f);
        }
        __syncthreads();
#pragma unroll 4
        for (int k = 0; k < cnt; k++) {
            const float4 p = tile[k];
#pragma unroll
            for (int r = 0; r < R; r++) {
                const float d = pnae_sqdist(p.x - qx[r], p.y - qy[r], p.z - qz[r]);
                if (d < best[r]) {
                    best[r] = d;
                    besti[r] = k0 + k;
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < R; r++) {
        int j = rb * (kFwdThreads * R) + r * kFwdThreads + (int)threadIdx.x;
        if (j < nq) {
            dist[j] = best[r];
            idx[j] = besti[r];
        }
    }
}

// Gradient: one thread per point of either cloud.
//   grad_a[j]      += 2 g (a_j - c_idx)      (own point)
//   grad_c[idx[j]] -= 2 g (a_j - c_idx)      (scatter, float atomics like the reference)
// Outputs are zeroed by the launcher (tf_nndistance_g.cu:153-154).
__global__ void __launch_bounds__(256)
nn_bwd_kernel(int b, int n, const float *__restrict__ xyz1, int m, const float *__restrict__ xyz2,
              const float *__restrict__ grad_dist1, const int *__restrict__ idx1,
              const float *__restrict__ grad_dist2, const int *__restrict__ idx2,
              float *__restrict__ grad_xyz1, float *__restrict__ grad_xyz2)
{
    const long long total1 = (long long)b * n, total = total1 + (long long)b * m;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const bool second = t >= total1;
        const long long u = second ? t - total1 : t;
        const int na = second ? m : n, nc = second ? n : m;
        const int i = (int)(u / na), j = (int)(u - (long long)i * na);
        const float *a = (second ? xyz2 : xyz1) + ((size_t)i * na + j) * 3;
        const int j2 = (second ? idx2 : idx1)[u];
        const float *c = (second ? xyz1 : xyz2) + ((size_t)i * nc + j2) * 3;
        float *ga = (second ? grad_xyz2 : grad_xyz1) + ((size_t)i * na + j) * 3;
        float *gc = (second ? grad_xyz1 : grad_xyz2) + ((size_t)i * nc + j2) * 3;
        const float g = __fmul_rn((second ? grad_dist2 : grad_dist1)[u], 2.0f);
#pragma unroll
        for (int ax = 0; ax < 3; ax++) {
            const float v = __fmul_rn(g, __fsub_rn(__ldg(a + ax), __ldg(c + ax)));
            atomicAdd(ga + ax, v);
            atomicAdd(gc + ax, -v);
        }
    }
}

}  // namespace

extern "C" size_t pnae_nn_distance_workspace_bytes(int b, int n, int m)
{
    (void)b; (void)n; (void)m;
    return 0;
}

extern "C" int pnae_nn_distance_fwd(int b, int n, const float *xyz1, int m, const float *xyz2,
                                    float *dist1, int *idx1, float *dist2, int *idx2,
                                    void *workspace, size_t workspace_bytes, void *stream)
{
    (void)workspace; (void)workspace_bytes;
    PNAE_REQUIRE(b >= 0 && n >= 1 && m >= 1, "nn_distance: need b>=0, n>=1, m>=1 (got b=%d n=%d m=%d)", b, n, m);
    PNAE_REQUIRE(xyz1 && xyz2 && dist1 && idx1 && dist2 && idx2, "nn_distance: NULL pointer");
    if (b == 0) return PNAE_OK;
    constexpr int R = 2;
    const int nb1 = (n + kFwdThreads * R - 1) / (kFwdThreads * R);
    const int nb2 = (m + kFwdThreads * R - 1) / (kFwdThreads * R);
    const long long grid = (long long)b * (nb1 + nb2);
    PNAE_REQUIRE(grid < (1ll << 31), "nn_distance: problem too large for one launch");
    nn_fwd_kernel<R><<<(unsigned)grid, kFwdThreads, 0, (cudaStream_t)stream>>>(n, xyz1, m, xyz2, dist1, idx1, dist2, idx2, nb1, nb2);
    PNAE_CUDA_OK(cudaGetLastError());
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
    cudaStream_t st = (cudaStream_t)stream;
    PNAE_CUDA_OK(cudaMemsetAsync(grad_xyz1, 0, sizeof(float) * (size_t)b * n * 3, st));
    PNAE_CUDA_OK(cudaMemsetAsync(grad_xyz2, 0, sizeof(float) * (size_t)b * m * 3, st));
    const long long total = (long long)b * n + (long long)b * m;
    const int grid = (int)min((total + 255) / 256, (long long)pnae_sm_count() * 8);
    nn_bwd_kernel<<<grid, 256, 0, st>>>(b, n, xyz1, m, xyz2, grad_dist1, idx1, grad_dist2, idx2, grad_xyz1, grad_xyz2);
    PNAE_CUDA_OK(cudaGetLastError());
    return PNAE_OK;
}
