// microbench_nn.cu -- the Chamfer sweep's inner loop in isolation (8 rows per lane in registers, columns as float4
// records broadcast from shared memory, row minima by FMNMX3, an 8-row column tree), in the arithmetic variants the
// kernel could use.  Prints issue cycles per (row, column) pair per SM sub-partition at the kernel's occupancy
// (4 CTAs x 4 warps per SM).   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_nn microbench_nn.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ float min3f(float a, float b, float c) { float r; asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }

constexpr int kR = 8, kCols = 32;

// MODE 0: exact (3 FADD, FMUL, 2 FFMA) + row min3 + column tree       (round-1 kernel)
// MODE 1: filter (FADD, 3 FFMA)        + row min3 + column tree       (round-2 kernel)
// MODE 2: filter arithmetic only (sum kept alive with one FADD per pair instead of minima)
// MODE 3: exact arithmetic only
// MODE 4: filter + row min3 only (no column tree)
// MODE 5: filter with 2-input minima everywhere
// MODE 6: filter, B as the chain's addend (3 FFMA) and the norm added for the column side only (FADD after the chain)
template <int MODE>
__global__ void __launch_bounds__(128, 4) kern(float *out, const float4 *cols, int iters)
{
    __shared__ float4 sc[4][kCols + 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    sc[warp][lane] = cols[lane];
    if (lane == 0) sc[warp][kCols] = cols[0];
    __syncwarp();
    float rx[kR], ry[kR], rz[kR], ra[kR], best[kR];
#pragma unroll
    for (int r = 0; r < kR; r++) { rx[r] = lane * 0.01f + r; ry[r] = lane * 0.02f - r; rz[r] = 0.5f * r + lane; ra[r] = rx[r] * rx[r] + ry[r] * ry[r] + rz[r] * rz[r]; best[r] = 3e38f; }
    float colacc = 3e38f;
    for (int it = 0; it < iters; it++) {
        float4 qn = sc[warp][0];
#pragma unroll 2
        for (int c = 0; c < kCols; c += 2) {
            const float4 q0 = qn, q1 = sc[warp][c + 1];
            qn = sc[warp][c + 2];
            float d0[kR], d1[kR];
#pragma unroll
            for (int r = 0; r < kR; r++) {
                if (MODE == 0 || MODE == 3) {
                    float dx = q0.x - rx[r], dy = q0.y - ry[r], dz = q0.z - rz[r];
                    d0[r] = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
                    dx = q1.x - rx[r]; dy = q1.y - ry[r]; dz = q1.z - rz[r];
                    d1[r] = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
                } else if (MODE == 6) {
                    d0[r] = __fmaf_rn(rz[r], q0.z, __fmaf_rn(ry[r], q0.y, __fmaf_rn(rx[r], q0.x, q0.w)));
                    d1[r] = __fmaf_rn(rz[r], q1.z, __fmaf_rn(ry[r], q1.y, __fmaf_rn(rx[r], q1.x, q1.w)));
                } else {
                    d0[r] = __fmaf_rn(rz[r], q0.z, __fmaf_rn(ry[r], q0.y, __fmaf_rn(rx[r], q0.x, __fadd_rn(ra[r], q0.w))));
                    d1[r] = __fmaf_rn(rz[r], q1.z, __fmaf_rn(ry[r], q1.y, __fmaf_rn(rx[r], q1.x, __fadd_rn(ra[r], q1.w))));
                }
                if (MODE == 2 || MODE == 3) best[r] = __fadd_rn(best[r], __fadd_rn(d0[r], d1[r]));
                else if (MODE == 5) best[r] = fminf(fminf(best[r], d0[r]), d1[r]);
                else best[r] = min3f(best[r], d0[r], d1[r]);
            }
            if (MODE == 0 || MODE == 1) {
                float m0 = fminf(min3f(min3f(min3f(d0[0], d0[1], d0[2]), d0[3], d0[4]), d0[5], d0[6]), d0[7]);
                float m1 = fminf(min3f(min3f(min3f(d1[0], d1[1], d1[2]), d1[3], d1[4]), d1[5], d1[6]), d1[7]);
                colacc = min3f(colacc, m0, m1);
            }
            if (MODE == 5) {
                float m0 = d0[0], m1 = d1[0];
#pragma unroll
                for (int r = 1; r < kR; r++) { m0 = fminf(m0, d0[r]); m1 = fminf(m1, d1[r]); }
                colacc = fminf(fminf(colacc, m0), m1);
            }
            if (MODE == 6) {
                float e0[kR], e1[kR];
#pragma unroll
                for (int r = 0; r < kR; r++) { e0[r] = __fadd_rn(d0[r], ra[r]); e1[r] = __fadd_rn(d1[r], ra[r]); }
                float m0 = fminf(min3f(min3f(min3f(e0[0], e0[1], e0[2]), e0[3], e0[4]), e0[5], e0[6]), e0[7]);
                float m1 = fminf(min3f(min3f(min3f(e1[0], e1[1], e1[2]), e1[3], e1[4]), e1[5], e1[6]), e1[7]);
                colacc = min3f(colacc, m0, m1);
            }
        }
    }
    float s = colacc;
#pragma unroll
    for (int r = 0; r < kR; r++) s += best[r];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char *name, int sms, float *out, const float4 *cols)
{
    const int blocks = sms * 4, threads = 128, iters = 2048;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    kern<MODE><<<blocks, threads>>>(out, cols, 16);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        CK(cudaEventRecord(e0));
        kern<MODE><<<blocks, threads>>>(out, cols, iters);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    // per SM sub-partition: 4 warps, each iters * kCols * kR pairs per lane; cycles = time * clock
    const double cycles = best * 1e-3 * clk_khz * 1e3;
    const double pairs_per_lane_per_smsp = 4.0 * iters * kCols * kR;
    printf("%-58s %8.3f ms  %6.2f issue cycles per pair (per lane, per sub-partition)\n", name, best, cycles / pairs_per_lane_per_smsp);
}

int main()
{
    int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    float *out; CK(cudaMalloc(&out, sizeof(float) * sms * 4 * 128));
    float4 h[32];
    for (int i = 0; i < 32; i++) h[i] = make_float4(0.1f * i, 0.2f * i, -0.05f * i, 0.01f * i * i);
    float4 *cols; CK(cudaMalloc(&cols, sizeof(h))); CK(cudaMemcpy(cols, h, sizeof(h), cudaMemcpyHostToDevice));
    run<3>("exact arithmetic only (3 FADD, FMUL, 2 FFMA)", sms, out, cols);
    run<2>("filter arithmetic only (FADD, 3 FFMA)", sms, out, cols);
    run<0>("exact + row min3 + column tree (round 1)", sms, out, cols);
    run<1>("filter + row min3 + column tree (round 2)", sms, out, cols);
    run<4>("filter + row min3 only", sms, out, cols);
    run<5>("filter + 2-input minima", sms, out, cols);
    run<6>("filter, B as addend; norm added for the column side", sms, out, cols);
    return 0;
}
