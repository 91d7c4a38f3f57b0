// microbench.cu -- per-SM issue rates of the instructions the Chamfer / EMD inner
// loops are made of, measured on the box (the hardware guides were measured on
// B300; this pins the B200 numbers the rooflines in DESIGN.md use).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu
// Prints one line per test: ops per clock per SM (thread-level ops).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#define ITERS 4096
#define NACC 8

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int MODE>
__global__ void __launch_bounds__(256) kern(float *out, float a, float b, int iters)
{
    float acc[NACC];
    float2 acc2[NACC];
    unsigned uacc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; i++) { acc[i] = threadIdx.x * 1e-3f + i; acc2[i] = make_float2(acc[i], acc[i] + 1.f); uacc[i] = threadIdx.x + i; }
    float2 a2 = make_float2(a, a * 1.0001f), b2 = make_float2(b, b * 0.999f);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) {
            if (MODE == 0) acc[i] = __fmaf_rn(acc[i], a, b);                          // FFMA
            if (MODE == 1) acc2[i] = __ffma2_rn(acc2[i], a2, b2);                     // FFMA2
            if (MODE == 2) acc[i] = fminf(acc[i], a + i);                             // FMNMX
            if (MODE == 3) { float r; asm volatile("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(acc[i]), "f"(a), "f"(b)); acc[i] = r; } // FMNMX3
            if (MODE == 4) { float r; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(acc[i])); acc[i] = r; }               // MUFU.EX2
            if (MODE == 5) acc[i] = __fadd_rn(acc[i], a);                             // FADD
            if (MODE == 6) acc2[i] = __fadd2_rn(acc2[i], a2);                         // FADD2
            if (MODE == 7) { acc[i] = __fmaf_rn(acc[i], a, b); float r; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(acc[i])); acc[i] = r; } // 1 FFMA + 1 MUFU
            if (MODE == 8) { uacc[i] = __reduce_min_sync(0xffffffffu, uacc[i]) + i; } // REDUX
            if (MODE == 9) { acc[i] = __shfl_xor_sync(0xffffffffu, acc[i], 1); }      // SHFL
            if (MODE == 10) {  // 4 FFMA + 1 MUFU
                acc[i] = __fmaf_rn(acc[i], a, b); acc[i] = __fmaf_rn(acc[i], a, b); acc[i] = __fmaf_rn(acc[i], a, b); acc[i] = __fmaf_rn(acc[i], a, b);
                float r; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(acc[i])); acc[i] = r; }
            if (MODE == 11) {  // 8 FFMA + 1 MUFU
#pragma unroll
                for (int q = 0; q < 8; q++) acc[i] = __fmaf_rn(acc[i], a, b);
                float r; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(acc[i])); acc[i] = r; }
            if (MODE == 12) {  // 4 FFMA2 + 2 MUFU  (packed math feeding two exps)
#pragma unroll
                for (int q = 0; q < 4; q++) acc2[i] = __ffma2_rn(acc2[i], a2, b2);
                float r0, r1; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(acc2[i].x)); asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(acc2[i].y));
                acc2[i] = make_float2(r0, r1); }
            if (MODE == 13) {  // 3 FFMA2 + 1 FMNMX3 (the Chamfer mix per 2 pairs, row side only)
#pragma unroll
                for (int q = 0; q < 3; q++) acc2[i] = __ffma2_rn(acc2[i], a2, b2);
                float r; asm volatile("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(acc[i]), "f"(acc2[i].x), "f"(acc2[i].y)); acc[i] = r; }
            if (MODE == 14) {  // 6 FFMA + 1 FMNMX3
#pragma unroll
                for (int q = 0; q < 6; q++) acc[i] = __fmaf_rn(acc[i], a, b);
                float r; asm volatile("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(acc2[i].x), "f"(acc[i]), "f"(b)); acc2[i].x = r; }
            if (MODE == 15) {  // 3 FFMA2 + 2 FMNMX3 (both-direction Chamfer mix per 2 pairs)
#pragma unroll
                for (int q = 0; q < 3; q++) acc2[i] = __ffma2_rn(acc2[i], a2, b2);
                float r, s; asm volatile("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(acc[i]), "f"(acc2[i].x), "f"(acc2[i].y));
                asm volatile("min.f32 %0, %1, %2, %3;" : "=f"(s) : "f"(a), "f"(acc2[i].x), "f"(acc2[i].y)); acc[i] = r + s * 0.f; }
            if (MODE == 16) { float r; asm volatile("min.f32 %0, %1, %2;" : "=f"(r) : "f"(acc[i]), "f"(a)); acc[i] = r; }   // FMNMX
            if (MODE == 17 || MODE == 18) {  // 6 FFMA + 1|2 FMNMX
#pragma unroll
                for (int q = 0; q < 6; q++) acc[i] = __fmaf_rn(acc[i], a, b);
                float r; asm volatile("min.f32 %0, %1, %2;" : "=f"(r) : "f"(acc2[i].x), "f"(acc[i])); acc2[i].x = r;
                if (MODE == 18) { asm volatile("min.f32 %0, %1, %2;" : "=f"(r) : "f"(acc2[i].y), "f"(acc[i])); acc2[i].y = r; } }
            if (MODE == 19) {  // 6 FFMA + compare/select pair (the reference's argmin update)
#pragma unroll
                for (int q = 0; q < 6; q++) acc[i] = __fmaf_rn(acc[i], a, b);
                if (acc[i] < acc2[i].x) { acc2[i].x = acc[i]; uacc[i] = it; } }
            if (MODE == 20) {  // 6 FFMA + 1 IADD3 (does the ALU pipe co-issue for free?)
#pragma unroll
                for (int q = 0; q < 6; q++) acc[i] = __fmaf_rn(acc[i], a, b);
                unsigned r; asm volatile("add.u32 %0, %1, %2;" : "=r"(r) : "r"(uacc[i]), "r"(it)); uacc[i] = r; }
            if (MODE == 21) {  // 6 FFMA + 2 integer min (d >= 0: u32 order == float order)
#pragma unroll
                for (int q = 0; q < 6; q++) acc[i] = __fmaf_rn(acc[i], a, b);
                unsigned r; asm volatile("min.u32 %0, %1, %2;" : "=r"(r) : "r"(uacc[i]), "r"(__float_as_uint(acc[i]))); uacc[i] = r;
                unsigned q2; asm volatile("min.u32 %0, %1, %2;" : "=r"(q2) : "r"(__float_as_uint(acc2[i].y)), "r"(__float_as_uint(acc[i]))); acc2[i].y = __uint_as_float(q2); }
            if (MODE == 22) { uacc[i] = __vimin3_u32(uacc[i], (unsigned)it, __float_as_uint(a)); }   // VIMNMX3
            if (MODE == 23) {  // EMD mix per 2 pairs: 8 packed + 2 MUFU
#pragma unroll
                for (int q = 0; q < 8; q++) acc2[i] = __ffma2_rn(acc2[i], a2, b2);
                float r0, r1; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(acc2[i].x)); asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(acc2[i].y));
                acc2[i] = make_float2(r0, r1); }
            if (MODE == 24) {  // EMD mix per pair, scalar: 8 FFMA-class + 1 MUFU is mode 11; here 9 + 1
#pragma unroll
                for (int q = 0; q < 9; q++) acc[i] = __fmaf_rn(acc[i], a, b);
                float r; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(acc[i])); acc[i] = r; }
            if (MODE == 25) {  // 6 FFMA2 + 2 MUFU
#pragma unroll
                for (int q = 0; q < 6; q++) acc2[i] = __ffma2_rn(acc2[i], a2, b2);
                float r0, r1; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(acc2[i].x)); asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(acc2[i].y));
                acc2[i] = make_float2(r0, r1); }
            if (MODE == 26) {  // 3 FFMA2 + 2 FMNMX (2-input) : Chamfer both directions, 2 pairs -> needs 4 mins; here 2
#pragma unroll
                for (int q = 0; q < 3; q++) acc2[i] = __ffma2_rn(acc2[i], a2, b2);
                float r; asm volatile("min.f32 %0, %1, %2;" : "=f"(r) : "f"(acc[i]), "f"(acc2[i].x)); acc[i] = r;
                asm volatile("min.f32 %0, %1, %2;" : "=f"(r) : "f"(acc[(i + 1) % NACC]), "f"(acc2[i].y)); acc[(i + 1) % NACC] = r; }
            if (MODE == 27) {  // 3 FFMA2 + 4 FMNMX
#pragma unroll
                for (int q = 0; q < 3; q++) acc2[i] = __ffma2_rn(acc2[i], a2, b2);
                float r; asm volatile("min.f32 %0, %1, %2;" : "=f"(r) : "f"(acc[i]), "f"(acc2[i].x)); acc[i] = r;
                asm volatile("min.f32 %0, %1, %2;" : "=f"(r) : "f"(acc[i]), "f"(acc2[i].y)); acc[i] = r;
                unsigned u; asm volatile("min.u32 %0, %1, %2;" : "=r"(u) : "r"(uacc[i]), "r"(__float_as_uint(acc2[i].x))); uacc[i] = u;
                asm volatile("min.u32 %0, %1, %2;" : "=r"(u) : "r"(uacc[i]), "r"(__float_as_uint(acc2[i].y))); uacc[i] = u; }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += acc[i] + acc2[i].x + acc2[i].y + (float)uacc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char *name, double ops_per_iter_per_acc, int sms, float *out)
{
    const int blocks = sms * 8, threads = 256;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    kern<MODE><<<blocks, threads>>>(out, 1.0001f, 0.5f, 64);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        CK(cudaEventRecord(e0));
        kern<MODE><<<blocks, threads>>>(out, 1.0001f, 0.5f, ITERS);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    double ops = (double)blocks * threads * ITERS * NACC * ops_per_iter_per_acc;
    double per_s = ops / (best * 1e-3);
    printf("%-28s %8.3f ms  %9.2f Gop/s  %7.2f op/clk/SM @max-clock(%d MHz)\n", name, best, per_s * 1e-9,
           per_s / ((double)clk_khz * 1e3) / sms, clk_khz / 1000);
}

int main()
{
    int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    float *out; CK(cudaMalloc(&out, sizeof(float) * sms * 8 * 256));
    printf("SMs=%d  (op = one thread-level instruction; a full-rate pipe is 128 op/clk/SM)\n", sms);
    run<0>("FFMA", 1, sms, out);
    run<1>("FFMA2 (instr)", 1, sms, out);
    run<5>("FADD", 1, sms, out);
    run<6>("FADD2 (instr)", 1, sms, out);
    run<2>("FMNMX", 1, sms, out);
    run<3>("FMNMX3", 1, sms, out);
    run<4>("MUFU.EX2", 1, sms, out);
    run<7>("1 FFMA + 1 MUFU (groups)", 1, sms, out);
    run<10>("4 FFMA + 1 MUFU (groups)", 1, sms, out);
    run<11>("8 FFMA + 1 MUFU (groups)", 1, sms, out);
    run<12>("4 FFMA2 + 2 MUFU (groups)", 1, sms, out);
    run<13>("3 FFMA2 + 1 FMNMX3 (groups)", 1, sms, out);
    run<14>("6 FFMA + 1 FMNMX3 (groups)", 1, sms, out);
    run<15>("3 FFMA2 + 2 FMNMX3 (groups)", 1, sms, out);
    run<16>("FMNMX", 1, sms, out);
    run<17>("6 FFMA + 1 FMNMX (groups)", 1, sms, out);
    run<18>("6 FFMA + 2 FMNMX (groups)", 1, sms, out);
    run<19>("6 FFMA + cmp/sel/sel (grp)", 1, sms, out);
    run<20>("6 FFMA + 1 IADD (groups)", 1, sms, out);
    run<21>("6 FFMA + 2 IMNMX (groups)", 1, sms, out);
    run<22>("VIMNMX3", 1, sms, out);
    run<23>("8 FFMA2 + 2 MUFU (groups)", 1, sms, out);
    run<24>("9 FFMA + 1 MUFU (groups)", 1, sms, out);
    run<25>("6 FFMA2 + 2 MUFU (groups)", 1, sms, out);
    run<26>("3 FFMA2 + 2 FMNMX (groups)", 1, sms, out);
    run<27>("3 FFMA2 + 2FMNMX+2IMNMX", 1, sms, out);
    run<8>("REDUX.MIN", 1, sms, out);
    run<9>("SHFL", 1, sms, out);
    return 0;
}
