// ref_gpu_shim.cu -- extern "C" doors onto the reference's own CUDA launchers.
//
// TEST INFRASTRUCTURE ONLY (see oracle/oracle.c header).  No reference code
// here: oracle/build.py compiles tf_ops/nn_distance/tf_nndistance_g.cu and
// tf_ops/approxmatch/tf_approxmatch_g.cu UNMODIFIED from /root/reference with
// `nvcc -O2 -gencode arch=compute_100a,code=sm_100a` (their own flags plus the
// arch) and links the objects with this shim into oracle/_ref/libref_gpu.so.
// It is the same-box incumbent and the GPU parity pin.
#include <cuda_runtime.h>
#include <math.h>

// C++-linkage launchers defined in the reference .cu files.
void NmDistanceKernelLauncher(int b, int n, const float *xyz, int m, const float *xyz2,
                              float *result, int *result_i, float *result2, int *result2_i);   // tf_nndistance_g.cu:128
void NmDistanceGradKernelLauncher(int b, int n, const float *xyz1, int m, const float *xyz2,
                                  const float *grad_dist1, const int *idx1,
                                  const float *grad_dist2, const int *idx2,
                                  float *grad_xyz1, float *grad_xyz2);                          // tf_nndistance_g.cu:152
void approxmatchLauncher(int b, int n, int m, const float *xyz1, const float *xyz2,
                         float *match, float *temp);                                            // tf_approxmatch_g.cu:180
void matchcostLauncher(int b, int n, int m, const float *xyz1, const float *xyz2,
                       const float *match, float *out);                                         // tf_approxmatch_g.cu:226
void matchcostgradLauncher(int b, int n, int m, const float *xyz1, const float *xyz2,
                           const float *match, float *grad1, float *grad2);                     // tf_approxmatch_g.cu:292

// level_j exactly as the reference kernel forms it (tf_approxmatch_g.cu:22-25),
// so a test can confirm powf(4,j) is an exact power of four on this toolchain.
__global__ void ref_levels_kernel(float *out)
{
    int t = 0;
    for (int j = 7; j >= -2; j--, t++) {
        float level = -powf(4.0f, j);
        if (j == -2) level = 0;
        out[t] = level;
    }
}

extern "C" {

// All pointers are DEVICE pointers; launches go to the legacy default stream
// like the reference's; the return value is cudaGetLastError() after the launch.
int ref_gpu_nn_distance(int b, int n, const float *xyz1, int m, const float *xyz2,
                        float *dist1, int *idx1, float *dist2, int *idx2)
{
    NmDistanceKernelLauncher(b, n, xyz1, m, xyz2, dist1, idx1, dist2, idx2);
    return (int)cudaGetLastError();
}

int ref_gpu_nn_distance_grad(int b, int n, const float *xyz1, int m, const float *xyz2,
                             const float *grad_dist1, const int *idx1,
                             const float *grad_dist2, const int *idx2,
                             float *grad_xyz1, float *grad_xyz2)
{
    NmDistanceGradKernelLauncher(b, n, xyz1, m, xyz2, grad_dist1, idx1, grad_dist2, idx2, grad_xyz1, grad_xyz2);
    return (int)cudaGetLastError();
}

// temp: (b, 2(n+m)) floats of scratch (tf_approxmatch.cpp:168)
int ref_gpu_approxmatch(int b, int n, int m, const float *xyz1, const float *xyz2, float *match, float *temp)
{
    approxmatchLauncher(b, n, m, xyz1, xyz2, match, temp);
    return (int)cudaGetLastError();
}

int ref_gpu_matchcost(int b, int n, int m, const float *xyz1, const float *xyz2, const float *match, float *cost)
{
    matchcostLauncher(b, n, m, xyz1, xyz2, match, cost);
    return (int)cudaGetLastError();
}

int ref_gpu_matchcostgrad(int b, int n, int m, const float *xyz1, const float *xyz2, const float *match,
                          float *grad1, float *grad2)
{
    matchcostgradLauncher(b, n, m, xyz1, xyz2, match, grad1, grad2);
    return (int)cudaGetLastError();
}

int ref_gpu_levels(float *out10_device)
{
    ref_levels_kernel<<<1, 1>>>(out10_device);
    return (int)cudaGetLastError();
}

int ref_gpu_sync(void) { return (int)cudaDeviceSynchronize(); }

}  // extern "C"
