// ref_cpu_shim.cpp -- extern "C" doors onto the reference's own CPU loops.
//
// TEST INFRASTRUCTURE ONLY (see oracle/oracle.c header).  This file contains no
// reference code: oracle/build.py cuts the TF-free function bodies out of the
// reference sources where they lie under /root/reference into a scratch
// directory outside the repo (REF_INC_DIR) and this shim #includes them:
//
//   ref_nnsearch.inc        tf_ops/nn_distance/tf_nndistance.cpp:21-43    nnsearch()
//   ref_nngrad_body.inc     tf_ops/nn_distance/tf_nndistance.cpp:126-163  body of NnDistanceGradOp::Compute
//   ref_approxmatch_cpu.inc tf_ops/approxmatch/tf_approxmatch.cpp:23-140  approxmatch_cpu, matchcost_cpu, matchcostgrad_cpu
//
// The result is oracle/_ref/libref_cpu.so (git-ignored, shipped to the GPU box).
#include <algorithm>
#include <vector>
#include <math.h>
#include <string.h>

#include "ref_nnsearch.inc"

static void ref_nngrad_impl(int b, int n, int m, const float *xyz1, const float *xyz2,
                            const float *grad_dist1, const int *idx1,
                            const float *grad_dist2, const int *idx2,
                            float *grad_xyz1, float *grad_xyz2)
{
#include "ref_nngrad_body.inc"
}

#include "ref_approxmatch_cpu.inc"

extern "C" {

// NnDistanceOp::Compute (tf_nndistance.cpp:79-80)
void ref_cpu_nn_distance(int b, int n, const float *xyz1, int m, const float *xyz2,
                         float *dist1, int *idx1, float *dist2, int *idx2)
{
    nnsearch(b, n, m, xyz1, xyz2, dist1, idx1);
    nnsearch(b, m, n, xyz2, xyz1, dist2, idx2);
}

void ref_cpu_nn_distance_grad(int b, int n, const float *xyz1, int m, const float *xyz2,
                              const float *grad_dist1, const int *idx1,
                              const float *grad_dist2, const int *idx2,
                              float *grad_xyz1, float *grad_xyz2)
{
    ref_nngrad_impl(b, n, m, xyz1, xyz2, grad_dist1, idx1, grad_dist2, idx2, grad_xyz1, grad_xyz2);
}

// NOTE: the CPU functions use an (n,m) match layout and 11 levels -- they are a
// timing baseline and a structural cross-check, not the parity oracle.
void ref_cpu_approxmatch(int b, int n, int m, const float *xyz1, const float *xyz2, float *match_nm)
{
    approxmatch_cpu(b, n, m, xyz1, xyz2, match_nm);
}

void ref_cpu_matchcost(int b, int n, int m, const float *xyz1, const float *xyz2,
                       const float *match_nm, float *cost)
{
    matchcost_cpu(b, n, m, xyz1, xyz2, match_nm, cost);
}

// matchcostgrad_cpu zeroes only grad1's x component (tf_approxmatch.cpp:108-109);
// pre-zero here so the caller sees the intended result.
void ref_cpu_matchcostgrad(int b, int n, int m, const float *xyz1, const float *xyz2,
                           const float *match_nm, float *grad1, float *grad2)
{
    memset(grad1, 0, sizeof(float) * (size_t)b * n * 3);
    matchcostgrad_cpu(b, n, m, xyz1, xyz2, match_nm, grad1, grad2);
}

}  // extern "C"
