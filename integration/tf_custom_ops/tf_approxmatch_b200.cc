// tf_approxmatch_b200.cc -- optional TensorFlow custom-op wrapper over libpnae.so.
//
// SOURCE ONLY (TensorFlow is not installable in the build image; not compiled or tested here).
// Registers the reference's three ops (tf_ops/approxmatch/tf_approxmatch.cpp:7-21) with the
// same names and signatures, GPU kernels only, so tf_approxmatch.py and models/model_emd.py
// work unchanged.  `match` stays the reference's dense (b,m,n) tensor on this path; the
// factor-form fast path (no dense tensor) is the fourth op, ApproxMatchCost, below.
#include "tensorflow/core/framework/op.h"
#include "tensorflow/core/framework/op_kernel.h"
#include "tensorflow/core/util/gpu_kernel_helper.h"

#include "pnae.h"

using namespace tensorflow;

REGISTER_OP("ApproxMatch").Input("xyz1: float32").Input("xyz2: float32").Output("match: float32");
REGISTER_OP("MatchCost").Input("xyz1: float32").Input("xyz2: float32").Input("match: float32").Output("cost: float32");
REGISTER_OP("MatchCostGrad").Input("xyz1: float32").Input("xyz2: float32").Input("match: float32")
    .Output("grad1: float32").Output("grad2: float32");
// fused: approx_match + match_cost + match_cost_grad without the dense tensor
REGISTER_OP("ApproxMatchCost").Input("xyz1: float32").Input("xyz2: float32")
    .Output("cost: float32").Output("grad1: float32").Output("grad2: float32");

namespace {
bool CheckClouds(OpKernelContext* ctx, const char* op, const Tensor& a, const Tensor& c, int* b, int* n, int* m) {
  if (!(a.dims() == 3 && a.shape().dim_size(2) == 3)) {
    ctx->SetStatus(errors::InvalidArgument(op, " expects (batch_size,num_points,3) xyz1 shape")); return false; }
  *b = a.shape().dim_size(0); *n = a.shape().dim_size(1);
  if (!(c.dims() == 3 && c.shape().dim_size(2) == 3 && c.shape().dim_size(0) == *b)) {
    ctx->SetStatus(errors::InvalidArgument(op, " expects (batch_size,num_points,3) xyz2 shape, and batch_size must match")); return false; }
  *m = c.shape().dim_size(1);
  return true;
}
bool CheckMatch(OpKernelContext* ctx, const Tensor& mt, int b, int n, int m) {
  if (!(mt.dims() == 3 && mt.shape().dim_size(0) == b && mt.shape().dim_size(1) == m && mt.shape().dim_size(2) == n)) {
    ctx->SetStatus(errors::InvalidArgument("MatchCost expects (batch_size,#query,#dataset) match shape")); return false; }
  return true;
}
// factors + workspace temporaries for pnae_approx_match
Status AllocScratch(OpKernelContext* ctx, int b, int n, int m, Tensor* factors, Tensor* ws, size_t* ws_bytes) {
  TF_RETURN_IF_ERROR(ctx->allocate_temp(DT_FLOAT, TensorShape{b, PNAE_NUM_LEVELS, n + m}, factors));
  *ws_bytes = pnae_approx_match_workspace_bytes(b, n, m);
  return ctx->allocate_temp(DT_UINT8, TensorShape{static_cast<int64>(*ws_bytes ? *ws_bytes : 1)}, ws);
}
}  // namespace

class ApproxMatchB200Op : public OpKernel {
 public:
  explicit ApproxMatchB200Op(OpKernelConstruction* c) : OpKernel(c) {}
  void Compute(OpKernelContext* ctx) override {
    const Tensor& x1 = ctx->input(0); const Tensor& x2 = ctx->input(1);
    int b, n, m;
    if (!CheckClouds(ctx, "ApproxMatch", x1, x2, &b, &n, &m)) return;
    Tensor* match;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape{b, m, n}, &match));
    Tensor factors, ws; size_t ws_bytes;
    OP_REQUIRES_OK(ctx, AllocScratch(ctx, b, n, m, &factors, &ws, &ws_bytes));
    const int rc = pnae_approx_match(b, n, m, x1.flat<float>().data(), x2.flat<float>().data(), factors.flat<float>().data(),
                                     match->flat<float>().data(), ws.flat<uint8>().data(), ws_bytes, (void*)GetGpuStream(ctx));
    OP_REQUIRES(ctx, rc == PNAE_OK, errors::Internal("ApproxMatch: ", pnae_last_error()));
  }
};
REGISTER_KERNEL_BUILDER(Name("ApproxMatch").Device(DEVICE_GPU), ApproxMatchB200Op);

class MatchCostB200Op : public OpKernel {
 public:
  explicit MatchCostB200Op(OpKernelConstruction* c) : OpKernel(c) {}
  void Compute(OpKernelContext* ctx) override {
    const Tensor& x1 = ctx->input(0); const Tensor& x2 = ctx->input(1); const Tensor& mt = ctx->input(2);
    int b, n, m;
    if (!CheckClouds(ctx, "MatchCost", x1, x2, &b, &n, &m) || !CheckMatch(ctx, mt, b, n, m)) return;
    Tensor* cost;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape{b}, &cost));
    const int rc = pnae_match_cost_fwd(b, n, m, x1.flat<float>().data(), x2.flat<float>().data(), mt.flat<float>().data(),
                                       cost->flat<float>().data(), (void*)GetGpuStream(ctx));
    OP_REQUIRES(ctx, rc == PNAE_OK, errors::Internal("MatchCost: ", pnae_last_error()));
  }
};
REGISTER_KERNEL_BUILDER(Name("MatchCost").Device(DEVICE_GPU), MatchCostB200Op);

class MatchCostGradB200Op : public OpKernel {
 public:
  explicit MatchCostGradB200Op(OpKernelConstruction* c) : OpKernel(c) {}
  void Compute(OpKernelContext* ctx) override {
    const Tensor& x1 = ctx->input(0); const Tensor& x2 = ctx->input(1); const Tensor& mt = ctx->input(2);
    int b, n, m;
    if (!CheckClouds(ctx, "MatchCostGrad", x1, x2, &b, &n, &m) || !CheckMatch(ctx, mt, b, n, m)) return;
    Tensor *g1, *g2;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape{b, n, 3}, &g1));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, TensorShape{b, m, 3}, &g2));
    const int rc = pnae_match_cost_bwd(b, n, m, x1.flat<float>().data(), x2.flat<float>().data(), mt.flat<float>().data(),
                                       g1->flat<float>().data(), g2->flat<float>().data(), (void*)GetGpuStream(ctx));
    OP_REQUIRES(ctx, rc == PNAE_OK, errors::Internal("MatchCostGrad: ", pnae_last_error()));
  }
};
REGISTER_KERNEL_BUILDER(Name("MatchCostGrad").Device(DEVICE_GPU), MatchCostGradB200Op);

class ApproxMatchCostB200Op : public OpKernel {
 public:
  explicit ApproxMatchCostB200Op(OpKernelConstruction* c) : OpKernel(c) {}
  void Compute(OpKernelContext* ctx) override {
    const Tensor& x1 = ctx->input(0); const Tensor& x2 = ctx->input(1);
    int b, n, m;
    if (!CheckClouds(ctx, "ApproxMatch", x1, x2, &b, &n, &m)) return;
    Tensor *cost, *g1, *g2;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape{b}, &cost));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, TensorShape{b, n, 3}, &g1));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(2, TensorShape{b, m, 3}, &g2));
    Tensor factors, ws; size_t ws_bytes;
    OP_REQUIRES_OK(ctx, AllocScratch(ctx, b, n, m, &factors, &ws, &ws_bytes));
    void* st = (void*)GetGpuStream(ctx);
    int rc = pnae_approx_match(b, n, m, x1.flat<float>().data(), x2.flat<float>().data(), factors.flat<float>().data(),
                               nullptr, ws.flat<uint8>().data(), ws_bytes, st);
    if (rc == PNAE_OK)
      rc = pnae_match_cost_factors(b, n, m, x1.flat<float>().data(), x2.flat<float>().data(), factors.flat<float>().data(),
                                   cost->flat<float>().data(), g1->flat<float>().data(), g2->flat<float>().data(), st);
    OP_REQUIRES(ctx, rc == PNAE_OK, errors::Internal("ApproxMatchCost: ", pnae_last_error()));
  }
};
REGISTER_KERNEL_BUILDER(Name("ApproxMatchCost").Device(DEVICE_GPU), ApproxMatchCostB200Op);
