// tf_nndistance_b200.cc -- optional TensorFlow custom-op wrapper over libpnae.so.
//
// SOURCE ONLY: TensorFlow is not installable in the build image, so this file is not compiled
// or tested here.  It registers the SAME two ops as the reference's
// tf_ops/nn_distance/tf_nndistance.cpp (names, inputs, outputs: :3-18), GPU kernels only
// (there is no CPU fallback), so tf_nndistance.py and models/*.py work unchanged:
//
//   g++ -std=c++14 -shared -fPIC tf_nndistance_b200.cc -o tf_nndistance_so.so \
//       -I$TF_INC -I<repo>/include -L$TF_LIB -ltensorflow_framework -L<repo>/pointnet_autoencoder_b200 -lpnae
#include "tensorflow/core/framework/op.h"
#include "tensorflow/core/framework/op_kernel.h"
#include "tensorflow/core/util/gpu_kernel_helper.h"

#include "pnae.h"

using namespace tensorflow;

REGISTER_OP("NnDistance")
    .Input("xyz1: float32")
    .Input("xyz2: float32")
    .Output("dist1: float32")
    .Output("idx1: int32")
    .Output("dist2: float32")
    .Output("idx2: int32");
REGISTER_OP("NnDistanceGrad")
    .Input("xyz1: float32")
    .Input("xyz2: float32")
    .Input("grad_dist1: float32")
    .Input("idx1: int32")
    .Input("grad_dist2: float32")
    .Input("idx2: int32")
    .Output("grad_xyz1: float32")
    .Output("grad_xyz2: float32");

namespace {
bool CheckClouds(OpKernelContext* ctx, const char* op, const Tensor& a, const Tensor& c, int* b, int* n, int* m) {
  // same conditions and messages as tf_nndistance.cpp:175-182
  if (a.dims() != 3) { ctx->SetStatus(errors::InvalidArgument(op, " requires xyz1 be of shape (batch,#points,3)")); return false; }
  if (a.shape().dim_size(2) != 3) { ctx->SetStatus(errors::InvalidArgument(op, " only accepts 3d point set xyz1")); return false; }
  if (c.dims() != 3) { ctx->SetStatus(errors::InvalidArgument(op, " requires xyz2 be of shape (batch,#points,3)")); return false; }
  if (c.shape().dim_size(2) != 3) { ctx->SetStatus(errors::InvalidArgument(op, " only accepts 3d point set xyz2")); return false; }
  *b = a.shape().dim_size(0); *n = a.shape().dim_size(1); *m = c.shape().dim_size(1);
  if (c.shape().dim_size(0) != *b) { ctx->SetStatus(errors::InvalidArgument(op, " expects xyz1 and xyz2 have same batch size")); return false; }
  return true;
}
}  // namespace

class NnDistanceB200Op : public OpKernel {
 public:
  explicit NnDistanceB200Op(OpKernelConstruction* c) : OpKernel(c) {}
  void Compute(OpKernelContext* ctx) override {
    const Tensor& x1 = ctx->input(0);
    const Tensor& x2 = ctx->input(1);
    int b, n, m;
    if (!CheckClouds(ctx, "NnDistance", x1, x2, &b, &n, &m)) return;
    Tensor *d1, *i1, *d2, *i2;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape{b, n}, &d1));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, TensorShape{b, n}, &i1));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(2, TensorShape{b, m}, &d2));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(3, TensorShape{b, m}, &i2));
    const size_t ws_bytes = pnae_nn_distance_workspace_bytes(b, n, m);
    Tensor ws;
    OP_REQUIRES_OK(ctx, ctx->allocate_temp(DT_UINT8, TensorShape{static_cast<int64>(ws_bytes ? ws_bytes : 1)}, &ws));
    const int rc = pnae_nn_distance_fwd(b, n, x1.flat<float>().data(), m, x2.flat<float>().data(),
                                        d1->flat<float>().data(), i1->flat<int>().data(),
                                        d2->flat<float>().data(), i2->flat<int>().data(),
                                        ws.flat<uint8>().data(), ws_bytes, (void*)GetGpuStream(ctx));
    OP_REQUIRES(ctx, rc == PNAE_OK, errors::Internal("NnDistance: ", pnae_last_error()));
  }
};
REGISTER_KERNEL_BUILDER(Name("NnDistance").Device(DEVICE_GPU), NnDistanceB200Op);

class NnDistanceGradB200Op : public OpKernel {
 public:
  explicit NnDistanceGradB200Op(OpKernelConstruction* c) : OpKernel(c) {}
  void Compute(OpKernelContext* ctx) override {
    const Tensor& x1 = ctx->input(0);
    const Tensor& x2 = ctx->input(1);
    const Tensor& g1 = ctx->input(2);
    const Tensor& i1 = ctx->input(3);
    const Tensor& g2 = ctx->input(4);
    const Tensor& i2 = ctx->input(5);
    int b, n, m;
    if (!CheckClouds(ctx, "NnDistanceGrad", x1, x2, &b, &n, &m)) return;
    // tf_nndistance.cpp:227-230
    OP_REQUIRES(ctx, g1.shape() == (TensorShape{b, n}), errors::InvalidArgument("NnDistanceGrad requires grad_dist1 be of shape(batch,#points)"));
    OP_REQUIRES(ctx, i1.shape() == (TensorShape{b, n}), errors::InvalidArgument("NnDistanceGrad requires idx1 be of shape(batch,#points)"));
    OP_REQUIRES(ctx, g2.shape() == (TensorShape{b, m}), errors::InvalidArgument("NnDistanceGrad requires grad_dist2 be of shape(batch,#points)"));
    OP_REQUIRES(ctx, i2.shape() == (TensorShape{b, m}), errors::InvalidArgument("NnDistanceGrad requires idx2 be of shape(batch,#points)"));
    Tensor *o1, *o2;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape{b, n, 3}, &o1));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, TensorShape{b, m, 3}, &o2));
    const int rc = pnae_nn_distance_bwd(b, n, x1.flat<float>().data(), m, x2.flat<float>().data(),
                                        g1.flat<float>().data(), i1.flat<int>().data(),
                                        g2.flat<float>().data(), i2.flat<int>().data(),
                                        o1->flat<float>().data(), o2->flat<float>().data(), (void*)GetGpuStream(ctx));
    OP_REQUIRES(ctx, rc == PNAE_OK, errors::Internal("NnDistanceGrad: ", pnae_last_error()));
  }
};
REGISTER_KERNEL_BUILDER(Name("NnDistanceGrad").Device(DEVICE_GPU), NnDistanceGradB200Op);
