/*
 * pnae.h -- C ABI of the B200-native reconstruction-loss ops of pointnet-autoencoder.
 *
 * Drop-in boundary.  Each entry point replaces one C++-linkage host launcher that
 * the reference's TensorFlow op glue calls (cited per function, paths relative to
 * the reference root).  Conventions shared by all of them, mirroring the reference:
 *   - plain pointers and ints only; every data pointer is a DEVICE pointer into
 *     memory the CALLER owns (outputs and scratch included: the library never
 *     allocates device memory and keeps no state between calls);
 *   - row-major float32 point sets (b, n, 3) / (b, m, 3), int32 indices;
 *   - outputs are fully overwritten (no pre-zeroing needed);
 *   - the Chamfer entry points take (b, n, xyz1, m, xyz2, ...), the EMD ones
 *     (b, n, m, xyz1, xyz2, ...) -- the reference's own argument-order quirk.
 * What is new relative to the reference launchers:
 *   - an explicit `stream` (a cudaStream_t passed as void*; NULL = legacy default
 *     stream, which is what the reference always used);
 *   - an int status: 0 on success, a negative pnae_status otherwise, with a
 *     thread-local message from pnae_last_error() (the reference checks nothing);
 *   - caller-provided workspace where an op needs scratch (the reference's TF glue
 *     did the same through allocate_temp, tf_approxmatch.cpp:168).
 * There is no CPU fallback: without a CUDA device every compute call fails with
 * PNAE_ERR_CUDA.
 */
#ifndef PNAE_H_
#define PNAE_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define PNAE_API __attribute__((visibility("default")))
#else
#define PNAE_API
#endif

#define PNAE_VERSION 100          /* 0.1.0 */
#define PNAE_NUM_LEVELS 10        /* j = 7..-2, tf_approxmatch_g.cu:21 */

typedef enum pnae_status {
    PNAE_OK = 0,
    PNAE_ERR_INVALID_ARG = -1,    /* bad shape / NULL pointer / misaligned pointer */
    PNAE_ERR_WORKSPACE = -2,      /* workspace too small */
    PNAE_ERR_CUDA = -3,           /* launch or runtime failure; see pnae_last_error() */
    PNAE_ERR_UNSUPPORTED = -4     /* device is not sm_100 */
} pnae_status;

PNAE_API int pnae_version(void);
/* Message of the last failing call made on this thread ("" if none). */
PNAE_API const char *pnae_last_error(void);
/* SM count and compute capability of the current device. */
PNAE_API int pnae_device_info(int *sm_count, int *cc_major, int *cc_minor);

/* FP32 issue-rate probe (measurement aid, not part of the reference's interface): launches 8 CTAs x 256 threads
 * per SM, each thread running 8 independent chains of `iters` FFMAs; *flop receives the FLOPs of the launch
 * (FMA = 2).  out: device scratch of at least sm_count*8*256 floats.  The caller times the launch. */
PNAE_API int pnae_fp32_probe(int iters, float *out, size_t out_floats, long long *flop, void *stream);

/* ---- Chamfer distance -------------------------------------------------- */

/* Scratch bytes pnae_nn_distance_fwd needs for these sizes (may be 0). */
PNAE_API size_t pnae_nn_distance_workspace_bytes(int b, int n, int m);

/* Introspection (host only, no device access): how pnae_nn_distance_fwd would split these sizes on a device
 * with sm_count SMs.  plan[9] = { row blocks, column chunks, row-key slots per row block in the workspace,
 * elements per launch, sweep warps, slots in use by a full launch, slots in use by the last launch,
 * rows per block, columns per chunk }.  tests/test_cabi.py checks the stream-K slot arithmetic against it. */
PNAE_API int pnae_nn_distance_plan(int b, int n, int m, int sm_count, int *plan);

/* Replaces NmDistanceKernelLauncher (tf_ops/nn_distance/tf_nndistance_g.cu:128-131).
 * dist1[i,j] = min_k |xyz1[i,j]-xyz2[i,k]|^2 (squared), idx1 = lowest-index argmin;
 * dist2/idx2 the same with the roles swapped.  n, m >= 1. */
PNAE_API int pnae_nn_distance_fwd(int b, int n, const float *xyz1, int m, const float *xyz2,
                         float *dist1, int *idx1, float *dist2, int *idx2,
                         void *workspace, size_t workspace_bytes, void *stream);

/* Replaces NmDistanceGradKernelLauncher (tf_nndistance_g.cu:152-157).
 * grad_xyz1 (b,n,3), grad_xyz2 (b,m,3); zeroed inside, like the reference. */
PNAE_API int pnae_nn_distance_bwd(int b, int n, const float *xyz1, int m, const float *xyz2,
                         const float *grad_dist1, const int *idx1,
                         const float *grad_dist2, const int *idx2,
                         float *grad_xyz1, float *grad_xyz2, void *stream);

/* Fused Chamfer loss + gradient (the loss of models/model.py:80-83 without the op seam):
 *   *loss      = w1 * sum(dist1) + w2 * sum(dist2)          (model.py: w1 = 100/(b*n), w2 = 100/(b*m))
 *   grad_xyz1/2 = d loss / d xyz1, d loss / d xyz2           ((b,n,3), (b,m,3), fully overwritten)
 * in the two launches of the forward (no gradient launch, no dist/idx round trip).  dist1/idx1 and
 * dist2/idx2 are optional outputs (NULL pairs to skip).  Same workspace as pnae_nn_distance_fwd. */
PNAE_API int pnae_chamfer_loss_grad(int b, int n, const float *xyz1, int m, const float *xyz2, float w1, float w2,
                                    float *loss, float *grad_xyz1, float *grad_xyz2,
                                    float *dist1, int *idx1, float *dist2, int *idx2,
                                    void *workspace, size_t workspace_bytes, void *stream);

/* NnDistance and NnDistanceGrad in ONE call, for callers whose upstream gradients do not depend on this call's
 * distances (every loss that is linear in dist1/dist2, e.g. the Chamfer loss of models/model.py:80-83, whose
 * grad_dist is the constant 100/(b*n)).  Replaces NmDistanceKernelLauncher followed by NmDistanceGradKernelLauncher
 * (tf_nndistance_g.cu:128-157): same outputs as pnae_nn_distance_fwd + pnae_nn_distance_bwd, two launches instead of
 * three and no idx/dist round trip between them.  Same workspace as pnae_nn_distance_fwd. */
PNAE_API int pnae_nn_distance_fwd_grad(int b, int n, const float *xyz1, int m, const float *xyz2,
                                       const float *grad_dist1, const float *grad_dist2,
                                       float *dist1, int *idx1, float *dist2, int *idx2,
                                       float *grad_xyz1, float *grad_xyz2,
                                       void *workspace, size_t workspace_bytes, void *stream);

/* One Chamfer step (pnae_nn_distance_fwd + pnae_nn_distance_bwd over FIXED buffers) captured into a
 * CUDA graph: at B=32, N=M=2048 the step is three kernels and ~56 us of GPU time, so one launch per
 * step instead of three is the difference between GPU-bound and host-bound.  The handle owns only
 * the graph objects; every buffer stays the caller's and must outlive the handle. */
PNAE_API int pnae_chamfer_graph_create(int b, int n, const float *xyz1, int m, const float *xyz2,
                                       float *dist1, int *idx1, float *dist2, int *idx2,
                                       const float *grad_dist1, const float *grad_dist2,
                                       float *grad_xyz1, float *grad_xyz2,
                                       void *workspace, size_t workspace_bytes, void **handle);
/* The same for `steps` consecutive steps over different inputs (xyz1[s], xyz2[s]) and shared outputs: one
 * launch replays them back to back (results of the last step remain in the output buffers). */
PNAE_API int pnae_chamfer_graph_create_multi(int steps, int b, int n, const float *const *xyz1, int m, const float *const *xyz2,
                                             float *dist1, int *idx1, float *dist2, int *idx2,
                                             const float *grad_dist1, const float *grad_dist2,
                                             float *grad_xyz1, float *grad_xyz2,
                                             void *workspace, size_t workspace_bytes, void **handle);
/* The same over pnae_nn_distance_fwd_grad (two kernels per step); gradient outputs are required. */
PNAE_API int pnae_chamfer_graph_create_fused_multi(int steps, int b, int n, const float *const *xyz1, int m, const float *const *xyz2,
                                                   float *dist1, int *idx1, float *dist2, int *idx2,
                                                   const float *grad_dist1, const float *grad_dist2,
                                                   float *grad_xyz1, float *grad_xyz2,
                                                   void *workspace, size_t workspace_bytes, void **handle);
/* The multi-step graph, software-pipelined: step s+1's sweep runs while step s's finalize (and gradient) resolve.  Steps
 * cycle through `nsets` >= 2 output sets and `nws` workspaces, 2 <= nws <= nsets (each of pnae_nn_distance_workspace_bytes):
 * the output pointer lists have `nsets` entries, `workspace` has `nws`, xyz1 / xyz2 `steps`; step s writes set s % nsets
 * -- after a launch the last step's results are in set (steps - 1) % nsets, the one before it in the set before that;
 * with nsets == steps every step keeps its own results.  Three workspaces let the sweeps follow each other without a gap.  fused != 0: sweep + finalize-with-gradients
 * (pnae_nn_distance_fwd_grad); fused == 0: pnae_nn_distance_fwd, plus pnae_nn_distance_bwd when gradient outputs are
 * given.  Results are those of the sequential graph, step for step.  The whole batch must fit one launch
 * (pnae_nn_distance_plan: be == b). */
PNAE_API int pnae_chamfer_graph_create_pipelined(int fused, int steps, int nsets, int nws, int b, int n, const float *const *xyz1, int m,
                                                 const float *const *xyz2, float *const *dist1, int *const *idx1,
                                                 float *const *dist2, int *const *idx2,
                                                 const float *grad_dist1, const float *grad_dist2,
                                                 float *const *grad_xyz1, float *const *grad_xyz2,
                                                 void *const *workspace, size_t workspace_bytes, void **handle);
PNAE_API int pnae_graph_launch(void *handle, void *stream);
PNAE_API int pnae_graph_destroy(void *handle);

/* Streaming host-buffer form of the Chamfer step: what a caller without device data runs (the reference feeds its
 * op from host memory through feed_dict and reads results through sess.run, train.py:196-206).  `depth` buffer sets
 * on three internal streams: the host->device copy of submission i+1, the kernels of submission i and the device->host
 * copy of submission i-1 overlap.  A submission is `steps` consecutive batches (>= 1) run as ONE CUDA graph.  All memory
 * is the caller's: per set a device xyz1 (steps,b,n,3), a device xyz2 (steps,b,m,3), a device result buffer of `steps`
 * blocks `out_stride` bytes apart and a PINNED host result buffer of the same layout; out_offsets[6] are the byte offsets
 * of { grad_xyz1, grad_xyz2, dist1, idx1, dist2, idx2 } inside a block and d2h_bytes the leading bytes of every block
 * copied back (the whole block, or only the gradients when they are laid out first).  One workspace
 * (pnae_nn_distance_workspace_bytes) is shared: the steps run in order on one stream.  fused != 0 selects
 * pnae_nn_distance_fwd_grad (two kernels per step) over fwd + bwd (three). */
PNAE_API int pnae_chamfer_host_pipeline_create(int depth, int steps, int b, int n, int m, int fused,
                                               float *const *d_xyz1, float *const *d_xyz2,
                                               void *const *d_out, void *const *h_out, const size_t *out_offsets,
                                               size_t out_stride, size_t d2h_bytes,
                                               const float *grad_dist1, const float *grad_dist2,
                                               void *workspace, size_t workspace_bytes, void **handle);
/* Enqueue one submission (`steps` batches, contiguous in h_xyz1 / h_xyz2) on host inputs (pinned for true asynchrony).  *retired = index of the buffer set whose results
 * are now complete in its host buffer and stay valid until the next submit, or -1 while the pipeline fills. */
PNAE_API int pnae_chamfer_host_pipeline_submit(void *handle, const float *h_xyz1, const float *h_xyz2, int *retired);
/* Wait for everything in flight: retired[0..*count) = completed buffer sets, oldest first (retired needs depth ints). */
PNAE_API int pnae_chamfer_host_pipeline_drain(void *handle, int *retired, int *count);
PNAE_API int pnae_chamfer_host_pipeline_destroy(void *handle);

/* ---- approximate earth mover's distance -------------------------------- */

/* Introspection (host only, no device access): plan[6] = { cooperative grid size, partial-sum slots per own
 * block, max(n,m), own points per task, streamed points per task, threads per CTA } for a device with sm_count
 * SMs.  tests/test_cabi.py checks the sweep's slot arithmetic against it. */
PNAE_API int pnae_approx_match_plan(int b, int n, int m, int sm_count, int *plan);

/* Scratch bytes pnae_approx_match needs (the reference's `temp`, tf_approxmatch.cpp:168). */
PNAE_API size_t pnae_approx_match_workspace_bytes(int b, int n, int m);

/* Replaces approxmatchLauncher (tf_ops/approxmatch/tf_approxmatch_g.cu:180-182).
 * xyz1 (b,n,3) "dataset", xyz2 (b,m,3) "query".
 *   factors: (b, PNAE_NUM_LEVELS, n+m) float32, REQUIRED.  Per level the vectors
 *            ratioL[0..n) then ratioR[0..m) of that level; they determine the soft
 *            assignment completely:
 *              match[i,l,k] = sum_j exp(level_j |xyz1[i,k]-xyz2[i,l]|^2) ratioL_j[k] ratioR_j[l]
 *            (tf_approxmatch_g.cu:145-153 is the only write to `match`).
 *   match:   (b, m, n) float32 or NULL.  The reference's dense output; pass NULL to
 *            keep the b*m*n tensor out of HBM and feed `factors` to
 *            pnae_match_cost_factors instead. */
PNAE_API int pnae_approx_match(int b, int n, int m, const float *xyz1, const float *xyz2,
                      float *factors, float *match,
                      void *workspace, size_t workspace_bytes, void *stream);

/* Dense (b,m,n) match from the factors (same accumulation order as the reference). */
PNAE_API int pnae_match_from_factors(int b, int n, int m, const float *xyz1, const float *xyz2,
                            const float *factors, float *match, void *stream);

/* Replaces matchcostLauncher (tf_approxmatch_g.cu:226-228): cost (b,) from a dense match. */
PNAE_API int pnae_match_cost_fwd(int b, int n, int m, const float *xyz1, const float *xyz2,
                        const float *match, float *cost, void *stream);

/* Replaces matchcostgradLauncher (tf_approxmatch_g.cu:292-295): grad1 (b,n,3), grad2 (b,m,3)
 * from a dense match (not yet scaled by the upstream grad_cost, as in the reference). */
PNAE_API int pnae_match_cost_bwd(int b, int n, int m, const float *xyz1, const float *xyz2,
                        const float *match, float *grad1, float *grad2, void *stream);

/* matchcost + matchcostgrad in one pass straight from the factors; the dense
 * tensor never exists.  grad1/grad2 may both be NULL (cost only). */
PNAE_API int pnae_match_cost_factors(int b, int n, int m, const float *xyz1, const float *xyz2,
                            const float *factors, float *cost, float *grad1, float *grad2,
                            void *stream);

/* ---- PointNet encoder: conv5 (1x1 conv = per-point linear map) + max-pool, fused ---------- */

/* Replaces, for the encoder's dominant layer, the TF graph segment of get_model
 * (models/model.py:57-66: tf_util.conv2d 128->1024 [utils/tf_util.py:155-185] followed by
 * tf_util.max_pool2d over the points [:368-391]).
 *   x_bf16  : (b, n, k)  bf16, the previous layer's activations (k = 64 or 128)
 *   wt_bf16 : (c, k)     bf16, the layer's weight TRANSPOSED (reference layout is [1,1,k,c]); c % 128 == 0
 * For y[i,p,ch] = sum_k x[i,p,k] * w[k,ch]  (fp32 accumulation on tcgen05 tensor cores, no bias) it writes,
 * per batch element i and channel ch, over the n points p:
 *   out_max, out_min, out_sum, out_sumsq : (b, c) float32 each.
 * The (b, n, c) activation is never written.  Bias, BatchNorm (batch statistics come from sum/sumsq,
 * folded inference statistics work the same way), ReLU and the max-pool then finish on (b, c):
 *   pooled = relu(s*(ext + bias) + t),  ext = out_max where the BN scale s >= 0, out_min where s < 0.
 * Optional (both or neither): sign (c,) float32 and out_arg (b, c) int32 -- out_arg[i,ch] is the index of the
 * first point attaining the maximum (sign[ch] >= 0) or the minimum (sign[ch] < 0) of channel ch; the
 * training backward routes the pooled gradient to that point. */
PNAE_API int pnae_encoder_conv_pool(int b, int n, int k, int c, const void *x_bf16, const void *wt_bf16,
                                    float *out_max, float *out_min, float *out_sum, float *out_sumsq,
                                    const float *sign, int *out_arg, int flags /* PNAE_OVERLAP_PREVIOUS or 0 */, void *stream);

/* ---- flags of the encoder entry points ---------------------------------------------------------
 * PNAE_STATS_ZEROED      the caller has already zeroed `stats` (all of it, tile-counter words included): the call
 *                        does not enqueue its own memset.  A chain of layers zeroes one arena once.
 * PNAE_OVERLAP_PREVIOUS  programmatic dependent launch: the kernel may be scheduled while the kernel enqueued
 *                        immediately before it on `stream` is still draining.  Its set-up -- staging this layer's
 *                        `w` / `bias` (`wt_bf16` for pnae_encoder_conv_pool), barriers, tensor memory -- runs under
 *                        that tail; every other read and every write waits until the predecessor has completed.
 *                        The caller promises that the immediately preceding kernel does not write `w`, `bias` or
 *                        `wt_bf16`.  Results are identical with and without the flag. */
#define PNAE_STATS_ZEROED 1
#define PNAE_OVERLAP_PREVIOUS 2

/* ---- PointNet encoder: layers 1-4 (3 -> 64 -> 64 -> 64 -> 128), one kernel per layer ---------- */

/* Replace, for the first four layers of get_model (models/model.py:40-56), the per-layer TF graph segment
 * conv2d -> bias -> BatchNorm -> ReLU (utils/tf_util.py:155-185, 514-533).  A layer kernel reads the previous layer's
 * RAW output (before BatchNorm), applies that layer's folded BatchNorm + ReLU  a = relu(s*y + t)  on the way in,
 * multiplies by its weight on the tensor cores (3xTF32: fp32 accuracy), adds the bias, writes its own raw output and
 * returns the per-channel sum and sum of squares of it: stats (2, kout) = { sum_p y[p,c], sum_p y[p,c]^2 }, from which
 * the caller forms the batch statistics of training-mode BatchNorm (or ignores them in inference mode).  For
 * pnae_mlp_layer the stats buffer must have room for 2*kout + kout/64 floats: the trailing words are the kernel's tile
 * counters.
 * npts = batch * points; all matrices row-major fp32; w is (kin, kout) like the reference's [1,1,kin,kout] kernel. */
PNAE_API int pnae_mlp_first(long long npts, const float *xyz, const float *w /* (3,64) */, const float *bias,
                            float *out /* (npts,64) */, float *stats /* (2,64) */, int flags, void *stream);
/* The previous layer's BatchNorm is given by its raw statistics (stats_prev (2,kin), training != 0: batch statistics, and
 * the kernel also performs TF's moving-average update of moving_mean_prev / moving_var_prev with the biased variance,
 * utils/tf_util.py:529-533) or by its moving statistics (training == 0; stats_prev may be NULL). */
PNAE_API int pnae_mlp_layer(long long npts, int kin /* 64 */, int kout /* 64 or 128 */, const float *in,
                            const float *stats_prev, const float *gamma_prev, const float *beta_prev,
                            float *moving_mean_prev, float *moving_var_prev, float eps, float decay, int training,
                            const float *w, const float *bias, float *out, float *stats, int flags, void *stream);
/* BatchNorm bookkeeping of one layer on its own (the layer kernels do this in their prologue; this entry exists for callers
 * that want the folded scale s = gamma/sqrt(var+eps) and shift t = beta - mean*s themselves). */
/* Layer 1 folded into layer 2: y1 = xyz @ w1 + b1 is affine in xyz, so its BatchNorm statistics follow from the moments
 * of xyz -- pnae_xyz_moments: moments (9 doubles, zero on entry unless the call zeroes them: flags) = the sums of x, y, z,
 * xx, xy, xz, yy, yz, zz over all points -- and layer 2's kernel forms relu(BatchNorm1(y1)) on the fly from xyz.  Same
 * results as pnae_mlp_first followed by pnae_mlp_layer up to the rounding of the statistics (double here), without the
 * (npts, 64) tensor in between; moving_mean1 / moving_var1 are updated as pnae_mlp_layer would (training != 0). */
PNAE_API int pnae_xyz_moments(long long npts, const float *xyz, double *moments /* (9) */, int flags, void *stream);
PNAE_API int pnae_mlp_layer_xyz(long long npts, const float *xyz, const double *moments, const float *w1 /* (3,64) */,
                                const float *b1, const float *gamma1, const float *beta1, float *moving_mean1,
                                float *moving_var1, float eps, float decay, int training, int kout /* 64 or 128 */,
                                const float *w /* (64,kout) */, const float *bias, float *out, float *stats, int flags,
                                void *stream);
PNAE_API int pnae_bn_fold(int k, const float *stats, double count, const float *gamma, const float *beta, float eps, float decay,
                          int training, float *moving_mean, float *moving_var, float *s_out, float *t_out, void *stream);
/* relu(BatchNorm(y)) of the last of these layers (BatchNorm given like in pnae_mlp_layer) as the bf16 (npts, k) operand of
 * pnae_encoder_conv_pool. */
PNAE_API int pnae_mlp_apply_bf16(long long npts, int k, const float *in, const float *stats, const float *gamma, const float *beta,
                                 float *moving_mean, float *moving_var, float eps, float decay, int training,
                                 void *out_bf16, int flags, void *stream);
/* conv5's bias + BatchNorm + ReLU + max-pool finish on (b, c), one launch: from pnae_encoder_conv_pool's max / min / sum /
 * sumsq to pooled (b,c) = relu((ext0 - mean0) * gamma * inv + beta), ext0 = max where gamma >= 0 else min; count = b * n.
 * Also returns what the backward needs: inv (c), mean0 (c) (of x @ w without the bias), ext0 (b,c), z (b,c). */
PNAE_API int pnae_conv5_finish(int b, int c, double count, const float *vmax, const float *vmin, const float *vsum, const float *vsq,
                               const float *bias, const float *gamma, const float *beta, float *moving_mean, float *moving_var,
                               float eps, float decay, int training, float *pooled, float *inv, float *mean0, float *ext0, float *z,
                               int flags, void *stream);

#ifdef __cplusplus
}
#endif
#endif  /* PNAE_H_ */
