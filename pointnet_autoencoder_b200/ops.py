"""Tensor-facing wrappers over the C ABI: shape validation (the reference's
OP_REQUIRES conditions and messages), output allocation and the launch on torch's
current stream.  torch is plumbing here (device memory + streams); all arithmetic
happens in libpnae.so.  No CPU path: a non-CUDA tensor is an error.

Reference glue being mirrored:
  NnDistanceGpuOp / NnDistanceGradGpuOp      tf_ops/nn_distance/tf_nndistance.cpp:169-254
  ApproxMatchGpuOp / MatchCostGpuOp / MatchCostGradGpuOp
                                             tf_ops/approxmatch/tf_approxmatch.cpp:145-295
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

NUM_LEVELS = _lib.NUM_LEVELS


def _require(cond, msg):
    if not cond:
        raise ValueError(msg)     # the reference raises errors::InvalidArgument(msg)


def _dev(t, name):
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor" % name)
    if not t.is_cuda:
        raise RuntimeError("%s is on %s: pointnet_autoencoder_b200 has no CPU path (CUDA tensors only)" % (name, t.device))
    return t


def _f32c(t):
    if t.dtype != torch.float32:
        raise TypeError("expected float32, got %s" % t.dtype)
    return t.contiguous()


def _i32c(t):
    if t.dtype != torch.int32:
        raise TypeError("expected int32, got %s" % t.dtype)
    return t.contiguous()


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream(t):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _check_pair(op, xyz1, xyz2, style):
    _dev(xyz1, "xyz1"); _dev(xyz2, "xyz2")
    if style == "nn":
        _require(xyz1.dim() == 3, "%s requires xyz1 be of shape (batch,#points,3)" % op)
        _require(xyz1.shape[2] == 3, "%s only accepts 3d point set xyz1" % op)
        _require(xyz2.dim() == 3, "%s requires xyz2 be of shape (batch,#points,3)" % op)
        _require(xyz2.shape[2] == 3, "%s only accepts 3d point set xyz2" % op)
        _require(xyz2.shape[0] == xyz1.shape[0], "%s expects xyz1 and xyz2 have same batch size" % op)
    else:
        _require(xyz1.dim() == 3 and xyz1.shape[2] == 3, "%s expects (batch_size,num_points,3) xyz1 shape" % op)
        _require(xyz2.dim() == 3 and xyz2.shape[2] == 3 and xyz2.shape[0] == xyz1.shape[0],
                 "%s expects (batch_size,num_points,3) xyz2 shape, and batch_size must match" % op)
    _require(xyz1.shape[1] >= 1 and xyz2.shape[1] >= 1, "%s needs at least one point per cloud" % op)
    _require(xyz1.device == xyz2.device, "%s expects xyz1 and xyz2 on the same device" % op)
    return _f32c(xyz1), _f32c(xyz2), xyz1.shape[0], xyz1.shape[1], xyz2.shape[1]


# ---------------------------------------------------------------------------
# Chamfer
# ---------------------------------------------------------------------------
def nn_distance_fwd(xyz1, xyz2):
    """NnDistance: -> dist1 (B,N) f32, idx1 (B,N) i32, dist2 (B,M) f32, idx2 (B,M) i32"""
    xyz1, xyz2, b, n, m = _check_pair("NnDistance", xyz1, xyz2, "nn")
    lib = _lib.load()
    dev = xyz1.device
    with torch.cuda.device(dev):
        dist1 = torch.empty((b, n), dtype=torch.float32, device=dev)
        idx1 = torch.empty((b, n), dtype=torch.int32, device=dev)
        dist2 = torch.empty((b, m), dtype=torch.float32, device=dev)
        idx2 = torch.empty((b, m), dtype=torch.int32, device=dev)
        wsb = lib.pnae_nn_distance_workspace_bytes(b, n, m)
        ws = torch.empty((wsb,), dtype=torch.uint8, device=dev) if wsb else None
        _lib.check(lib.pnae_nn_distance_fwd(b, n, _p(xyz1), m, _p(xyz2), _p(dist1), _p(idx1), _p(dist2), _p(idx2),
                                            _p(ws), wsb, _stream(xyz1)))
    return dist1, idx1, dist2, idx2


def nn_distance_bwd(xyz1, xyz2, grad_dist1, idx1, grad_dist2, idx2):
    """NnDistanceGrad: -> grad_xyz1 (B,N,3), grad_xyz2 (B,M,3)"""
    op = "NnDistanceGrad"
    xyz1, xyz2, b, n, m = _check_pair(op, xyz1, xyz2, "nn")
    _require(tuple(grad_dist1.shape) == (b, n), "%s requires grad_dist1 be of shape(batch,#points)" % op)
    _require(tuple(idx1.shape) == (b, n), "%s requires idx1 be of shape(batch,#points)" % op)
    _require(tuple(grad_dist2.shape) == (b, m), "%s requires grad_dist2 be of shape(batch,#points)" % op)
    _require(tuple(idx2.shape) == (b, m), "%s requires idx2 be of shape(batch,#points)" % op)
    g1 = _f32c(_dev(grad_dist1, "grad_dist1")); g2 = _f32c(_dev(grad_dist2, "grad_dist2"))
    i1 = _i32c(_dev(idx1, "idx1")); i2 = _i32c(_dev(idx2, "idx2"))
    lib = _lib.load()
    dev = xyz1.device
    with torch.cuda.device(dev):
        o1 = torch.empty((b, n, 3), dtype=torch.float32, device=dev)
        o2 = torch.empty((b, m, 3), dtype=torch.float32, device=dev)
        _lib.check(lib.pnae_nn_distance_bwd(b, n, _p(xyz1), m, _p(xyz2), _p(g1), _p(i1), _p(g2), _p(i2),
                                            _p(o1), _p(o2), _stream(xyz1)))
    return o1, o2


def nn_distance_fwd_grad(xyz1, xyz2, grad_dist1, grad_dist2):
    """NnDistance + NnDistanceGrad in one call (two launches) for upstream gradients known beforehand:
    -> dist1, idx1, dist2, idx2, grad_xyz1 (B,N,3), grad_xyz2 (B,M,3)"""
    op = "NnDistanceGrad"
    xyz1, xyz2, b, n, m = _check_pair("NnDistance", xyz1, xyz2, "nn")
    _require(tuple(grad_dist1.shape) == (b, n), "%s requires grad_dist1 be of shape(batch,#points)" % op)
    _require(tuple(grad_dist2.shape) == (b, m), "%s requires grad_dist2 be of shape(batch,#points)" % op)
    g1 = _f32c(_dev(grad_dist1, "grad_dist1")); g2 = _f32c(_dev(grad_dist2, "grad_dist2"))
    lib = _lib.load()
    dev = xyz1.device
    with torch.cuda.device(dev):
        dist1 = torch.empty((b, n), dtype=torch.float32, device=dev); idx1 = torch.empty((b, n), dtype=torch.int32, device=dev)
        dist2 = torch.empty((b, m), dtype=torch.float32, device=dev); idx2 = torch.empty((b, m), dtype=torch.int32, device=dev)
        o1 = torch.empty((b, n, 3), dtype=torch.float32, device=dev); o2 = torch.empty((b, m, 3), dtype=torch.float32, device=dev)
        wsb = lib.pnae_nn_distance_workspace_bytes(b, n, m)
        ws = torch.empty((max(wsb, 1),), dtype=torch.uint8, device=dev)
        _lib.check(lib.pnae_nn_distance_fwd_grad(b, n, _p(xyz1), m, _p(xyz2), _p(g1), _p(g2), _p(dist1), _p(idx1), _p(dist2), _p(idx2),
                                                 _p(o1), _p(o2), _p(ws), wsb, _stream(xyz1)))
    return dist1, idx1, dist2, idx2, o1, o2


def chamfer_loss_grad(xyz1, xyz2, w1, w2):
    """Fused: loss = w1*sum(dist1) + w2*sum(dist2) -> (loss (), grad_xyz1 (B,N,3), grad_xyz2 (B,M,3)); two launches"""
    xyz1, xyz2, b, n, m = _check_pair("NnDistance", xyz1, xyz2, "nn")
    lib = _lib.load()
    dev = xyz1.device
    with torch.cuda.device(dev):
        loss = torch.empty((), dtype=torch.float32, device=dev)
        g1 = torch.empty((b, n, 3), dtype=torch.float32, device=dev)
        g2 = torch.empty((b, m, 3), dtype=torch.float32, device=dev)
        wsb = lib.pnae_nn_distance_workspace_bytes(b, n, m)
        ws = torch.empty((max(wsb, 1),), dtype=torch.uint8, device=dev)
        _lib.check(lib.pnae_chamfer_loss_grad(b, n, _p(xyz1), m, _p(xyz2), float(w1), float(w2), _p(loss), _p(g1), _p(g2),
                                              None, None, None, None, _p(ws), wsb, _stream(xyz1)))
    return loss, g1, g2


# ---------------------------------------------------------------------------
# approximate EMD
# ---------------------------------------------------------------------------
def approx_match_factors(xyz1, xyz2, dense=False):
    """ApproxMatch: -> factors (B,10,N+M) [, match (B,M,N) if dense]"""
    xyz1, xyz2, b, n, m = _check_pair("ApproxMatch", xyz1, xyz2, "emd")
    lib = _lib.load()
    dev = xyz1.device
    with torch.cuda.device(dev):
        factors = torch.empty((b, NUM_LEVELS, n + m), dtype=torch.float32, device=dev)
        match = torch.empty((b, m, n), dtype=torch.float32, device=dev) if dense else None
        wsb = lib.pnae_approx_match_workspace_bytes(b, n, m)
        ws = torch.empty((max(wsb, 1),), dtype=torch.uint8, device=dev)
        _lib.check(lib.pnae_approx_match(b, n, m, _p(xyz1), _p(xyz2), _p(factors), _p(match), _p(ws), wsb, _stream(xyz1)))
    return (factors, match) if dense else factors


def match_from_factors(xyz1, xyz2, factors):
    xyz1, xyz2, b, n, m = _check_pair("ApproxMatch", xyz1, xyz2, "emd")
    _require(tuple(factors.shape) == (b, NUM_LEVELS, n + m), "factors must be (batch_size,10,#dataset+#query)")
    factors = _f32c(_dev(factors, "factors"))
    lib = _lib.load()
    dev = xyz1.device
    with torch.cuda.device(dev):
        match = torch.empty((b, m, n), dtype=torch.float32, device=dev)
        _lib.check(lib.pnae_match_from_factors(b, n, m, _p(xyz1), _p(xyz2), _p(factors), _p(match), _stream(xyz1)))
    return match


def _check_match(op, match, b, n, m):
    _dev(match, "match")
    _require(match.dim() == 3 and match.shape[0] == b and match.shape[1] == m and match.shape[2] == n,
             "%s expects (batch_size,#query,#dataset) match shape" % op)
    return _f32c(match)


def match_cost_dense_fwd(xyz1, xyz2, match):
    """MatchCost over a dense (B,M,N) match: -> cost (B,)"""
    xyz1, xyz2, b, n, m = _check_pair("MatchCost", xyz1, xyz2, "emd")
    match = _check_match("MatchCost", match, b, n, m)
    lib = _lib.load()
    dev = xyz1.device
    with torch.cuda.device(dev):
        cost = torch.empty((b,), dtype=torch.float32, device=dev)
        _lib.check(lib.pnae_match_cost_fwd(b, n, m, _p(xyz1), _p(xyz2), _p(match), _p(cost), _stream(xyz1)))
    return cost


def match_cost_dense_bwd(xyz1, xyz2, match):
    """MatchCostGrad over a dense match: -> grad1 (B,N,3), grad2 (B,M,3) (unscaled)"""
    xyz1, xyz2, b, n, m = _check_pair("MatchCostGrad", xyz1, xyz2, "emd")
    match = _check_match("MatchCost", match, b, n, m)     # sic: the reference's message says MatchCost (tf_approxmatch.cpp:280)
    lib = _lib.load()
    dev = xyz1.device
    with torch.cuda.device(dev):
        g1 = torch.empty((b, n, 3), dtype=torch.float32, device=dev)
        g2 = torch.empty((b, m, 3), dtype=torch.float32, device=dev)
        _lib.check(lib.pnae_match_cost_bwd(b, n, m, _p(xyz1), _p(xyz2), _p(match), _p(g1), _p(g2), _stream(xyz1)))
    return g1, g2


def match_cost_factors(xyz1, xyz2, factors, with_grad=True):
    """MatchCost (+MatchCostGrad) straight from the factors: -> cost[, grad1, grad2]"""
    xyz1, xyz2, b, n, m = _check_pair("MatchCost", xyz1, xyz2, "emd")
    _require(tuple(factors.shape) == (b, NUM_LEVELS, n + m), "MatchCost expects (batch_size,10,#dataset+#query) factors")
    factors = _f32c(_dev(factors, "factors"))
    lib = _lib.load()
    dev = xyz1.device
    with torch.cuda.device(dev):
        cost = torch.empty((b,), dtype=torch.float32, device=dev)
        g1 = torch.empty((b, n, 3), dtype=torch.float32, device=dev) if with_grad else None
        g2 = torch.empty((b, m, 3), dtype=torch.float32, device=dev) if with_grad else None
        _lib.check(lib.pnae_match_cost_factors(b, n, m, _p(xyz1), _p(xyz2), _p(factors), _p(cost), _p(g1), _p(g2), _stream(xyz1)))
    return (cost, g1, g2) if with_grad else cost


# ---------------------------------------------------------------------------
# encoder: conv5 + max-pool (models/model.py:57-66)
# ---------------------------------------------------------------------------
# flags of the encoder entry points (include/pnae.h)
STATS_ZEROED = 1
OVERLAP_PREVIOUS = 2


def _overlap_flag(overlap, *pairs):
    """OVERLAP_PREVIOUS if asked for and none of the (given, used) parameter tensors had to be converted: a conversion
    is a kernel that writes the parameter immediately before the call, which is what the flag's contract excludes"""
    if not overlap:
        return 0
    return OVERLAP_PREVIOUS if all(a.data_ptr() == b.data_ptr() for a, b in pairs) else 0


def encoder_conv_pool(x_bf16, wt_bf16, sign=None, overlap=False):
    """x (B,N,K) bf16, wt (C,K) bf16 -> max, min, sum, sumsq of x @ wt.T over the points, each (B,C) fp32.
    With `sign` (C,) fp32 also -> arg (B,C) int32: first point attaining the max (sign>=0) / min (sign<0).
    overlap: PNAE_OVERLAP_PREVIOUS -- the kernel enqueued just before this call does not write `wt_bf16`."""
    _dev(x_bf16, "x"); _dev(wt_bf16, "wt")
    _require(x_bf16.dim() == 3 and wt_bf16.dim() == 2 and x_bf16.shape[2] == wt_bf16.shape[1],
             "encoder_conv_pool expects x (batch,#points,k) and wt (c,k)")
    if x_bf16.dtype != torch.bfloat16 or wt_bf16.dtype != torch.bfloat16:
        raise TypeError("encoder_conv_pool expects bfloat16 operands")
    x = x_bf16.contiguous(); wt = wt_bf16.contiguous()
    b, n, k = x.shape
    c = wt.shape[0]
    lib = _lib.load()
    dev = x.device
    with torch.cuda.device(dev):
        outs = [torch.empty((b, c), dtype=torch.float32, device=dev) for _ in range(4)]
        arg = None
        if sign is not None:
            _require(tuple(sign.shape) == (c,), "encoder_conv_pool expects sign of shape (c,)")
            sign = _f32c(_dev(sign, "sign"))
            arg = torch.empty((b, c), dtype=torch.int32, device=dev)
        _lib.check(lib.pnae_encoder_conv_pool(b, n, k, c, _p(x), _p(wt), _p(outs[0]), _p(outs[1]), _p(outs[2]), _p(outs[3]),
                                              _p(sign), _p(arg), _overlap_flag(overlap, (wt_bf16, wt)), _stream(x)))
    return tuple(outs) if arg is None else tuple(outs) + (arg,)


# ---------------------------------------------------------------------------
# encoder layers 1-4 (csrc/shared_mlp.cu)
# ---------------------------------------------------------------------------
def _stats_arg(stats_out, words, dev):
    """the statistics buffer of a layer call: a fresh one (the call zeroes it) or the caller's zeroed view (STATS_ZEROED)"""
    if stats_out is None:
        return torch.empty((words,), dtype=torch.float32, device=dev), 0
    _require(stats_out.dtype == torch.float32 and stats_out.is_contiguous() and stats_out.numel() == words and stats_out.device == dev,
             "stats_out must be a contiguous zeroed fp32 tensor of %d words on the input's device" % words)
    return stats_out, STATS_ZEROED


def mlp_stats_words(kout):
    """words of the statistics buffer of an mlp_layer call with kout output channels: sums, sums of squares, tile counters"""
    return 2 * kout + kout // 64


def mlp_first(xyz, w, bias, stats_out=None, overlap=False):
    """layer 1: xyz (B,N,3) fp32, w (3,64), bias (64,) -> raw output (B*N,64) fp32, stats (2,64) = per-channel sum / sum of squares.
    stats_out: a zeroed (128,) view to accumulate into (saves the call's own memset); overlap: PNAE_OVERLAP_PREVIOUS."""
    _dev(xyz, "xyz")
    _require(xyz.dim() == 3 and xyz.shape[2] == 3 and tuple(w.shape) == (3, 64) and tuple(bias.shape) == (64,), "mlp_first expects xyz (batch,#points,3), w (3,64), bias (64,)")
    w0, bias0 = w, bias
    x = _f32c(xyz); w = _f32c(_dev(w, "w")); bias = _f32c(_dev(bias, "bias"))
    npts = x.shape[0] * x.shape[1]
    lib = _lib.load()
    with torch.cuda.device(x.device):
        out = torch.empty((npts, 64), dtype=torch.float32, device=x.device)
        stats, zf = _stats_arg(stats_out, 128, x.device)
        _lib.check(lib.pnae_mlp_first(npts, _p(x), _p(w), _p(bias), _p(out), _p(stats), zf | _overlap_flag(overlap, (w0, w), (bias0, bias)), _stream(x)))
    return out, stats.view(2, 64)


def xyz_moments(xyz, out=None, overlap=False):
    """xyz (B,N,3) fp32 -> (9,) float64: sums of x, y, z, xx, xy, xz, yy, yz, zz over all points (what layer 1's BatchNorm
    statistics follow from, see mlp_layer_xyz).  out: a zeroed (9,) float64 view to accumulate into."""
    _dev(xyz, "xyz")
    _require(xyz.dim() == 3 and xyz.shape[2] == 3, "xyz_moments expects xyz (batch,#points,3)")
    x = _f32c(xyz)
    lib = _lib.load()
    with torch.cuda.device(x.device):
        if out is None:
            mom, zf = torch.empty((9,), dtype=torch.float64, device=x.device), 0
        else:
            _require(out.dtype == torch.float64 and out.is_contiguous() and out.numel() == 9 and out.device == x.device,
                     "out must be a contiguous zeroed float64 tensor of 9 elements on the input's device")
            mom, zf = out, STATS_ZEROED
        _lib.check(lib.pnae_xyz_moments(x.shape[0] * x.shape[1], _p(x), _p(mom), zf | (OVERLAP_PREVIOUS if overlap else 0), _stream(x)))
    return mom


def mlp_layer_xyz(xyz, moments, w1, b1, gamma1, beta1, moving_mean1, moving_var1, training, decay, eps, w, bias,
                  stats_out=None, overlap=False):
    """layers 1 and 2 in one kernel: relu(BatchNorm1(xyz @ w1 + b1)) @ w + bias -> raw output (B*N,kout) fp32, stats (2,kout),
    with layer 1's batch statistics derived from `moments` (xyz_moments); the (B*N,64) tensor of layer 1 is never formed.
    moving_mean1 / moving_var1 are updated in place when training."""
    _dev(xyz, "xyz")
    _require(xyz.dim() == 3 and xyz.shape[2] == 3 and tuple(w1.shape) == (3, 64) and tuple(b1.shape) == (64,) and w.shape[0] == 64
             and tuple(bias.shape) == (w.shape[1],), "mlp_layer_xyz expects xyz (batch,#points,3), w1 (3,64), b1 (64,), w (64,kout), bias (kout,)")
    kout = w.shape[1]
    g, b = _bn_args(64, None, gamma1, beta1, moving_mean1, moving_var1, False)
    if training:
        _require(moments is not None and moments.dtype == torch.float64 and moments.numel() == 9 and moments.is_contiguous(),
                 "mlp_layer_xyz needs the (9,) float64 moments of xyz in training mode")
    ws = (w1, b1, w, bias)
    x = _f32c(xyz); w1c = _f32c(_dev(w1, "w1")); b1c = _f32c(_dev(b1, "b1")); wc = _f32c(_dev(w, "w")); bc = _f32c(_dev(bias, "bias"))
    npts = x.shape[0] * x.shape[1]
    lib = _lib.load()
    with torch.cuda.device(x.device):
        out = torch.empty((npts, kout), dtype=torch.float32, device=x.device)
        buf, zf = _stats_arg(stats_out, 2 * kout + kout // 64, x.device)
        _lib.check(lib.pnae_mlp_layer_xyz(npts, _p(x), _p(moments) if moments is not None else None, _p(w1c), _p(b1c), _p(g), _p(b),
                                          _p(moving_mean1), _p(moving_var1), float(eps), float(decay), int(bool(training)), kout,
                                          _p(wc), _p(bc), _p(out), _p(buf),
                                          zf | _overlap_flag(overlap, *zip(ws, (w1c, b1c, wc, bc))), _stream(x)))
    return out, buf[: 2 * kout].view(2, kout)


def _bn_args(k, stats, gamma, beta, moving_mean, moving_var, training):
    _require(tuple(gamma.shape) == (k,) and tuple(beta.shape) == (k,) and tuple(moving_mean.shape) == (k,) and tuple(moving_var.shape) == (k,),
             "BatchNorm parameters must have one entry per input channel")
    _require(moving_mean.is_contiguous() and moving_var.is_contiguous() and moving_mean.dtype == torch.float32 and moving_var.dtype == torch.float32,
             "moving statistics must be contiguous float32 (they are updated in place)")
    _require((not training) or (stats is not None and stats.is_contiguous() and tuple(stats.shape) == (2, k)), "training mode needs the previous layer's stats (2,k)")
    return _f32c(_dev(gamma, "gamma")), _f32c(_dev(beta, "beta"))


def mlp_layer(y_prev, stats_prev, gamma_prev, beta_prev, moving_mean_prev, moving_var_prev, training, decay, eps, w, bias,
              stats_out=None, overlap=False):
    """layers 2-4: relu(BatchNorm_prev(y_prev)) @ w + bias -> raw output (T,kout) fp32, stats (2,kout).
    BatchNorm_prev uses batch statistics formed from stats_prev (training; the moving statistics are updated in place,
    TF convention) or the moving statistics (inference).
    stats_out: a zeroed (mlp_stats_words(kout),) view to accumulate into; overlap: PNAE_OVERLAP_PREVIOUS."""
    _dev(y_prev, "y_prev")
    kin, kout = w.shape
    _require(y_prev.dim() == 2 and y_prev.shape[1] == kin and tuple(bias.shape) == (kout,), "mlp_layer expects y_prev (points,kin), w (kin,kout), bias (kout,)")
    g, b = _bn_args(kin, stats_prev, gamma_prev, beta_prev, moving_mean_prev, moving_var_prev, training)
    w0, bias0 = w, bias
    y = _f32c(y_prev); w = _f32c(_dev(w, "w")); bias = _f32c(_dev(bias, "bias"))
    lib = _lib.load()
    with torch.cuda.device(y.device):
        out = torch.empty((y.shape[0], kout), dtype=torch.float32, device=y.device)
        buf, zf = _stats_arg(stats_out, 2 * kout + kout // 64, y.device)                       # statistics + the kernel's tile counters
        _lib.check(lib.pnae_mlp_layer(y.shape[0], kin, kout, _p(y), _p(stats_prev), _p(g), _p(b), _p(moving_mean_prev), _p(moving_var_prev),
                                      float(eps), float(decay), int(bool(training)), _p(w), _p(bias), _p(out), _p(buf),
                                      zf | _overlap_flag(overlap, (w0, w), (bias0, bias)), _stream(y)))
    return out, buf[: 2 * kout].view(2, kout)


def bn_fold(stats, count, gamma, beta, moving_mean, moving_var, training, decay, eps):
    """-> (s, t): folded BatchNorm scale / shift of one layer (batch statistics from `stats` and a moving-average update when
    training, moving statistics otherwise); one launch"""
    k = gamma.shape[0]
    g = _f32c(_dev(gamma, "gamma")); b = _f32c(_dev(beta, "beta"))
    _require(moving_mean.is_contiguous() and moving_var.is_contiguous() and moving_mean.dtype == torch.float32, "bn_fold expects contiguous fp32 moving statistics")
    lib = _lib.load()
    with torch.cuda.device(g.device):
        s = torch.empty((k,), dtype=torch.float32, device=g.device); t = torch.empty((k,), dtype=torch.float32, device=g.device)
        _lib.check(lib.pnae_bn_fold(k, _p(stats) if stats is not None else None, float(count), _p(g), _p(b), float(eps), float(decay), int(bool(training)),
                                    _p(moving_mean), _p(moving_var), _p(s), _p(t), _stream(g)))
    return s, t


def mlp_apply_bf16(y, stats, gamma, beta, moving_mean, moving_var, training, decay, eps, overlap=False):
    """relu(BatchNorm(y)) -> bf16 (T,k): the K-major operand of encoder_conv_pool (BatchNorm given as in mlp_layer)"""
    _dev(y, "y")
    k = y.shape[1]
    _require(y.dim() == 2 and k % 4 == 0 and k <= 256, "mlp_apply_bf16 expects y (points,k), k % 4 == 0, k <= 256")
    g, b = _bn_args(k, stats, gamma, beta, moving_mean, moving_var, training)
    y = _f32c(y)
    lib = _lib.load()
    with torch.cuda.device(y.device):
        out = torch.empty((y.shape[0], k), dtype=torch.bfloat16, device=y.device)
        _lib.check(lib.pnae_mlp_apply_bf16(y.shape[0], k, _p(y), _p(stats), _p(g), _p(b), _p(moving_mean), _p(moving_var),
                                           float(eps), float(decay), int(bool(training)), _p(out), OVERLAP_PREVIOUS if overlap else 0, _stream(y)))
    return out


def conv5_finish(vmax, vmin, vsum, vsq, count, bias, gamma, beta, moving_mean, moving_var, training, decay, eps, overlap=False):
    """conv5's bias + BatchNorm + ReLU + max-pool finish on (B,C) in one launch -> pooled, inv (C), mean0 (C), ext0 (B,C), z (B,C)"""
    b, c = vmax.shape
    g, be = _bn_args(c, None, gamma, beta, moving_mean, moving_var, False)
    bias = _f32c(_dev(bias, "bias"))
    lib = _lib.load()
    dev = vmax.device
    with torch.cuda.device(dev):
        f = dict(dtype=torch.float32, device=dev)
        pooled = torch.empty((b, c), **f); inv = torch.empty((c,), **f); mean0 = torch.empty((c,), **f)
        ext0 = torch.empty((b, c), **f); z = torch.empty((b, c), **f)
        _lib.check(lib.pnae_conv5_finish(b, c, float(count), _p(_f32c(vmax)), _p(_f32c(vmin)), _p(_f32c(vsum)), _p(_f32c(vsq)), _p(bias), _p(g), _p(be),
                                         _p(moving_mean), _p(moving_var), float(eps), float(decay), int(bool(training)),
                                         _p(pooled), _p(inv), _p(mean0), _p(ext0), _p(z), OVERLAP_PREVIOUS if overlap else 0, _stream(vmax)))
    return pooled, inv, mean0, ext0, z
