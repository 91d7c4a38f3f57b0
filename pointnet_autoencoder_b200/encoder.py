"""PointNet encoder of the reference's `get_model` (models/model.py:40-66, identical in all six
model files): per point 3 -> 64 -> 64 -> 64 -> 128 -> 1024, each layer
`tf_util.conv2d` = 1x1 conv (a per-point linear map) + bias + BatchNorm + ReLU
(utils/tf_util.py:155-185), then `tf_util.max_pool2d` over the points (:368-391).

Layers 1-4 (11% of the FLOPs, memory-bound) are one kernel per layer (csrc/shared_mlp.cu: the previous layer's
BatchNorm + ReLU applied on the way in, 3xTF32 tensor-core product, bias, raw output, BatchNorm statistics in the
epilogue), the last one emitting the bf16 operand of layer 5 directly.  Layer 5 + bias + BN +
ReLU + max-pool is the fused tcgen05 kernel `pnae_encoder_conv_pool` (csrc/encoder.cu): the
(B, N, 1024) activation is never written; the kernel returns per-(batch, channel) max / min / sum /
sum-of-squares of the raw GEMM output and the rest finishes on a (B, 1024) tensor:

    BN scale s = gamma / sqrt(var + eps),  shift t = beta - mean * s      (batch or running statistics)
    max_n relu(s*y_n + t) = relu(s * max_n y_n + t)   if s >= 0,   relu(s * min_n y_n + t)   if s < 0

TF semantics kept: bias is added before BN (tf_util.py:176-181), BN eps = 1e-3 and biased batch
variance (tf.contrib.layers.batch_norm defaults, tf_util.py:529-533), `bn_decay` is TF's decay
(moving = decay*moving + (1-decay)*batch), Xavier-uniform weights / zero bias (tf_util.py:42,174).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops

BN_EPS = 1e-3


class _Conv5Pool(torch.autograd.Function):
    """pooled = max_n relu(BN(x @ w + bias)), forward on tensor cores (csrc/encoder.cu), never forming
    the (B, N, C) activation -- in the backward either.

    Backward.  Only the arg-extremum point n*(b,c) receives the pooled gradient, but training-mode BN
    couples all points through its statistics:
        dy[b,n,c] = s_c * ( dz[b,n,c] - dbeta_c/T - xhat[b,n,c] * dgamma_c/T ),   T = B*N,
        dz[b,n,c] = g[b,c] * [z* > 0] * delta(n, n*(b,c)).
    With y0 = x @ w this is  dy = S + a_c + q_c * y0  (S sparse: one entry per (b,c)), hence
        dW = gather(x, n*)^T (s g~) + sum_n x (x) a + (X^T X) w diag(q)        (a K x K Gram matrix)
        dX = x (w diag(q) w^T) + w a + scatter(s g~ w^T at n*)                 (a K x K matrix)
    i.e. two GEMMs with inner/outer size K=128 instead of three with C=1024, plus gathers on (B,C)."""

    @staticmethod
    def forward(ctx, x, w, bias, gamma, beta, run_mean, run_var, training, decay):
        b, n, k = x.shape
        xb = x.detach().to(torch.bfloat16)
        wtb = w.detach().t().contiguous().to(torch.bfloat16)
        need_arg = any(ctx.needs_input_grad[:5])
        if need_arg:
            vmax, vmin, vsum, vsq, arg = ops.encoder_conv_pool(xb, wtb, sign=gamma.detach())
        else:
            vmax, vmin, vsum, vsq = ops.encoder_conv_pool(xb, wtb)
            arg = None
        cnt = float(b * n)
        if training:
            mean0 = vsum.sum(0) / cnt                          # of y0 = x @ w (no bias)
            var = (vsq.sum(0) / cnt - mean0 * mean0).clamp_min(0.0)
            with torch.no_grad():
                run_mean.mul_(decay).add_(mean0 + bias, alpha=1.0 - decay)
                run_var.mul_(decay).add_(var, alpha=1.0 - decay)
        else:
            mean0, var = run_mean - bias, run_var
        inv = torch.rsqrt(var + BN_EPS)
        s = gamma * inv
        ext0 = torch.where(gamma >= 0, vmax, vmin)             # the extremum the pool selects, of y0
        z = (ext0 - mean0) * s + beta
        pooled = F.relu(z)
        if need_arg:
            ctx.save_for_backward(x, w, bias, gamma, inv, mean0, ext0, arg, z)
        ctx.training = training
        ctx.cnt = cnt
        return pooled

    @staticmethod
    def backward(ctx, grad_pooled):
        x, w, bias, gamma, inv, mean0, ext0, arg, z = ctx.saved_tensors
        b, n, k = x.shape
        c = w.shape[1]
        s = gamma * inv
        gt = grad_pooled * (z > 0).to(grad_pooled.dtype)       # (B,C) gradient reaching z at the arg-extremum
        xhat_star = (ext0 - mean0) * inv
        dbeta = gt.sum(0)
        dgamma = (gt * xhat_star).sum(0)
        sg = gt * s                                            # sparse part of dy, one value per (b,c)
        idx = arg.long().unsqueeze(-1).expand(b, c, k)
        xstar = x.gather(1, idx)                               # (B,C,K): the selected points
        dw = torch.einsum("bck,bc->kc", xstar, sg)
        dx = torch.zeros_like(x)
        dx.scatter_add_(1, idx, sg.unsqueeze(-1) * w.t().unsqueeze(0))
        if ctx.training:
            t = ctx.cnt
            q = -s * dgamma * inv / t
            a = -s * dbeta / t + s * dgamma * mean0 * inv / t
            x2 = x.reshape(-1, k)
            gram = x2.t() @ x2                                 # (K,K)
            dw = dw + torch.outer(x2.sum(0), a) + (gram @ w) * q
            dx = dx + (x2 @ ((w * q) @ w.t()) + w @ a).view(b, n, k)
            dbias = torch.zeros_like(bias)                     # BN subtracts the mean: the bias has no effect
        else:
            dbias = sg.sum(0)
        return dx, dw, dbias, dgamma, dbeta, None, None, None, None


def _conv5_forward(ctx, xb, x_saved, w, bias, gamma, beta, run_mean, run_var, training, decay, need_arg):
    """the conv5 + BN + ReLU + max-pool forward shared by _Conv5Pool and _EncoderChain; xb (B,N,K) bf16"""
    b, n, k = xb.shape
    wtb = w.detach().t().contiguous().to(torch.bfloat16)
    if need_arg:
        vmax, vmin, vsum, vsq, arg = ops.encoder_conv_pool(xb, wtb, sign=gamma.detach())
    else:
        vmax, vmin, vsum, vsq = ops.encoder_conv_pool(xb, wtb)
        arg = None
    cnt = float(b * n)
    if training:
        mean0 = vsum.sum(0) / cnt                          # of y0 = x @ w (no bias)
        var = (vsq.sum(0) / cnt - mean0 * mean0).clamp_min(0.0)
        with torch.no_grad():
            run_mean.mul_(decay).add_(mean0 + bias, alpha=1.0 - decay)
            run_var.mul_(decay).add_(var, alpha=1.0 - decay)
    else:
        mean0, var = run_mean - bias, run_var
    inv = torch.rsqrt(var + BN_EPS)
    s = gamma * inv
    ext0 = torch.where(gamma >= 0, vmax, vmin)             # the extremum the pool selects, of y0
    z = (ext0 - mean0) * s + beta
    return F.relu(z), (inv, mean0, ext0, arg, z, cnt)


def _conv5_backward(grad_pooled, x, w, bias, gamma, inv, mean0, ext0, arg, z, cnt, training):
    """(dx, dw, dbias, dgamma, dbeta) of pooled = max_n relu(BN(x @ w + bias)); see _Conv5Pool"""
    b, n, k = x.shape
    c = w.shape[1]
    s = gamma * inv
    gt = grad_pooled * (z > 0).to(grad_pooled.dtype)       # (B,C) gradient reaching z at the arg-extremum
    xhat_star = (ext0 - mean0) * inv
    dbeta = gt.sum(0)
    dgamma = (gt * xhat_star).sum(0)
    sg = gt * s                                            # sparse part of dy, one value per (b,c)
    idx = arg.long().unsqueeze(-1).expand(b, c, k)
    xstar = x.gather(1, idx)                               # (B,C,K): the selected points
    dw = torch.einsum("bck,bc->kc", xstar, sg)
    dx = torch.zeros_like(x)
    dx.scatter_add_(1, idx, sg.unsqueeze(-1) * w.t().unsqueeze(0))
    if training:
        q = -s * dgamma * inv / cnt
        a = -s * dbeta / cnt + s * dgamma * mean0 * inv / cnt
        x2 = x.reshape(-1, k)
        gram = x2.t() @ x2                                 # (K,K)
        dw = dw + torch.outer(x2.sum(0), a) + (gram @ w) * q
        dx = dx + (x2 @ ((w * q) @ w.t()) + w @ a).view(b, n, k)
        dbias = torch.zeros_like(bias)                     # BN subtracts the mean: the bias has no effect
    else:
        dbias = sg.sum(0)
    return dx, dw, dbias, dgamma, dbeta


class _EncoderChain(torch.autograd.Function):
    """The whole encoder, (B,N,3) -> (B,1024): layers 1-4 through csrc/shared_mlp.cu (one kernel per layer, no
    normalised activation in HBM, the bf16 operand of conv5 emitted directly), conv5 + BN + ReLU + max-pool through
    csrc/encoder.cu.  Arguments after `decay`: for each of the five layers weight, bias, gamma, beta, moving_mean,
    moving_var.

    Backward: conv5 analytically (see _Conv5Pool); layers 1-4 by re-running their plain library formulation under
    autograd (they are 11 % of the FLOPs, and only their forward is on the measured path)."""

    @staticmethod
    def forward(ctx, x, training, decay, grad_mode, *tensors):
        # grad_mode: torch.is_grad_enabled() at the call (inside forward() it is always off, and needs_input_grad does not
        # tell): without a backward to come the cheaper kernel variant that does not track the arg-extremum is enough
        b, n, _ = x.shape
        layers = [tensors[6 * i: 6 * i + 6] for i in range(5)]
        cnt = float(b * n)

        # Everything that is not one of the seven chain kernels first: conv5's bf16 weights and ONE zeroed arena for the
        # layers' statistics.  The seven then follow each other directly and each is enqueued as a programmatic
        # dependent launch: its set-up (weights to shared memory, barriers, tensor memory) runs under its predecessor's tail.
        w5, b5, g5, be5, mm5, mv5 = layers[4]
        wtb = w5.detach().t().contiguous().to(torch.bfloat16)
        # arena: the 9 float64 moments of xyz first (18 words), then the statistics of layers 2-4
        words = [18] + [ops.mlp_stats_words(lay[0].shape[1]) for lay in layers[1:4]]
        offs = [0]
        for wd in words:
            offs.append(offs[-1] + (wd + 3) // 4 * 4)
        arena = torch.zeros(offs[-1], dtype=torch.float32, device=x.device)
        view = lambda i: arena[offs[i]: offs[i] + words[i]]
        # layer 1 is folded into layer 2's kernel: its BatchNorm statistics follow from the moments of xyz, and its
        # (B*N,64) output is never written
        l1, l2 = layers[0], layers[1]
        mom = ops.xyz_moments(x.detach(), out=view(0).view(torch.float64), overlap=True) if training else None
        y, st = ops.mlp_layer_xyz(x.detach(), mom, l1[0].detach(), l1[1].detach(), l1[2].detach(), l1[3].detach(), l1[4], l1[5],
                                  training, decay, BN_EPS, l2[0].detach(), l2[1].detach(), stats_out=view(1), overlap=True)
        prev = l2
        for i, lay in enumerate(layers[2:4], start=2):
            # the previous layer's BatchNorm (+ moving-average update) and ReLU happen inside this layer's kernel
            y, st_next = ops.mlp_layer(y, st, prev[2].detach(), prev[3].detach(), prev[4], prev[5], training, decay, BN_EPS,
                                       lay[0].detach(), lay[1].detach(), stats_out=view(i), overlap=True)
            st, prev = st_next, lay
        xb = ops.mlp_apply_bf16(y, st, prev[2].detach(), prev[3].detach(), prev[4], prev[5], training, decay, BN_EPS, overlap=True).view(b, n, -1)
        need_arg = bool(grad_mode) and any(ctx.needs_input_grad)
        if need_arg:
            vmax, vmin, vsum, vsq, arg = ops.encoder_conv_pool(xb, wtb, sign=g5.detach(), overlap=True)
        else:
            vmax, vmin, vsum, vsq = ops.encoder_conv_pool(xb, wtb, overlap=True)
            arg = None
        pooled, inv, mean0, ext0, z = ops.conv5_finish(vmax, vmin, vsum, vsq, cnt, b5.detach(), g5.detach(), be5.detach(), mm5, mv5,
                                                       training, decay, BN_EPS, overlap=True)
        cnt5 = cnt
        if need_arg:
            ctx.save_for_backward(x, xb, inv, mean0, ext0, arg, z, *tensors)
        ctx.training, ctx.cnt = training, cnt5
        return pooled

    @staticmethod
    def backward(ctx, grad_pooled):
        x, xb, inv, mean0, ext0, arg, z = ctx.saved_tensors[:7]
        tensors = ctx.saved_tensors[7:]
        layers = [tensors[6 * i: 6 * i + 6] for i in range(5)]
        w5, b5, g5, be5 = layers[4][:4]
        dx4, dw5, db5, dg5, dbe5 = _conv5_backward(grad_pooled, xb.float(), w5, b5, g5, inv, mean0, ext0, arg, z, ctx.cnt, ctx.training)
        # layers 1-4: the plain formulation under autograd (statistics are recomputed, moving averages untouched)
        with torch.enable_grad():
            xin = x.detach().requires_grad_(ctx.needs_input_grad[0])
            params = [[p_.detach().requires_grad_(True) for p_ in lay[:4]] for lay in layers[:4]]
            net = xin.reshape(-1, 3)
            for (w_, bias_, gamma_, beta_), lay in zip(params, layers[:4]):
                yy = torch.addmm(bias_, net, w_)
                yy = F.batch_norm(yy, None if ctx.training else lay[4], None if ctx.training else lay[5], gamma_, beta_, ctx.training, 0.0, BN_EPS)
                net = F.relu(yy)
            wanted = [p_ for lay in params for p_ in lay] + ([xin] if ctx.needs_input_grad[0] else [])
            grads = torch.autograd.grad(net, wanted, grad_outputs=dx4.reshape(-1, net.shape[1]), allow_unused=True)
        out = [grads[-1] if ctx.needs_input_grad[0] else None, None, None, None]
        for i in range(4):
            out += list(grads[4 * i: 4 * i + 4]) + [None, None]
        out += [dw5, db5, dg5, dbe5, None, None]
        return tuple(out)


class SharedMLPLayer(nn.Module):
    """tf_util.conv2d with a [1,1] (or [1,3] first-layer) kernel = per-point linear + bias + BN + ReLU."""

    def __init__(self, cin, cout):
        super().__init__()
        bound = math.sqrt(6.0 / (cin + cout))                 # xavier_initializer (uniform), tf_util.py:42
        self.weight = nn.Parameter(torch.empty(cin, cout).uniform_(-bound, bound))
        self.bias = nn.Parameter(torch.zeros(cout))
        self.gamma = nn.Parameter(torch.ones(cout))
        self.beta = nn.Parameter(torch.zeros(cout))
        self.register_buffer("moving_mean", torch.zeros(cout))
        self.register_buffer("moving_var", torch.ones(cout))

    def forward(self, x, bn_decay=0.9):
        b, n, _ = x.shape
        y = torch.addmm(self.bias, x.reshape(b * n, -1), self.weight)
        # library batch-norm kernel (fused statistics + normalise, fused backward).  Its moving variance
        # uses the unbiased estimate (x T/(T-1), T = B*N >= 65536 here: 1.5e-5 relative to TF's biased one).
        y = F.batch_norm(y, self.moving_mean, self.moving_var, self.gamma, self.beta, self.training,
                         1.0 - bn_decay, BN_EPS)
        return F.relu(y).view(b, n, -1)


class PointNetEncoder(nn.Module):
    """point_cloud (B, N, 3) -> global feature (B, 1024)   (models/model.py:40-66)"""

    def __init__(self, fused=True):
        super().__init__()
        self.layers = nn.ModuleList([SharedMLPLayer(3, 64), SharedMLPLayer(64, 64), SharedMLPLayer(64, 64),
                                     SharedMLPLayer(64, 128)])
        self.conv5 = SharedMLPLayer(128, 1024)
        self.fused = fused

    def forward(self, point_cloud, bn_decay=0.9):
        if self.fused is True:
            # every layer a hand-written kernel: csrc/shared_mlp.cu (1-4) + csrc/encoder.cu (5 + pool)
            flat = []
            for layer in list(self.layers) + [self.conv5]:
                flat += [layer.weight, layer.bias, layer.gamma, layer.beta, layer.moving_mean, layer.moving_var]
            return _EncoderChain.apply(point_cloud, self.training, bn_decay, torch.is_grad_enabled(), *flat)
        net = point_cloud
        for layer in self.layers:
            net = layer(net, bn_decay)
        l5 = self.conv5
        if self.fused == "conv5":                            # round-1 form: library layers 1-4, fused conv5 + pool
            return _Conv5Pool.apply(net, l5.weight, l5.bias, l5.gamma, l5.beta, l5.moving_mean, l5.moving_var,
                                    self.training, bn_decay)
        return l5(net, bn_decay).amax(dim=1)                 # unfused reference path (library ops only)
