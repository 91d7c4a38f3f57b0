"""PointNet encoder of the reference's `get_model` (models/model.py:40-66, identical in all six
model files): per point 3 -> 64 -> 64 -> 64 -> 128 -> 1024, each layer
`tf_util.conv2d` = 1x1 conv (a per-point linear map) + bias + BatchNorm + ReLU
(utils/tf_util.py:155-185), then `tf_util.max_pool2d` over the points (:368-391).

Layers 1-4 (11% of the FLOPs) run as plain library GEMMs (torch / cuBLAS).  Layer 5 + bias + BN +
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
    """pooled = max_n relu(BN(x @ w + bias)) with the forward on tensor cores; the backward
    re-evaluates the layer with library GEMMs (only the arg-max points carry gradient through the
    pool, but training-mode BN couples all points through its statistics)."""

    @staticmethod
    def forward(ctx, x, w, bias, gamma, beta, run_mean, run_var, training, decay):
        b, n, k = x.shape
        c = w.shape[1]
        vmax, vmin, vsum, vsq = ops.encoder_conv_pool(x.to(torch.bfloat16), w.t().contiguous().to(torch.bfloat16))
        if training:
            cnt = float(b * n)
            mean_acc = vsum.sum(0) / cnt                       # of the GEMM output without bias
            var = (vsq.sum(0) / cnt - mean_acc * mean_acc).clamp_min(0.0)
            mean = mean_acc + bias
            with torch.no_grad():
                run_mean.mul_(decay).add_(mean, alpha=1.0 - decay)
                run_var.mul_(decay).add_(var, alpha=1.0 - decay)
        else:
            mean, var = run_mean, run_var
        s = gamma * torch.rsqrt(var + BN_EPS)
        t = beta - mean * s
        ext = torch.where(s >= 0, vmax, vmin) + bias
        pooled = F.relu(ext * s + t)
        ctx.save_for_backward(x, w, bias, gamma, beta, run_mean, run_var)
        ctx.training = training
        return pooled

    @staticmethod
    def backward(ctx, grad_pooled):
        x, w, bias, gamma, beta, run_mean, run_var = ctx.saved_tensors
        with torch.enable_grad():
            xx = x.detach().requires_grad_(True)
            ww = w.detach().requires_grad_(True)
            bb = bias.detach().requires_grad_(True)
            gg = gamma.detach().requires_grad_(True)
            be = beta.detach().requires_grad_(True)
            y = xx @ ww + bb
            if ctx.training:
                mean = y.mean(dim=(0, 1)); var = y.var(dim=(0, 1), unbiased=False)
            else:
                mean, var = run_mean, run_var
            out = F.relu((y - mean) * torch.rsqrt(var + BN_EPS) * gg + be).amax(dim=1)
            grads = torch.autograd.grad(out, (xx, ww, bb, gg, be), grad_pooled)
        return grads[0], grads[1], grads[2], grads[3], grads[4], None, None, None, None


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
        y = x @ self.weight + self.bias
        if self.training:
            mean = y.mean(dim=(0, 1)); var = y.var(dim=(0, 1), unbiased=False)
            with torch.no_grad():
                self.moving_mean.mul_(bn_decay).add_(mean.detach(), alpha=1.0 - bn_decay)
                self.moving_var.mul_(bn_decay).add_(var.detach(), alpha=1.0 - bn_decay)
        else:
            mean, var = self.moving_mean, self.moving_var
        return F.relu((y - mean) * torch.rsqrt(var + BN_EPS) * self.gamma + self.beta)


class PointNetEncoder(nn.Module):
    """point_cloud (B, N, 3) -> global feature (B, 1024)   (models/model.py:40-66)"""

    def __init__(self, fused=True):
        super().__init__()
        self.layers = nn.ModuleList([SharedMLPLayer(3, 64), SharedMLPLayer(64, 64), SharedMLPLayer(64, 64),
                                     SharedMLPLayer(64, 128)])
        self.conv5 = SharedMLPLayer(128, 1024)
        self.fused = fused

    def forward(self, point_cloud, bn_decay=0.9):
        net = point_cloud
        for layer in self.layers:
            net = layer(net, bn_decay)
        l5 = self.conv5
        if self.fused:
            return _Conv5Pool.apply(net, l5.weight, l5.bias, l5.gamma, l5.beta, l5.moving_mean, l5.moving_var,
                                    self.training, bn_decay)
        return l5(net, bn_decay).amax(dim=1)                 # unfused reference path (library ops only)
