"""The reference's autoencoders with the B200-native loss ops on the hot path.

Only what the measured configurations need: the shared PointNet encoder
(encoder.py, fused tcgen05 conv5 + pool), the FC decoder of models/model.py:68-75, the up-conv decoder
of models/model_upconv.py:62-81, and the two losses (models/model.py:77-83, models/model_emd.py:79-89)
built on tf_nndistance / tf_approxmatch exactly as the reference's get_loss does.  Decoders are plain
library ops (torch / cuBLAS / cuDNN): they are outside the hot path.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .encoder import BN_EPS, PointNetEncoder
from .tf_ops.approxmatch import tf_approxmatch
from .tf_ops.nn_distance import tf_nndistance


class _FC(nn.Module):
    """tf_util.fully_connected (utils/tf_util.py:329-365): linear + BN + ReLU (or no activation)"""

    def __init__(self, cin, cout, bn=True, act=True):
        super().__init__()
        self.lin = nn.Linear(cin, cout)
        nn.init.xavier_uniform_(self.lin.weight); nn.init.zeros_(self.lin.bias)
        self.bn = nn.BatchNorm1d(cout, eps=BN_EPS) if bn else None
        self.act = act

    def forward(self, x, bn_decay=0.9):
        x = self.lin(x)
        if self.bn is not None:
            self.bn.momentum = 1.0 - bn_decay
            x = self.bn(x)
        return F.relu(x) if self.act else x


class _UpConv(nn.Module):
    """tf_util.conv2d_transpose (utils/tf_util.py:188-262), VALID padding: out = (in-1)*stride + k"""

    def __init__(self, cin, cout, k, s, bn=True, act=True):
        super().__init__()
        self.conv = nn.ConvTranspose2d(cin, cout, k, s)
        nn.init.xavier_uniform_(self.conv.weight); nn.init.zeros_(self.conv.bias)
        self.bn = nn.BatchNorm2d(cout, eps=BN_EPS) if bn else None
        self.act = act

    def forward(self, x, bn_decay=0.9):
        x = self.conv(x)
        if self.bn is not None:
            self.bn.momentum = 1.0 - bn_decay
            x = self.bn(x)
        return F.relu(x) if self.act else x


class AutoEncoderFC(nn.Module):
    """models/model.py get_model: encoder -> fc 1024 -> fc 1024 -> fc num_point*3"""

    def __init__(self, num_point=2048, fused_encoder=True):
        super().__init__()
        self.num_point = num_point
        self.encoder = PointNetEncoder(fused=fused_encoder)
        self.fc1 = _FC(1024, 1024); self.fc2 = _FC(1024, 1024); self.fc3 = _FC(1024, num_point * 3, bn=False, act=False)

    def forward(self, point_cloud, bn_decay=0.9):
        emb = self.encoder(point_cloud, bn_decay)
        net = self.fc2(self.fc1(emb, bn_decay), bn_decay)
        return self.fc3(net).view(-1, self.num_point, 3), {"embedding": emb}


class AutoEncoderUpconv(nn.Module):
    """models/model_upconv.py get_model (num_point must be 2048: the decoder emits a 32x64 xyz map)"""

    def __init__(self, fused_encoder=True):
        super().__init__()
        self.encoder = PointNetEncoder(fused=fused_encoder)
        self.fc00 = _FC(1024, 1024)
        self.up1 = _UpConv(512, 512, (2, 2), (2, 2)); self.up2 = _UpConv(512, 256, (3, 3), (1, 1))
        self.up3 = _UpConv(256, 256, (4, 5), (2, 3)); self.up4 = _UpConv(256, 128, (5, 7), (3, 3))
        self.up5 = _UpConv(128, 3, (1, 1), (1, 1), bn=False, act=False)

    def forward(self, point_cloud, bn_decay=0.9):
        assert point_cloud.shape[1] == 2048
        emb = self.fc00(self.encoder(point_cloud, bn_decay), bn_decay)
        net = emb.view(-1, 1, 2, 512).permute(0, 3, 1, 2)          # TF NHWC (B,1,2,512) -> NCHW
        for up in (self.up1, self.up2, self.up3, self.up4):
            net = up(net, bn_decay)
        net = self.up5(net)                                         # (B,3,32,64)
        return net.permute(0, 2, 3, 1).reshape(-1, 2048, 3), {"embedding": emb}


def nn_distance_cpu(xyz1, xyz2):
    """The reference's pure-TF Chamfer (tf_ops/nn_distance/tf_nndistance_cpu.py:4-25, used by models/model_cpu.py:81),
    restated with the same broadcast formulation: a (B,N,M,3) difference tensor, squared, summed over the last axis,
    reduce_min / argmin over either point axis (int64 indices, as tf.argmin returns).  BASELINE ONLY (configs[0]): it
    materialises B*N*M*3 floats and is not accelerated; works on any device."""
    n, m = xyz1.shape[1], xyz2.shape[1]
    a = xyz1[:, :, None, :].expand(-1, n, m, -1)
    c = xyz2[:, None, :, :].expand(-1, n, m, -1)
    d = ((a - c) ** 2).sum(-1)
    dist1, idx1 = d.min(dim=2)
    dist2, idx2 = d.min(dim=1)
    return dist1, idx1, dist2, idx2


def chamfer_loss(pred, label):
    """models/model.py:77-83 -> (loss*100, pcloss)"""
    d_fwd, _, d_bwd, _ = tf_nndistance.nn_distance(pred, label)
    loss = torch.mean(d_fwd + d_bwd)
    return loss * 100, loss


class _ChamferLossFused(torch.autograd.Function):
    """loss*100 of models/model.py:77-83 through pnae_chamfer_loss_grad: forward and both gradients in the two
    launches of the forward; backward only scales the stashed gradients by the upstream scalar."""

    @staticmethod
    def forward(ctx, pred, label):
        from . import ops
        b, n, _ = pred.shape
        m = label.shape[1]
        loss, g1, g2 = ops.chamfer_loss_grad(pred, label, 100.0 / (b * n), 100.0 / (b * m))
        ctx.save_for_backward(g1, g2)
        return loss

    @staticmethod
    def backward(ctx, grad_loss):
        g1, g2 = ctx.saved_tensors
        return (g1 * grad_loss if ctx.needs_input_grad[0] else None), (g2 * grad_loss if ctx.needs_input_grad[1] else None)


def chamfer_loss_fused(pred, label):
    """Same value and gradients as chamfer_loss(pred, label)[0] (requires n == m like the reference's
    reduce_mean(dists_forward + dists_backward)), one op instead of two, no dist/idx tensors."""
    assert pred.shape[1] == label.shape[1]
    loss = _ChamferLossFused.apply(pred, label)
    return loss, loss.detach() / 100


def emd_loss(pred, label):
    """models/model_emd.py:79-89 -> (mean match_cost, pcloss); Chamfer is still evaluated, as in the reference"""
    d_fwd, _, d_bwd, _ = tf_nndistance.nn_distance(pred, label)
    pcloss = torch.mean(d_fwd + d_bwd)
    match = tf_approxmatch.approx_match(label, pred)
    return torch.mean(tf_approxmatch.match_cost(label, pred, match)), pcloss


# ---- train.py schedules (train.py:57-60, 74-92) -------------------------------------------------
BN_INIT_DECAY, BN_DECAY_DECAY_RATE, BN_DECAY_CLIP = 0.5, 0.5, 0.99


def get_bn_decay(step, batch_size, decay_step=200000):
    bn_momentum = BN_INIT_DECAY * (BN_DECAY_DECAY_RATE ** ((step * batch_size) // float(decay_step)))
    return min(BN_DECAY_CLIP, 1 - bn_momentum)


def get_learning_rate(step, batch_size, base_lr=0.001, decay_step=200000, decay_rate=0.7):
    # staircase exponential decay; the reference's max(lr, 1e-5) clip is a no-op (misspelt variable, train.py:81)
    return base_lr * (decay_rate ** ((step * batch_size) // decay_step))
