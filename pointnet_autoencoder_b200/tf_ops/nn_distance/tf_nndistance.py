""" Compute Chamfer's Distance -- B200-native drop-in for the reference module
tf_ops/nn_distance/tf_nndistance.py.

Same module name, same function, same argument and output order, dtypes and shapes
(tf_nndistance.py:14-24), and the same registered gradient (tf_nndistance.py:31-37):
the gradient w.r.t. xyz1/xyz2 comes from NnDistanceGrad, gradients flowing into the
two index outputs are ignored.  Tensors are torch CUDA tensors instead of TF ones;
the arithmetic runs in libpnae.so (hand-written sm_100a kernels) -- there is no
CPU path.  Like the reference module it can be imported bare after
`sys.path.append(<...>/tf_ops/nn_distance)` (models/model.py:16-17).
"""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

import torch  # noqa: E402

from pointnet_autoencoder_b200 import ops as _ops  # noqa: E402


class _NnDistance(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz1, xyz2):
        dist1, idx1, dist2, idx2 = _ops.nn_distance_fwd(xyz1, xyz2)
        ctx.save_for_backward(xyz1, xyz2, idx1, idx2)
        ctx.mark_non_differentiable(idx1, idx2)
        return dist1, idx1, dist2, idx2

    @staticmethod
    def backward(ctx, grad_dist1, grad_idx1, grad_dist2, grad_idx2):
        xyz1, xyz2, idx1, idx2 = ctx.saved_tensors
        if grad_dist1 is None:
            grad_dist1 = torch.zeros(idx1.shape, dtype=torch.float32, device=idx1.device)
        if grad_dist2 is None:
            grad_dist2 = torch.zeros(idx2.shape, dtype=torch.float32, device=idx2.device)
        return nn_distance_grad(xyz1, xyz2, grad_dist1, idx1, grad_dist2, idx2)


def nn_distance(xyz1, xyz2):
    '''
Computes the distance of nearest neighbors for a pair of point clouds
input: xyz1: (batch_size,#points_1,3)  the first point cloud
input: xyz2: (batch_size,#points_2,3)  the second point cloud
output: dist1: (batch_size,#point_1)   distance from first to second
output: idx1:  (batch_size,#point_1)   nearest neighbor from first to second
output: dist2: (batch_size,#point_2)   distance from second to first
output: idx2:  (batch_size,#point_2)   nearest neighbor from second to first
    '''
    return _NnDistance.apply(xyz1, xyz2)


def nn_distance_grad(xyz1, xyz2, grad_dist1, idx1, grad_dist2, idx2):
    '''The NnDistanceGrad op (tf_nndistance.cpp:10-18): -> grad_xyz1, grad_xyz2'''
    return _ops.nn_distance_bwd(xyz1, xyz2, grad_dist1, idx1, grad_dist2, idx2)
