""" Approxmiate algorithm for computing the Earch Mover's Distance -- B200-native
drop-in for the reference module tf_ops/approxmatch/tf_approxmatch.py.

Same module name, functions, argument order and output shapes
(tf_approxmatch.py:13-51):

    match = approx_match(xyz1, xyz2)          # no gradient (ops.NoGradient('ApproxMatch'))
    cost  = match_cost(xyz1, xyz2, match)     # gradient [grad_1*gc, grad_2*gc, None]

`approx_match` returns a `Match`: a tensor-like handle on the (batch,#query,#dataset)
soft assignment that stores only the per-level factors ratioL_j / ratioR_j
(10*(n+m) floats per element instead of n*m).  `match_cost` recognises it and
evaluates cost and both gradients in one fused pass without the dense tensor ever
reaching HBM.  Anything else that touches a `Match` (torch functions, `.dense()`,
indexing, `.cpu()`) materialises the dense float32 (B,M,N) tensor on demand, and
`match_cost` equally accepts a plain dense tensor -- so code written against the
reference API keeps working.  `approx_match(..., dense=True)` returns the dense
tensor directly.
"""
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

import torch  # noqa: E402

from pointnet_autoencoder_b200 import ops as _ops  # noqa: E402


class Match(object):
    """Factor form of the soft assignment:
    match[i,l,k] = sum_j exp(level_j*|xyz1[i,k]-xyz2[i,l]|^2) * ratioL_j[i,k] * ratioR_j[i,l]
    (the only write to `match` in the reference: tf_approxmatch_g.cu:145-153)."""

    def __init__(self, xyz1, xyz2, factors):
        # The handle keeps its OWN copy of the clouds (1.5 MB at B=32, N=2048): the factors only mean something together
        # with the coordinates they were computed from, and the caller may overwrite its tensors in place afterwards.
        # The source tensors' identity and version counters are remembered so match_cost can tell whether it is being
        # called with exactly those (unmodified) clouds.
        self._src = tuple((t.data_ptr(), t._version, tuple(t.shape), t.is_contiguous()) for t in (xyz1, xyz2))
        self.xyz1 = xyz1.detach().clone()
        self.xyz2 = xyz2.detach().clone()
        self.factors = factors          # (B, 10, N+M)
        self._dense = None

    def computed_from(self, xyz1, xyz2):
        """True iff (xyz1, xyz2) are the very tensors this match was computed from, unchanged since"""
        return all(t.is_contiguous() and (t.data_ptr(), t._version, tuple(t.shape), True) == src
                   for t, src in zip((xyz1, xyz2), self._src))

    @property
    def shape(self):
        return torch.Size((self.xyz1.shape[0], self.xyz2.shape[1], self.xyz1.shape[1]))

    def size(self, dim=None):
        return self.shape if dim is None else self.shape[dim]

    def dim(self):
        return 3

    @property
    def dtype(self):
        return torch.float32

    @property
    def device(self):
        return self.factors.device

    @property
    def requires_grad(self):
        return False

    def dense(self):
        """The reference's (batch,#query,#dataset) tensor, materialised once."""
        if self._dense is None:
            self._dense = _ops.match_from_factors(self.xyz1, self.xyz2, self.factors)
        return self._dense

    def detach(self):
        return self

    def cpu(self):
        return self.dense().cpu()

    def numpy(self):
        return self.dense().cpu().numpy()

    def __getitem__(self, item):
        return self.dense()[item]

    def __repr__(self):
        return "Match(shape=%s, device=%s, factor form)" % (tuple(self.shape), self.device)

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        def conv(a):
            if isinstance(a, Match):
                return a.dense()
            if isinstance(a, (list, tuple)):
                return type(a)(conv(x) for x in a)
            return a
        return func(*conv(args), **{k: conv(v) for k, v in (kwargs or {}).items()})


def approx_match(xyz1, xyz2, dense=False):
    '''
input:
    xyz1 : batch_size * #dataset_points * 3
    xyz2 : batch_size * #query_points * 3
returns:
    match : batch_size * #query_points * #dataset_points
    '''
    with torch.no_grad():          # ops.NoGradient('ApproxMatch'), tf_approxmatch.py:22
        xyz1 = xyz1.detach()
        xyz2 = xyz2.detach()
        if dense:
            return _ops.approx_match_factors(xyz1, xyz2, dense=True)[1]
        return Match(xyz1, xyz2, _ops.approx_match_factors(xyz1, xyz2))


class _MatchCostFactors(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz1, xyz2, handle):
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            cost, g1, g2 = _ops.match_cost_factors(xyz1, xyz2, handle.factors, with_grad=True)
            ctx.save_for_backward(g1, g2)
        else:
            cost = _ops.match_cost_factors(xyz1, xyz2, handle.factors, with_grad=False)
        return cost

    @staticmethod
    def backward(ctx, grad_cost):
        g1, g2 = ctx.saved_tensors
        gc = grad_cost[:, None, None]
        return g1 * gc, g2 * gc, None     # tf_approxmatch.py:51


class _MatchCostDense(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz1, xyz2, match):
        ctx.save_for_backward(xyz1, xyz2, match)
        return _ops.match_cost_dense_fwd(xyz1, xyz2, match)

    @staticmethod
    def backward(ctx, grad_cost):
        xyz1, xyz2, match = ctx.saved_tensors
        g1, g2 = match_cost_grad(xyz1, xyz2, match)
        gc = grad_cost[:, None, None]
        return g1 * gc, g2 * gc, None     # match is treated as a constant


def match_cost(xyz1, xyz2, match):
    '''
input:
    xyz1 : batch_size * #dataset_points * 3
    xyz2 : batch_size * #query_points * 3
    match : batch_size * #query_points * #dataset_points
returns:
    cost : batch_size
    '''
    if isinstance(match, Match):
        b, m, n = match.shape
        if not (xyz1.dim() == 3 and xyz2.dim() == 3 and xyz1.shape[0] == b and xyz1.shape[1] == n and xyz2.shape[1] == m):
            raise ValueError("MatchCost expects (batch_size,#query,#dataset) match shape")
        # The factors re-evaluate exp(level*d) from the coordinates, so the fused path is the
        # reference's "match is a constant" only for the clouds the match was computed from.
        # (same storage AND same version counter: an in-place update of the clouds between approx_match and match_cost
        # must see the match as the constant it was, i.e. the dense tensor built from the handle's own snapshot)
        if match.computed_from(xyz1, xyz2):
            return _MatchCostFactors.apply(xyz1, xyz2, match)
        match = match.dense()
    return _MatchCostDense.apply(xyz1, xyz2, match.detach())


def match_cost_grad(xyz1, xyz2, match):
    '''The MatchCostGrad op (tf_approxmatch.cpp:16-21): -> grad1, grad2 (not yet scaled by grad_cost)'''
    if isinstance(match, Match):
        if match.computed_from(xyz1, xyz2):
            _, g1, g2 = _ops.match_cost_factors(xyz1, xyz2, match.factors, with_grad=True)
            return g1, g2
        match = match.dense()
    return _ops.match_cost_dense_bwd(xyz1, xyz2, match)
