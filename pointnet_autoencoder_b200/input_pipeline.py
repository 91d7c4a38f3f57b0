"""The reference's host-side input pipeline as batched tensor ops.

The reference prepares every batch in Python loops on the host (train.py:170-178, 196-201):
`pc_normalize` once per cloud at load time (part_dataset.py:12-19), a fresh resample-with-replacement
to `npoints` on every access (part_dataset.py:118-121), and a random rotation about the up (y) axis per
shape (part_dataset.py:21-39).  With the training step at a few milliseconds that loop would dominate, so
the same three steps are expressed here over whole batches; the tensors may live on the GPU, in which
case a batch never touches the host.  Plain library ops (gather, bmm): this is plumbing around the hot
path, not part of it.  Randomness comes from a caller-supplied `torch.Generator`, so a run is
reproducible; the random stream itself necessarily differs from numpy's global RNG in the reference.
"""
from __future__ import annotations

import math

import torch


def pc_normalize(clouds, lengths=None):
    """part_dataset.py:12-19 for a padded batch: clouds (S, P, 3), lengths (S,) valid points per cloud
    (None: all P).  Each cloud is centred on the mean of its valid points and divided by its largest
    point radius; padding rows come back as zeros."""
    s, p, _ = clouds.shape
    if lengths is None:
        lengths = torch.full((s,), p, dtype=torch.long, device=clouds.device)
    valid = (torch.arange(p, device=clouds.device)[None, :] < lengths[:, None]).unsqueeze(-1)      # (S, P, 1)
    x = torch.where(valid, clouds, torch.zeros_like(clouds))
    centroid = x.sum(dim=1, keepdim=True) / lengths.clamp(min=1).to(clouds.dtype)[:, None, None]
    x = torch.where(valid, x - centroid, torch.zeros_like(x))
    radius = x.pow(2).sum(dim=2).sqrt().amax(dim=1)                                                   # (S,)
    return x / radius[:, None, None]


def resample(clouds, lengths, npoints, generator=None):
    """part_dataset.py:118-121: for every cloud draw `npoints` indices uniformly WITH replacement from its
    valid points and gather them -> (S, npoints, 3).  Also returns the indices (S, npoints)."""
    s = clouds.shape[0]
    u = torch.rand((s, npoints), generator=generator, device=clouds.device)
    idx = (u * lengths[:, None].to(u.dtype)).long()
    idx = torch.minimum(idx, (lengths[:, None] - 1).clamp(min=0))           # u*len can round up to len in fp32
    return torch.gather(clouds, 1, idx.unsqueeze(-1).expand(s, npoints, 3)), idx


def rotation_matrices_y(angles):
    """(B,) angles -> (B, 3, 3) matrices [[c,0,s],[0,1,0],[-s,0,c]] of part_dataset.py:33-36"""
    c, s = torch.cos(angles), torch.sin(angles)
    z, o = torch.zeros_like(c), torch.ones_like(c)
    return torch.stack([torch.stack([c, z, s], -1), torch.stack([z, o, z], -1), torch.stack([-s, z, c], -1)], -2)


def rotate_y(batch, angles=None, generator=None):
    """part_dataset.py:21-39: every shape of batch (B, N, 3) times its own rotation about y (row vectors times the
    matrix, as np.dot(shape_pc, rotation_matrix) does).  angles (B,) in radians; None draws U(0, 2 pi) per shape."""
    if angles is None:
        angles = torch.rand((batch.shape[0],), generator=generator, device=batch.device) * (2 * math.pi)
    return torch.bmm(batch, rotation_matrices_y(angles.to(batch.dtype)))


class DeviceDataset:
    """A set of variable-length clouds kept normalised and padded on one device; `batch(ids)` does what
    train.py:196-201 does for those ids (resample every cloud to `npoints`, rotate unless told not to)."""

    def __init__(self, clouds, npoints=2048, device="cpu", normalize=True):
        lengths = torch.tensor([int(c.shape[0]) for c in clouds], dtype=torch.long)
        padded = torch.zeros((len(clouds), int(lengths.max()), 3), dtype=torch.float32)
        for i, c in enumerate(clouds):
            padded[i, : c.shape[0]] = torch.as_tensor(c, dtype=torch.float32)
        self.lengths = lengths.to(device)
        self.points = padded.to(device)
        if normalize:
            self.points = pc_normalize(self.points, self.lengths)
        self.npoints = npoints

    def __len__(self):
        return self.points.shape[0]

    def batch(self, ids, generator=None, rotate=True):
        ids = torch.as_tensor(ids, dtype=torch.long, device=self.points.device)
        pts, _ = resample(self.points[ids], self.lengths[ids], self.npoints, generator)
        return rotate_y(pts, generator=generator) if rotate else pts
