"""Synthetic point clouds for parity tests and benchmarks (no dataset ships with
the reference: `data/` holds only a .gitignore, and there is no network).

Two generators, both specified in SURVEY.md section 8(d):

  s_randn  -- mirrors the reference's only timing loop
              (tf_ops/nn_distance/tf_nndistance.py:45-49): np.random.seed(100),
              xyz1=randn(B,N,3), xyz2=randn(B,M,3), float32.
  s_chair  -- autoencoder-like: N points area-uniform on a box-composite chair,
              random anisotropic scale, then the reference's pc_normalize
              (part_dataset.py:12-19: subtract centroid, divide by max radius);
              label=cloud, pred=label[perm]+N(0,0.02^2) (a partially trained AE).
"""
from __future__ import annotations

import numpy as np


def s_randn(b, n, m, seed=100):
    rs = np.random.RandomState(seed)
    xyz1 = rs.randn(b, n, 3).astype(np.float32)
    xyz2 = rs.randn(b, m, 3).astype(np.float32)
    return xyz1, xyz2


def pc_normalize(pc):
    """part_dataset.py:12-19"""
    pc = pc - np.mean(pc, axis=0)
    return pc / np.max(np.sqrt(np.sum(pc ** 2, axis=1)))


# (centre, size) of the boxes making up the chair: seat, back, four legs
_BOXES = [
    ((0.0, 0.0, 0.0), (0.5, 0.06, 0.5)),
    ((0.0, 0.33, -0.22), (0.5, 0.6, 0.06)),
    ((-0.22, -0.255, -0.22), (0.06, 0.45, 0.06)),
    ((0.22, -0.255, -0.22), (0.06, 0.45, 0.06)),
    ((-0.22, -0.255, 0.22), (0.06, 0.45, 0.06)),
    ((0.22, -0.255, 0.22), (0.06, 0.45, 0.06)),
]


def _chair_surface(n, rng):
    faces = []  # (area, centre, size, fixed axis, sign)
    for c, s in _BOXES:
        for ax in range(3):
            a, bb = [s[i] for i in range(3) if i != ax]
            for sg in (-1.0, 1.0):
                faces.append((a * bb, c, s, ax, sg))
    areas = np.array([f[0] for f in faces])
    cen = np.array([f[1] for f in faces]); siz = np.array([f[2] for f in faces])
    axs = np.array([f[3] for f in faces]); sgn = np.array([f[4] for f in faces])
    pick = rng.choice(len(faces), size=n, p=areas / areas.sum())
    u = rng.uniform(-0.5, 0.5, size=(n, 3))
    pts = cen[pick] + u * siz[pick]
    rows = np.arange(n)
    ax = axs[pick]
    pts[rows, ax] = cen[pick, ax] + sgn[pick] * 0.5 * siz[pick, ax]
    return pts


def s_chair(b, n, first_id=0, noise=0.02):
    """-> label (B,N,3), pred (B,N,3) float32"""
    label = np.empty((b, n, 3), np.float32)
    pred = np.empty((b, n, 3), np.float32)
    for i in range(b):
        rng = np.random.default_rng(1000 + first_id + i)
        pts = _chair_surface(n, rng) * rng.uniform(0.8, 1.2, size=3)
        pts = pc_normalize(pts)
        label[i] = pts.astype(np.float32)
        pred[i] = (pts[rng.permutation(n)] + rng.normal(0.0, noise, size=(n, 3))).astype(np.float32)
    return label, pred
