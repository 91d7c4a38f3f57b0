"""Batched input pipeline (pointnet_autoencoder_b200/input_pipeline.py) against per-cloud numpy restatements of
the reference's host loops (part_dataset.py:12-39,118-121; train.py:196-201).  CPU only: the ops are
device-agnostic library calls."""
import numpy as np
import torch

from pointnet_autoencoder_b200 import input_pipeline as ip
from pointnet_autoencoder_b200 import synthetic


def _ragged(seed=3, count=5):
    rs = np.random.RandomState(seed)
    return [(rs.randn(int(rs.randint(40, 300)), 3) * rs.uniform(0.3, 3.0) + rs.randn(3)).astype(np.float32) for _ in range(count)]


def test_pc_normalize_matches_per_cloud_reference():
    clouds = _ragged()
    ds = ip.DeviceDataset(clouds, npoints=64)
    for i, c in enumerate(clouds):
        ref = synthetic.pc_normalize(c.astype(np.float64))          # part_dataset.py:12-19
        got = ds.points[i, : len(c)].numpy()
        assert np.abs(got - ref).max() <= 2e-6
        assert not ds.points[i, len(c):].any()                     # padding stays zero
        assert abs(np.sqrt((got ** 2).sum(1)).max() - 1.0) <= 1e-6  # unit ball


def test_rotate_y_matches_np_dot():
    rs = np.random.RandomState(0)
    batch = rs.randn(4, 50, 3).astype(np.float32)
    ang = rs.uniform(0, 2 * np.pi, 4)
    got = ip.rotate_y(torch.from_numpy(batch), torch.from_numpy(ang)).numpy()
    for k in range(4):
        c, s = np.cos(ang[k]), np.sin(ang[k])
        rot = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])          # part_dataset.py:33-36
        assert np.abs(got[k] - batch[k].astype(np.float64) @ rot).max() <= 1e-5
    assert np.allclose(got[..., 1], batch[..., 1])                  # y is the rotation axis
    g = torch.Generator().manual_seed(5)
    r = ip.rotate_y(torch.from_numpy(batch), generator=g).numpy()
    assert np.allclose(np.sqrt((r ** 2).sum(-1)), np.sqrt((batch ** 2).sum(-1)), atol=1e-5)   # an isometry per shape


def test_resample_draws_valid_points_with_replacement():
    clouds = _ragged(seed=9, count=6)
    ds = ip.DeviceDataset(clouds, npoints=500, normalize=False)
    g = torch.Generator().manual_seed(11)
    pts, idx = ip.resample(ds.points, ds.lengths, 500, g)
    assert pts.shape == (6, 500, 3)
    for i, c in enumerate(clouds):
        assert int(idx[i].min()) >= 0 and int(idx[i].max()) < len(c)
        assert np.array_equal(pts[i].numpy(), c[idx[i].numpy()])
        assert len(np.unique(idx[i].numpy())) < 500                 # 500 draws from < 300 points repeat
        counts = np.bincount(idx[i].numpy(), minlength=len(c))     # roughly uniform: nobody is starved or favoured 10x
        assert counts.max() <= 12 * 500 / len(c) + 12
    g2 = torch.Generator().manual_seed(11)
    assert torch.equal(ip.resample(ds.points, ds.lengths, 500, g2)[1], idx)                   # reproducible


def test_device_dataset_batch_shape_and_range():
    ds = ip.DeviceDataset(_ragged(count=8), npoints=128)
    g = torch.Generator().manual_seed(2)
    b = ds.batch([1, 5, 2], generator=g)
    assert b.shape == (3, 128, 3) and b.dtype == torch.float32
    assert float(b.pow(2).sum(-1).sqrt().max()) <= 1.0 + 1e-5       # still inside the unit ball after rotation
    assert torch.equal(ds.batch([4], rotate=False, generator=torch.Generator().manual_seed(7)),
                       ds.batch([4], rotate=False, generator=torch.Generator().manual_seed(7)))
