"""Golden fixtures = what the REFERENCE's own CUDA kernels returned on a B200
(tests/golden/make_golden.py; oracle/_ref/libref_gpu.so is tf_nndistance_g.cu and
tf_approxmatch_g.cu compiled unmodified for sm_100a).

CPU half (runs everywhere): the C restatement must reproduce them -- this is what
pins the oracle.  GPU half (-m gpu): the product's CUDA path must reproduce them.
Chamfer is bit-exact; EMD is compared at the north-star tolerances (the hardware
exp2/rsqrt approximations cannot be restated bit-for-bit on a CPU).
"""
import glob
import os

import numpy as np
import pytest

import oracle

GOLDEN = os.path.dirname(os.path.abspath(__file__)) + "/golden"
NN = sorted(glob.glob(GOLDEN + "/nn_*.npz"))
EMD = sorted(glob.glob(GOLDEN + "/emd_*.npz"))
O = oracle.cpu


def ids(paths):
    return [os.path.basename(p)[:-4] for p in paths]


def close_scaled(a, ref, rel):
    a = np.asarray(a, np.float64); ref = np.asarray(ref, np.float64)
    assert np.abs(a - ref).max() <= rel * max(np.abs(ref).max(), 1e-30)


def test_fixtures_exist():
    assert len(NN) >= 3 and len(EMD) >= 4
    lv = np.load(GOLDEN + "/levels.npz")["levels"]
    # -powf(4,j) on the device is an exact power of four (tf_approxmatch_g.cu:22)
    assert lv.tolist() == [-16384.0, -4096.0, -1024.0, -256.0, -64.0, -16.0, -4.0, -1.0, -0.25, 0.0]


@pytest.mark.parametrize("path", NN, ids=ids(NN))
def test_oracle_chamfer_matches_reference_gpu(path):
    g = np.load(path)
    d1, i1, d2, i2 = O.nn_distance(g["xyz1"], g["xyz2"], contract=True)
    assert np.array_equal(d1, g["dist1"]) and np.array_equal(d2, g["dist2"])      # bit-exact
    assert np.array_equal(i1, g["idx1"]) and np.array_equal(i2, g["idx2"])
    o1, o2 = O.nn_distance_grad(g["xyz1"], g["xyz2"], g["grad_dist1"], g["idx1"], g["grad_dist2"], g["idx2"])
    close_scaled(o1, g["grad_xyz1"], 1e-5)      # the reference scatters with float atomics: order differs
    close_scaled(o2, g["grad_xyz2"], 1e-5)


@pytest.mark.parametrize("path", EMD, ids=ids(EMD))
def test_oracle_emd_matches_reference_gpu(path):
    g = np.load(path)
    xyz1, xyz2 = g["xyz1"], g["xyz2"]
    n, m = xyz1.shape[1], xyz2.shape[1]
    scale = max(1.0, n / m)
    match = O.approx_match(xyz1, xyz2)
    np.testing.assert_allclose(match, g["match"], rtol=0, atol=2e-4 * scale)
    assert np.abs(match - g["match"]).mean() < 2e-7 * scale
    np.testing.assert_allclose(O.match_cost(xyz1, xyz2, match), g["cost"], rtol=1e-5)
    # the oracle's cost/grad functions on the REFERENCE's match: isolates them from approx_match
    np.testing.assert_allclose(O.match_cost(xyz1, xyz2, g["match"]), g["cost"], rtol=2e-6)
    g1, g2 = O.match_cost_grad(xyz1, xyz2, g["match"])
    close_scaled(g1, g["grad1"], 2e-6)
    close_scaled(g2, g["grad2"], 2e-6)
    g1, g2 = O.match_cost_grad(xyz1, xyz2, match)
    close_scaled(g1, g["grad1"], 1e-4)
    close_scaled(g2, g["grad2"], 1e-4)


@pytest.mark.gpu
@pytest.mark.parametrize("path", NN, ids=ids(NN))
def test_cuda_chamfer_matches_reference_gpu(path):
    import torch
    from pointnet_autoencoder_b200.tf_ops.nn_distance import tf_nndistance
    g = np.load(path)
    x1 = torch.from_numpy(g["xyz1"]).cuda(); x2 = torch.from_numpy(g["xyz2"]).cuda()
    d1, i1, d2, i2 = tf_nndistance.nn_distance(x1, x2)
    assert np.array_equal(d1.cpu().numpy(), g["dist1"]) and np.array_equal(d2.cpu().numpy(), g["dist2"])
    assert np.array_equal(i1.cpu().numpy(), g["idx1"]) and np.array_equal(i2.cpu().numpy(), g["idx2"])
    o1, o2 = tf_nndistance.nn_distance_grad(x1, x2, torch.from_numpy(g["grad_dist1"]).cuda(), i1,
                                            torch.from_numpy(g["grad_dist2"]).cuda(), i2)
    close_scaled(o1.cpu().numpy(), g["grad_xyz1"], 1e-5)
    close_scaled(o2.cpu().numpy(), g["grad_xyz2"], 1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("path", EMD, ids=ids(EMD))
def test_cuda_emd_matches_reference_gpu(path):
    import torch
    from pointnet_autoencoder_b200.tf_ops.approxmatch import tf_approxmatch
    g = np.load(path)
    x1 = torch.from_numpy(g["xyz1"]).cuda().requires_grad_(True)
    x2 = torch.from_numpy(g["xyz2"]).cuda().requires_grad_(True)
    n, m = x1.shape[1], x2.shape[1]
    scale = max(1.0, n / m)
    match = tf_approxmatch.approx_match(x1, x2)
    np.testing.assert_allclose(match.dense().cpu().numpy(), g["match"], rtol=0, atol=2e-5 * scale)
    cost = tf_approxmatch.match_cost(x1, x2, match)
    cost.sum().backward()
    np.testing.assert_allclose(cost.detach().cpu().numpy(), g["cost"], rtol=1e-5)
    close_scaled(x1.grad.cpu().numpy(), g["grad1"], 1e-4)
    close_scaled(x2.grad.cpu().numpy(), g["grad2"], 1e-4)
