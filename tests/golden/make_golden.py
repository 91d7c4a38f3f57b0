"""Generate the golden fixtures from the REFERENCE's own CUDA kernels.

Runs on a GPU box (the reference kernels cannot run in the build container):

    gpurun -- 'python tests/golden/make_golden.py gpurun_out/golden'

then copy gpurun_out/golden/*.npz into tests/golden/ and commit them.  The
fixtures hold the inputs and what oracle/_ref/libref_gpu.so (tf_nndistance_g.cu
and tf_approxmatch_g.cu compiled unmodified for sm_100a) returned for them.
The reference ships no golden vectors of its own; these
are "outputs of the reference itself run here".
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import numpy as np
import torch

import oracle
from pointnet_autoencoder_b200 import synthetic

NN_CASES = {
    "nn_randn_2x300x77": ("randn", 2, 300, 77, 11),
    "nn_chair_2x512x512": ("chair", 2, 512, 512, 0),
    "nn_randn_1x1030x513": ("randn", 1, 1030, 513, 5),
}
EMD_CASES = {
    "emd_chair_2x128x128": ("chair", 2, 128, 128, 0),
    "emd_chair_2x200x50": ("chair", 2, 200, 50, 0),
    "emd_randn_1x96x96": ("randn", 1, 96, 96, 7),
    "emd_chair_1x64x256": ("chair", 1, 64, 256, 0),
}


def clouds(gen, b, n, m, seed):
    if gen == "randn":
        return synthetic.s_randn(b, n, m, seed=seed)
    label, pred = synthetic.s_chair(b, max(n, m), first_id=seed)
    return np.ascontiguousarray(label[:, :n]), np.ascontiguousarray(pred[:, :m])


def main(out_dir):
    os.makedirs(out_dir, exist_ok=True)
    R = oracle.ref_gpu
    cu = lambda a: torch.from_numpy(a).cuda()
    for name, (gen, b, n, m, seed) in NN_CASES.items():
        xyz1, xyz2 = clouds(gen, b, n, m, seed)
        d1, i1, d2, i2 = R.nn_distance(cu(xyz1), cu(xyz2))
        rs = np.random.RandomState(1)
        g1 = rs.randn(b, n).astype(np.float32); g2 = rs.randn(b, m).astype(np.float32)
        o1, o2 = R.nn_distance_grad(cu(xyz1), cu(xyz2), cu(g1), i1, cu(g2), i2)
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), xyz1=xyz1, xyz2=xyz2,
                            dist1=d1.cpu().numpy(), idx1=i1.cpu().numpy(), dist2=d2.cpu().numpy(), idx2=i2.cpu().numpy(),
                            grad_dist1=g1, grad_dist2=g2, grad_xyz1=o1.cpu().numpy(), grad_xyz2=o2.cpu().numpy())
        print("wrote", name)
    for name, (gen, b, n, m, seed) in EMD_CASES.items():
        xyz1, xyz2 = clouds(gen, b, n, m, seed)
        match = R.approx_match(cu(xyz1), cu(xyz2))
        cost = R.match_cost(cu(xyz1), cu(xyz2), match)
        g1, g2 = R.match_cost_grad(cu(xyz1), cu(xyz2), match)
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), xyz1=xyz1, xyz2=xyz2,
                            match=match.cpu().numpy(), cost=cost.cpu().numpy(),
                            grad1=g1.cpu().numpy(), grad2=g2.cpu().numpy())
        print("wrote", name)
    np.savez_compressed(os.path.join(out_dir, "levels.npz"), levels=R.levels())


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden")
