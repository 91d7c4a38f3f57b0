"""Per-element distance of the product's and of the reference kernels' EMD results to the fp64 evaluation of the
whole pipeline (oracle_emd_fp64): which of the two fp32 implementations is closer to the truth, element by element.
    python tools/emd_truth_table.py [chair|randn] [B N]   ->  table on stdout (profiles/r2_emd_truth_table.txt)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import oracle
from pointnet_autoencoder_b200 import ops, synthetic

gen = sys.argv[1] if len(sys.argv) > 1 else "chair"
b = int(sys.argv[2]) if len(sys.argv) > 2 else 32
n = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
if gen == "randn":
    xyz1, xyz2 = synthetic.s_randn(b, n, n, seed=7)
else:
    xyz1, xyz2 = synthetic.s_chair(b, n)
x1 = torch.from_numpy(xyz1).cuda(); x2 = torch.from_numpy(xyz2).cuda()
rm = oracle.ref_gpu.approx_match(x1, x2)
rc = oracle.ref_gpu.match_cost(x1, x2, rm).cpu().numpy()
rg1, rg2 = [t.cpu().numpy() for t in oracle.ref_gpu.match_cost_grad(x1, x2, rm)]
fac = ops.approx_match_factors(x1, x2)
pc, pg1, pg2 = [t.cpu().numpy() for t in ops.match_cost_factors(x1, x2, fac)]
sc = lambda a, r: float(np.abs(a - r).max() / max(np.abs(r).max(), 1e-30))
print("# %s B=%d N=M=%d: distance to the fp64 truth (cost: relative; gradients: max-norm relative to the gradient scale)" % (gen, b, n))
print("# elem  cost_prod  cost_ref   grad_prod  grad_ref   prod_vs_ref(grad)")
wins = 0
rows = []
for e in range(b):
    c, t1, t2 = oracle.cpu.emd_fp64(xyz1[e:e + 1], xyz2[e:e + 1])
    ep = max(sc(pg1[e], t1[0]), sc(pg2[e], t2[0])); er = max(sc(rg1[e], t1[0]), sc(rg2[e], t2[0]))
    wins += ep <= er
    rows.append((ep, er))
    print("%5d  %.2e  %.2e   %.2e  %.2e   %.2e" % (e, abs(pc[e] - c[0]) / c[0], abs(rc[e] - c[0]) / c[0], ep, er,
                                                    max(sc(pg1[e], rg1[e]), sc(pg2[e], rg2[e]))))
r = np.array(rows)
print("# product closer to the truth than the reference kernels on %d of %d elements; median %.2e vs %.2e; max %.2e vs %.2e"
      % (wins, b, np.median(r[:, 0]), np.median(r[:, 1]), r[:, 0].max(), r[:, 1].max()))
