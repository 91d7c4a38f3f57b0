"""How far apart are three evaluations of the SAME approx-EMD algorithm?  (product kernels, the reference's own
CUDA kernels, the CPU oracle).  Prints max|diff| / max|ref| for cost and gradients on a few shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oracle
from pointnet_autoencoder_b200 import ops, synthetic

def sc(a, r):
    return float(np.abs(a - r).max() / max(np.abs(r).max(), 1e-30))

print("%-18s %-7s | cost: mine-orc  ref-orc  mine-ref | grad1: mine-orc  ref-orc  mine-ref | grad2: mine-orc ref-orc mine-ref" % ("shape", "data"))
for gen, b, n, m in [("chair", 5, 76, 163), ("chair", 2, 600, 600), ("chair", 2, 128, 128), ("chair", 2, 1024, 1024), ("randn", 2, 512, 512), ("chair", 1, 400, 100)]:
    if gen == "randn":
        xyz1, xyz2 = synthetic.s_randn(b, n, m, seed=7)
    else:
        label, pred = synthetic.s_chair(b, max(n, m), first_id=n + m)
        xyz1 = np.ascontiguousarray(label[:, :n]); xyz2 = np.ascontiguousarray(pred[:, :m])
    x1 = torch.from_numpy(xyz1).cuda(); x2 = torch.from_numpy(xyz2).cuda()
    fac = ops.approx_match_factors(x1, x2)
    c_m, g1_m, g2_m = [t.cpu().numpy() for t in ops.match_cost_factors(x1, x2, fac)]
    om = oracle.cpu.approx_match(xyz1, xyz2)
    c_o = oracle.cpu.match_cost(xyz1, xyz2, om); g1_o, g2_o = oracle.cpu.match_cost_grad(xyz1, xyz2, om)
    rm = oracle.ref_gpu.approx_match(x1, x2)
    c_r = oracle.ref_gpu.match_cost(x1, x2, rm).cpu().numpy()
    g1_r, g2_r = [t.cpu().numpy() for t in oracle.ref_gpu.match_cost_grad(x1, x2, rm)]
    print("%-18s %-7s | %.1e %.1e %.1e | %.1e %.1e %.1e | %.1e %.1e %.1e" % (
        "%dx%dx%d" % (b, n, m), gen, sc(c_m, c_o), sc(c_r, c_o), sc(c_m, c_r),
        sc(g1_m, g1_o), sc(g1_r, g1_o), sc(g1_m, g1_r), sc(g2_m, g2_o), sc(g2_r, g2_o), sc(g2_m, g2_r)))
