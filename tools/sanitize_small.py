"""Small-shape pass over every kernel for compute-sanitizer (one tool per gpurun call):
    compute-sanitizer --tool memcheck  python tools/sanitize_small.py
    compute-sanitizer --tool racecheck python tools/sanitize_small.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pointnet_autoencoder_b200 import ops, synthetic

for (b, n, m) in [(2, 300, 77), (1, 513, 1030), (3, 64, 64), (2, 1, 5)]:
    a, c = synthetic.s_randn(b, n, m, seed=1)
    x1 = torch.from_numpy(a).cuda(); x2 = torch.from_numpy(c).cuda()
    d1, i1, d2, i2 = ops.nn_distance_fwd(x1, x2)
    ops.nn_distance_bwd(x1, x2, torch.ones_like(d1), i1, torch.ones_like(d2), i2)
    fac, dense = ops.approx_match_factors(x1, x2, dense=True)
    ops.match_cost_factors(x1, x2, fac)
    ops.match_cost_dense_fwd(x1, x2, dense); ops.match_cost_dense_bwd(x1, x2, dense)
xe = torch.randn(2, 300, 128, device="cuda").to(torch.bfloat16)
we = torch.randn(256, 128, device="cuda").to(torch.bfloat16)
ops.encoder_conv_pool(xe, we, sign=torch.randn(256, device="cuda"))
torch.cuda.synchronize()
print("sanitize_small: done")
