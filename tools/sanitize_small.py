"""Small-shape pass over every kernel for compute-sanitizer (one tool per gpurun call):
    compute-sanitizer --tool memcheck  python tools/sanitize_small.py
    compute-sanitizer --tool racecheck python tools/sanitize_small.py
(round 2: compute-sanitizer is closed on the GPU pool; the script still serves as a plain small-shape pass over every
entry point, including the encoder chain with its programmatic dependent launches, forward and backward)
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
ops.nn_distance_fwd_grad(x1, x2, torch.ones_like(d1), torch.ones_like(d2))
# the encoder chain: moments + layers 1-2 in one kernel, tcgen05 layers, bf16 apply, conv5, finish (programmatic dependent launches)
from pointnet_autoencoder_b200.encoder import PointNetEncoder
enc = PointNetEncoder(fused=True).cuda().train()
pc = torch.randn(3, 333, 3, device="cuda")
out = enc(pc)
out.sum().backward()
with torch.no_grad():
    enc.eval()(pc)
y1, st1 = ops.mlp_first(pc, torch.randn(3, 64, device="cuda"), torch.randn(64, device="cuda"))
one = torch.ones(64, device="cuda")
ops.mlp_layer(y1, st1, one, torch.zeros_like(one), torch.zeros_like(one), one.clone(), True, 0.9, 1e-3, torch.randn(64, 128, device="cuda"), torch.randn(128, device="cuda"))
torch.cuda.synchronize()
print("sanitize_small: done")
