"""Run each hot op a few times at the headline size (for ncu: -k regex:<kernel> -s <skip> -c <count>)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pointnet_autoencoder_b200 import ops, synthetic

which = sys.argv[1] if len(sys.argv) > 1 else "chamfer"
b, n, m = 32, 2048, 2048
x1n, x2n = synthetic.s_randn(b, n, m)
x1 = torch.from_numpy(x1n).cuda(); x2 = torch.from_numpy(x2n).cuda()
g1 = torch.full((b, n), 100.0 / (b * n), device="cuda"); g2 = torch.full((b, m), 100.0 / (b * m), device="cuda")
xe = torch.randn(b, n, 128, device="cuda").to(torch.bfloat16)
we = (torch.randn(1024, 128, device="cuda") / 128 ** 0.5).to(torch.bfloat16)
for _ in range(3):
    if which in ("chamfer", "all"):
        ops.nn_distance_fwd_grad(x1, x2, g1, g2)               # the benchmarked two-kernel step
        d1, i1, d2, i2 = ops.nn_distance_fwd(x1, x2)           # and the two-op form
        ops.nn_distance_bwd(x1, x2, g1, i1, g2, i2)
    if which in ("emd", "all"):
        fac = ops.approx_match_factors(x1, x2)
        ops.match_cost_factors(x1, x2, fac)
    if which in ("enc", "all"):
        ops.encoder_conv_pool(xe, we)
torch.cuda.synchronize()
print("ok")
