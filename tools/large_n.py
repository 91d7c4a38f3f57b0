"""BASELINE configs[4] sizes on one GPU: N=M in {4096, 8192, 16384}, per-GPU batch 8 (= B=64 over 8 GPUs).
Parity against the reference's own CUDA kernels where their 32-bit indexing allows, and timings."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oracle
from pointnet_autoencoder_b200 import ops, synthetic

def t_ms(fn, it=3):
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / it

peak = 148 * 128 * 2 * 1.965e9
for n in (4096, 8192, 16384):
    b = 8
    xyz1, xyz2 = synthetic.s_randn(b, n, n, seed=n)
    x1 = torch.from_numpy(xyz1).cuda(); x2 = torch.from_numpy(xyz2).cuda()
    d1, i1, d2, i2 = ops.nn_distance_fwd(x1, x2)
    rb = 2
    r = oracle.ref_gpu.nn_distance(x1[:rb].contiguous(), x2[:rb].contiguous())
    exact = all(torch.equal(a[:rb], c) for a, c in zip((d1, i1, d2, i2), r))
    tf = t_ms(lambda: ops.nn_distance_fwd(x1, x2))
    g1 = torch.ones_like(d1); g2 = torch.ones_like(d2)
    tb = t_ms(lambda: ops.nn_distance_bwd(x1, x2, g1, i1, g2, i2))
    pairs = b * n * n
    msg = "N=%5d B=%d chamfer fwd %8.3f ms (%4.1f%% fp32 peak) bwd %6.3f ms  bit-exact vs reference kernels: %s" % (n, b, tf, 100 * 16 * pairs / (tf * 1e-3) / peak, tb, exact)
    fac = ops.approx_match_factors(x1, x2)
    ta = t_ms(lambda: ops.approx_match_factors(x1, x2), it=2)
    cost, q1, q2 = ops.match_cost_factors(x1, x2, fac)
    tc = t_ms(lambda: ops.match_cost_factors(x1, x2, fac), it=2)
    msg += " | approx_match %8.2f ms match_cost+grad %7.2f ms (EMD %4.1f%% fp32 peak)" % (ta, tc, 100 * 423 * pairs / ((ta + tc) * 1e-3) / peak)
    if n <= 8192:
        rm = oracle.ref_gpu.approx_match(x1[:1].contiguous(), x2[:1].contiguous())
        rc = oracle.ref_gpu.match_cost(x1[:1].contiguous(), x2[:1].contiguous(), rm)
        msg += " cost rel.err vs ref %.1e" % float(((cost[:1] - rc).abs() / rc.abs()).max())
    print(msg, flush=True)
