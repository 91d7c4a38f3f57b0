"""Kernel-only timing: capture `reps` back-to-back calls of one op in a CUDA graph and
replay it, so host launch overhead is out of the picture.
    python tools/graph_time.py [B N M]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from pointnet_autoencoder_b200 import ops, synthetic


def graph_time(fn, reps=20):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    for _ in range(20):
        g.replay()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    n = 20
    e0.record()
    for _ in range(n):
        g.replay()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / (n * reps)


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    b, n, m = (int(args[0]), int(args[1]), int(args[2])) if len(args) >= 3 else (32, 2048, 2048)
    gen = [a.split("=")[1] for a in sys.argv if a.startswith("--gen=")]
    gen = gen[0] if gen else "randn"
    if gen == "randn":
        x1n, x2n = synthetic.s_randn(b, n, m)
    elif gen == "chair":
        x2n, x1n = synthetic.s_chair(b, max(n, m)); x1n = np.ascontiguousarray(x1n[:, :n]); x2n = np.ascontiguousarray(x2n[:, :m])
    else:       # "dups": label resampled with replacement, pred = label + noise
        rs = np.random.RandomState(0)
        src = rs.uniform(-1, 1, (b, m, 3)).astype(np.float32)
        x2n = np.take_along_axis(src, rs.randint(0, m // 2, (b, m))[:, :, None].repeat(3, 2), 1)
        x1n = np.take_along_axis(x2n, rs.randint(0, m, (b, n))[:, :, None].repeat(3, 2), 1) + (rs.randn(b, n, 3) * 0.02).astype(np.float32)
        x1n = np.ascontiguousarray(x1n); x2n = np.ascontiguousarray(x2n)
    print("inputs: %s B=%d N=%d M=%d" % (gen, b, n, m))
    x1 = torch.from_numpy(x1n).cuda(); x2 = torch.from_numpy(x2n).cuda()
    pairs = b * n * m
    peak = 148 * 128 * 2 * 1.965e9
    d1, i1, d2, i2 = ops.nn_distance_fwd(x1, x2)
    g1 = torch.full((b, n), 100.0 / (b * n), device="cuda"); g2 = torch.full((b, m), 100.0 / (b * m), device="cuda")
    t = graph_time(lambda: ops.nn_distance_fwd(x1, x2))
    print("nn_distance fwd  %8.2f us  %5.1f%% of fp32 peak (16 flop/pair)" % (t * 1e3, 100 * 16 * pairs / (t * 1e-3) / peak))
    t2 = graph_time(lambda: ops.nn_distance_bwd(x1, x2, g1, i1, g2, i2))
    print("nn_distance bwd  %8.2f us" % (t2 * 1e3))
    print("fwd+bwd          %8.2f us  %5.1f%% of fp32 peak  %.0f Gpairs/s" % ((t + t2) * 1e3, 100 * 16 * pairs / ((t + t2) * 1e-3) / peak, pairs / ((t + t2) * 1e-3) / 1e9))
    t3 = graph_time(lambda: ops.nn_distance_fwd_grad(x1, x2, g1, g2))
    print("fwd_grad (fused) %8.2f us  %5.1f%% of fp32 peak  %.0f Gpairs/s" % (t3 * 1e3, 100 * 16 * pairs / (t3 * 1e-3) / peak, pairs / (t3 * 1e-3) / 1e9))
    if "--emd" in sys.argv:
        fac = ops.approx_match_factors(x1, x2)
        t = graph_time(lambda: ops.approx_match_factors(x1, x2), reps=2)
        print("approx_match     %8.2f us  %5.1f%% of fp32 peak (380 flop/pair)" % (t * 1e3, 100 * 380 * pairs / (t * 1e-3) / peak))
        t2 = graph_time(lambda: ops.match_cost_factors(x1, x2, fac), reps=2)
        print("match_cost f+g   %8.2f us" % (t2 * 1e3))
        print("EMD fwd+grad     %8.2f us  %5.1f%% of fp32 peak (423 flop/pair)" % ((t + t2) * 1e3, 100 * 423 * pairs / ((t + t2) * 1e-3) / peak))


def enc():
    b, n, k, c = 32, 2048, 128, 1024
    x = torch.randn(b, n, k, device="cuda").to(torch.bfloat16)
    w = (torch.randn(k, c, device="cuda") / k ** 0.5)
    wt = w.t().contiguous().to(torch.bfloat16)
    t = graph_time(lambda: ops.encoder_conv_pool(x, wt), reps=10)
    flop = 2.0 * b * n * k * c
    to = graph_time(lambda: ops.encoder_conv_pool(x, wt, overlap=True), reps=10)
    print("  ... as a programmatic dependent launch (PNAE_OVERLAP_PREVIOUS, the way the encoder chain enqueues it): %.2f us  %.1f TFLOP/s" % (to * 1e3, flop / (to * 1e-3) / 1e12))
    print("encoder conv5+pool (tcgen05)   %8.2f us  %7.1f TFLOP/s  (%.1f%% of the 1660.6 TF/s measured bf16 burst peak)" % (t * 1e3, flop / (t * 1e-3) / 1e12, 100 * flop / (t * 1e-3) / 1660.6e12))
    xf = x.float()

    def lib():
        y = torch.relu(xf @ w)
        return y.amax(1)
    t2 = graph_time(lib, reps=5)
    print("torch fp32 matmul+relu+amax    %8.2f us" % (t2 * 1e3))
    xb = x; wb = w.to(torch.bfloat16)

    def lib16():
        y = xb @ wb
        return y.amax(1), y.amin(1), y.float().sum(1)
    t3 = graph_time(lib16, reps=5)
    print("torch bf16 matmul+amax/amin/sum %7.2f us" % (t3 * 1e3))


if __name__ == "__main__":
    if "--enc" in sys.argv:
        enc()
        sys.exit(0)
    main()
