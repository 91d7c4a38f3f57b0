"""Developer timing: each op of the product next to the reference's own CUDA kernel
(oracle/_ref/libref_gpu.so) on the same GPU, CUDA events, median of `iters`.
Not the bench contract (bench.py is); a quick look while optimising.

    python tools/quick_time.py [B N M] [--no-ref]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

from pointnet_autoencoder_b200 import ops, synthetic


def timeit(fn, iters=20, warm=3):
    """Back-to-back launches between two events (no host sync inside), after ~60 ms of
    continuous warm-up so the SM clock has ramped; returns (ms per call, same)."""
    import time
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); e1.synchronize()
    one = max(e0.elapsed_time(e1), 1e-3)
    nwarm = int(min(max(60.0 / one, 3), 2000))
    niter = int(min(max(100.0 / one, 5), 2000))
    for _ in range(nwarm):
        fn()
    e0.record()
    for _ in range(niter):
        fn()
    e1.record(); e1.synchronize()
    ms = e0.elapsed_time(e1) / niter
    return ms, ms


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    b, n, m = (int(args[0]), int(args[1]), int(args[2])) if len(args) >= 3 else (32, 2048, 2048)
    use_ref = "--no-ref" not in sys.argv
    label, pred = synthetic.s_chair(b, max(n, m))
    x1 = torch.from_numpy(np.ascontiguousarray(label[:, :n])).cuda()
    x2 = torch.from_numpy(np.ascontiguousarray(pred[:, :m])).cuda()
    pairs = b * n * m
    peak = 148 * 128 * 2 * 1.965e9
    print("B=%d N=%d M=%d pairs=%.1fM  (fp32 peak used for %%: %.1f TFLOP/s)" % (b, n, m, pairs / 1e6, peak / 1e12))

    d1, i1, d2, i2 = ops.nn_distance_fwd(x1, x2)
    g1 = torch.full((b, n), 100.0 / (b * n), device="cuda"); g2 = torch.full((b, m), 100.0 / (b * m), device="cuda")
    rows = []
    rows.append(("nn_distance fwd", timeit(lambda: ops.nn_distance_fwd(x1, x2)), 16 * pairs))
    rows.append(("nn_distance bwd", timeit(lambda: ops.nn_distance_bwd(x1, x2, g1, i1, g2, i2)), 0))
    fac = ops.approx_match_factors(x1, x2)
    rows.append(("approx_match (factors)", timeit(lambda: ops.approx_match_factors(x1, x2), iters=5, warm=1), 380 * pairs))
    rows.append(("match_cost_factors cost+grad", timeit(lambda: ops.match_cost_factors(x1, x2, fac), iters=5, warm=1), 43 * pairs))
    rows.append(("match_cost_factors cost only", timeit(lambda: ops.match_cost_factors(x1, x2, fac, with_grad=False), iters=5, warm=1), 11 * pairs))
    if b * n * m * 4 <= 8 << 30:
        rows.append(("approx_match (dense RMW)", timeit(lambda: ops.approx_match_factors(x1, x2, dense=True), iters=3, warm=1), 380 * pairs))
        dense = ops.match_from_factors(x1, x2, fac)
        rows.append(("match_from_factors", timeit(lambda: ops.match_from_factors(x1, x2, fac), iters=5, warm=1), 0))
        rows.append(("match_cost dense fwd", timeit(lambda: ops.match_cost_dense_fwd(x1, x2, dense), iters=5, warm=1), 11 * pairs))
        rows.append(("match_cost dense bwd", timeit(lambda: ops.match_cost_dense_bwd(x1, x2, dense), iters=5, warm=1), 32 * pairs))
    for name, (med, best), flop in rows:
        print("  mine  %-30s median %9.3f ms  best %9.3f ms  %6.1f%% of fp32 peak" % (name, med, best, 100 * flop / (best * 1e-3) / peak))

    if use_ref:
        import oracle
        R = oracle.ref_gpu
        if not R.available():
            print("  (reference GPU library not present)")
            return
        lib = R.lib
        p = R._p
        rd1 = torch.empty_like(d1); ri1 = torch.empty_like(i1); rd2 = torch.empty_like(d2); ri2 = torch.empty_like(i2)
        o1 = torch.empty((b, n, 3), device="cuda"); o2 = torch.empty((b, m, 3), device="cuda")
        rows = []
        rows.append(("nn_distance fwd", timeit(lambda: lib.ref_gpu_nn_distance(b, n, p(x1), m, p(x2), p(rd1), p(ri1), p(rd2), p(ri2))), 16 * pairs))
        rows.append(("nn_distance bwd", timeit(lambda: lib.ref_gpu_nn_distance_grad(b, n, p(x1), m, p(x2), p(g1), p(i1), p(g2), p(i2), p(o1), p(o2))), 0))
        if b * n * m < 2 ** 31:
            match = torch.empty((b, m, n), device="cuda"); temp = torch.empty((max(b, 32), 2 * (n + m)), device="cuda")
            cost = torch.empty((b,), device="cuda")
            rows.append(("approxmatch", timeit(lambda: lib.ref_gpu_approxmatch(b, n, m, p(x1), p(x2), p(match), p(temp)), iters=3, warm=1), 380 * pairs))
            rows.append(("matchcost", timeit(lambda: lib.ref_gpu_matchcost(b, n, m, p(x1), p(x2), p(match), p(cost)), iters=3, warm=1), 11 * pairs))
            rows.append(("matchcostgrad", timeit(lambda: lib.ref_gpu_matchcostgrad(b, n, m, p(x1), p(x2), p(match), p(o1), p(o2)), iters=3, warm=1), 32 * pairs))
        for name, (med, best), flop in rows:
            print("  ref   %-30s median %9.3f ms  best %9.3f ms  %6.1f%% of fp32 peak" % (name, med, best, 100 * flop / (best * 1e-3) / peak))


if __name__ == "__main__":
    main()
