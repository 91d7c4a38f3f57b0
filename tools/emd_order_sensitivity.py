"""How far does a change of summation order alone move the EMD gradients on one cloud?  The fp32 restatement is run
with its per-point sums formed sequentially (the reference's order) and as partial sums over blocks of 64..1024
streamed points; each result is compared with the fp64 evaluation of the pipeline.  CPU only.
    python tools/emd_order_sensitivity.py [chair|randn] [element ...]
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import oracle
from pointnet_autoencoder_b200 import synthetic

gen = sys.argv[1] if len(sys.argv) > 1 else "chair"
elems = [int(a) for a in sys.argv[2:]] or [10, 0]
n = 2048
O = oracle.cpu
fp = C.POINTER(C.c_float)
sc = lambda a, r: float(np.abs(a - r).max() / max(np.abs(r).max(), 1e-30))
if gen == "chair":
    x1, x2 = synthetic.s_chair(max(elems) + 1, n)
else:
    x1, x2 = synthetic.s_randn(max(elems) + 1, n, n, seed=7)
print("# %s N=M=%d: max-norm distance of the fp32 gradients to the fp64 truth, by summation order" % (gen, n))
for e in elems:
    a = np.ascontiguousarray(x1[e:e + 1]); c = np.ascontiguousarray(x2[e:e + 1])
    _, t1, t2 = O.emd_fp64(a, c)
    row = []
    for chunk in (0, 64, 128, 256, 512, 1024):
        fac = O.approx_match_order(a, c, chunk)
        _, g1, g2 = O.match_cost_factors(a, c, fac)
        row.append(max(sc(g1, t1), sc(g2, t2)))
    print("element %2d   sequential %.2e | blocks of 64 %.2e  128 %.2e  256 %.2e  512 %.2e  1024 %.2e" % ((e,) + tuple(row)), flush=True)
