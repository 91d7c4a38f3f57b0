"""Phase timeline of two mlp_layer CTAs (needs a libpnae built with -DPNAE_MLP_TRACE, see tools/build_variant.sh)."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from pointnet_autoencoder_b200 import _lib, ops

b, n = 32, 2048
g = torch.Generator(device="cuda").manual_seed(0)
rnd = lambda *sh: torch.randn(*sh, device="cuda", generator=g)
y1, st1 = ops.mlp_first(rnd(b, n, 3), rnd(3, 64), rnd(64))
gam, bet, mm, mv = torch.ones(64, device="cuda"), torch.zeros(64, device="cuda"), torch.zeros(64, device="cuda"), torch.ones(64, device="cuda")
w, bias = rnd(64, 64) / 8, rnd(64)
for _ in range(5):
    ops.mlp_layer(y1, st1, gam, bet, mm, mv, True, 0.9, 1e-3, w, bias)
torch.cuda.synchronize()
buf = np.zeros((2, 32), np.int64)
assert _lib.load().pnae_debug_mlp_trace(buf.ctypes.data_as(ctypes.c_void_p)) == 0
for cta in range(2):
    t0 = buf[cta, 0]
    us = lambda i: (buf[cta, i] - t0) / 1.9e3
    print("CTA %d: W staged %.2f, set-up done %.2f" % (cta, us(1), us(2)))
    i = 3
    while i + 3 < 31 and buf[cta, i]:
        print("   tile: stage free + rows loaded %.2f  operands written %.2f  MMAs issued %.2f  epilogue done %.2f" % (us(i), us(i + 1), us(i + 2), us(i + 3)))
        i += 4
    print("   exit %.2f" % us(31))
