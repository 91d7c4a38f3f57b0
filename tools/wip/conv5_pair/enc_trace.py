"""Print the pipeline timeline of two conv5 CTAs (needs a libpnae built with -DPNAE_ENC_TRACE, see tools/build_variant.sh)."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from pointnet_autoencoder_b200 import _lib, ops

b, n, k, c = 32, 2048, 128, 1024
x = torch.randn(b, n, k, device="cuda").to(torch.bfloat16)
wt = (torch.randn(c, k, device="cuda") / k ** 0.5).to(torch.bfloat16)
for _ in range(5):
    ops.encoder_conv_pool(x, wt)
torch.cuda.synchronize()
lib = _lib.load()
buf = np.zeros((2, 16, 8), np.int64)
rc = lib.pnae_debug_enc_trace(buf.ctypes.data_as(ctypes.c_void_p))
assert rc == 0, rc
clk = 1.9   # GHz, nominal: cycles -> us below are approximate
names = ["load issued", "acc free", "data ready", "mma issued", "epi start", "epi done", "tmem read", "reduced"]
for cta in range(2):
    t0 = buf[cta, 15, 0]
    print("CTA %d: setup done +%.2f us, exit +%.2f us" % (cta, (buf[cta, 15, 1] - t0) / clk / 1e3, (buf[cta, 15, 2] - t0) / clk / 1e3))
    print("  tile " + " ".join("%11s" % s for s in names))
    for t in range(15):
        if buf[cta, t].any():
            print("  %4d " % t + " ".join("%11.2f" % ((v - t0) / clk / 1e3) for v in buf[cta, t]))
