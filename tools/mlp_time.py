"""Kernel-only timing of the encoder's layer kernels and of the whole forward (CUDA-graph replay, B=32, N=2048)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pointnet_autoencoder_b200 import ops
from pointnet_autoencoder_b200.encoder import PointNetEncoder
from tools.graph_time import graph_time

b, n = 32, 2048
g = torch.Generator(device="cuda").manual_seed(0)
rnd = lambda *sh: torch.randn(*sh, device="cuda", generator=g)
xyz = rnd(b, n, 3)
w1, b1 = rnd(3, 64), rnd(64)
y1, st1 = ops.mlp_first(xyz, w1, b1)
print("mlp_first 3->64         %7.2f us" % (1e3 * graph_time(lambda: ops.mlp_first(xyz, w1, b1))))
gam, bet, mm, mv = torch.ones(64, device="cuda"), torch.zeros(64, device="cuda"), torch.zeros(64, device="cuda"), torch.ones(64, device="cuda")
for kout in (64, 128):
    w, bias = rnd(64, kout) / 8, rnd(kout)
    t = graph_time(lambda: ops.mlp_layer(y1, st1, gam, bet, mm, mv, True, 0.9, 1e-3, w, bias))
    mb = (y1.numel() + y1.shape[0] * kout) * 4 / 1e6
    print("mlp_layer 64->%-3d       %7.2f us   %.1f MB moved = %.2f TB/s (memset + kernel)" % (kout, 1e3 * t, mb, mb / 1e6 / (t * 1e-3)))
y4, st4 = ops.mlp_layer(y1, st1, gam, bet, mm, mv, True, 0.9, 1e-3, rnd(64, 128) / 8, rnd(128))
g4, b4, mm4, mv4 = torch.ones(128, device="cuda"), torch.zeros(128, device="cuda"), torch.zeros(128, device="cuda"), torch.ones(128, device="cuda")
print("mlp_apply_bf16 128      %7.2f us" % (1e3 * graph_time(lambda: ops.mlp_apply_bf16(y4, st4, g4, b4, mm4, mv4, True, 0.9, 1e-3))))
enc = PointNetEncoder(fused=True).cuda().train()
with torch.no_grad():
    print("encoder forward         %7.2f us" % (1e3 * graph_time(lambda: enc(xyz), reps=5)))
