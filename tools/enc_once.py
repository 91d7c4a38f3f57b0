"""A few encoder forwards (training-mode BN) at B=32, N=2048 for an ncu launch list."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pointnet_autoencoder_b200.encoder import PointNetEncoder

enc = PointNetEncoder(fused=True).cuda().train()
pc = torch.randn(32, 2048, 3, device="cuda")
with torch.no_grad():
    for _ in range(3):
        enc(pc)
torch.cuda.synchronize()
print("ok")
