"""Sequential vs software-pipelined multi-step Chamfer graphs (B=32, N=M=2048, fused fwd+grad and forward only)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from pointnet_autoencoder_b200 import synthetic
from pointnet_autoencoder_b200.graphs import ChamferStep

b, n, m, steps, ng = 32, 2048, 2048, 8, 16
x1 = [torch.randn(b, n, 3, device="cuda") for _ in range(steps * ng)]
x2 = [torch.randn(b, m, 3, device="cuda") for _ in range(steps * ng)]
for kw in (dict(fused=True), dict(fused=True, pipelined=True), dict(forward_only=True), dict(forward_only=True, pipelined=True), dict(), dict(pipelined=True)):
    first = ChamferStep(x1[:steps], x2[:steps], **kw)
    gs = [first] + [ChamferStep(x1[g * steps:(g + 1) * steps], x2[g * steps:(g + 1) * steps], share_buffers_with=first, **kw) for g in range(1, ng)]
    for g in gs:
        g.run()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        for g in gs:
            g.run()
    e1.record(); e1.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (10 * ng * steps)
    print("%-45s %7.2f us/step  %6.0f Gpairs/s" % (kw, us, b * n * m / us / 1e3))
