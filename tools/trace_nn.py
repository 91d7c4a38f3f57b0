"""Per-warp timeline of the Chamfer sweep (build with PNAE_NVCC_DEFS=-DPNAE_NN_TRACE): when each warp enters,
gets past the dependency wait, has its first data, and exits -- relative to the first warp's entry.
    PNAE_NVCC_DEFS=-DPNAE_NN_TRACE python -m pointnet_autoencoder_b200.build && PNAE_NVCC_DEFS=-DPNAE_NN_TRACE python tools/trace_nn.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from pointnet_autoencoder_b200 import ops, synthetic

b, n, m = 32, 2048, 2048
x1n, x2n = synthetic.s_randn(b, n, m)
x1 = torch.from_numpy(x1n).cuda(); x2 = torch.from_numpy(x2n).cuda()
nw = torch.cuda.get_device_properties(0).multi_processor_count * 4 * 4
trace = torch.zeros(nw * 4, dtype=torch.int64, device="cuda")
os.environ["PNAE_NN_TRACE_PTR"] = "%x" % trace.data_ptr()
for _ in range(5):
    ops.nn_distance_fwd(x1, x2)          # back to back: the last call's sweep follows a finalize, as in the bench
    trace.zero_()
ops.nn_distance_fwd(x1, x2); ops.nn_distance_fwd(x1, x2)
trace.zero_()
torch.cuda.synchronize()
for _ in range(3):
    ops.nn_distance_fwd(x1, x2)
torch.cuda.synchronize()
t = trace.cpu().numpy().reshape(nw, 4).astype(np.float64)
t0 = t[:, 0].min()
t = (t - t0) / 1e3
names = ["entry", "past dependency wait", "first rows+chunk landed", "exit"]
for i, nm in enumerate(names):
    c = t[:, i]
    print("%-26s min %7.2f  p10 %7.2f  median %7.2f  p90 %7.2f  max %7.2f us" % (nm, c.min(), np.percentile(c, 10), np.median(c), np.percentile(c, 90), c.max()))
print("busy (first data -> exit)  min %7.2f  median %7.2f  max %7.2f us" % ((t[:, 3] - t[:, 2]).min(), np.median(t[:, 3] - t[:, 2]), (t[:, 3] - t[:, 2]).max()))
