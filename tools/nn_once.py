"""A few eager Chamfer forwards + fused steps on one input family (for an ncu launch list).
    python tools/nn_once.py [randn|chair|dups] [B N M] [--lib path/to/other/libpnae.so (forward only, raw ctypes)]
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from pointnet_autoencoder_b200 import synthetic

args = [a for a in sys.argv[1:] if not a.startswith("--")]
gen = args[0] if args else "randn"
b, n, m = (int(args[1]), int(args[2]), int(args[3])) if len(args) >= 4 else (32, 2048, 2048)
if gen == "randn":
    x1n, x2n = synthetic.s_randn(b, n, m)
elif gen == "chair":
    x2n, x1n = synthetic.s_chair(b, max(n, m)); x1n = np.ascontiguousarray(x1n[:, :n]); x2n = np.ascontiguousarray(x2n[:, :m])
else:
    rs = np.random.RandomState(0)
    src = rs.uniform(-1, 1, (b, m, 3)).astype(np.float32)
    x2n = np.ascontiguousarray(np.take_along_axis(src, rs.randint(0, m // 2, (b, m))[:, :, None].repeat(3, 2), 1))
    x1n = np.ascontiguousarray(np.take_along_axis(x2n, rs.randint(0, m, (b, n))[:, :, None].repeat(3, 2), 1) + (rs.randn(b, n, 3) * 0.02).astype(np.float32))
x1 = torch.from_numpy(x1n).cuda(); x2 = torch.from_numpy(x2n).cuda()
libarg = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--lib=")]
if libarg:
    lib = C.CDLL(libarg[0])
    lib.pnae_nn_distance_workspace_bytes.restype = C.c_size_t
    wsb = lib.pnae_nn_distance_workspace_bytes(b, n, m)
    ws = torch.empty((wsb,), dtype=torch.uint8, device="cuda")
    d1 = torch.empty((b, n), device="cuda"); i1 = torch.empty((b, n), dtype=torch.int32, device="cuda")
    d2 = torch.empty((b, m), device="cuda"); i2 = torch.empty((b, m), dtype=torch.int32, device="cuda")
    p = lambda t: C.c_void_p(t.data_ptr())
    reps = 200 if "--time" in sys.argv else 3
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    for it in range(2):
        e0.record()
        for _ in range(reps):
            rc = lib.pnae_nn_distance_fwd(b, n, p(x1), C.c_int(m), p(x2), p(d1), p(i1), p(d2), p(i2), p(ws), C.c_size_t(wsb), None)
            assert rc == 0
        e1.record(); e1.synchronize()
    print("%s %s fwd (eager back-to-back launches on the legacy stream) %.2f us" % (libarg[0], gen, e0.elapsed_time(e1) / reps * 1e3))
else:
    from pointnet_autoencoder_b200 import ops
    g1 = torch.full((b, n), 100.0 / (b * n), device="cuda"); g2 = torch.full((b, m), 100.0 / (b * m), device="cuda")
    for _ in range(3):
        ops.nn_distance_fwd(x1, x2)
        ops.nn_distance_fwd_grad(x1, x2, g1, g2)
    torch.cuda.synchronize()
    print("ok")
