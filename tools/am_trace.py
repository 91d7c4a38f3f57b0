"""Per-sweep timeline of three approx_match CTAs (needs a libpnae built with -DPNAE_AM_TRACE, see tools/build_variant.sh)."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from pointnet_autoencoder_b200 import _lib, ops, synthetic

b, n = 32, 2048
x2n, x1n = synthetic.s_chair(b, n)
x1 = torch.from_numpy(np.ascontiguousarray(x1n)).cuda(); x2 = torch.from_numpy(np.ascontiguousarray(x2n)).cuda()
for _ in range(3):
    ops.approx_match_factors(x1, x2)
torch.cuda.synchronize()
buf = np.zeros((3, 64), np.int64)
assert _lib.load().pnae_debug_am_trace(buf.ctypes.data_as(ctypes.c_void_p)) == 0
for cta in range(3):
    t = (buf[cta] - buf[cta, 0]) / 1.925e3
    # stamps: [A0 start, A0 end] then per level [B start, B end, CA start, CA end] ..., then last level start, end
    print("CTA %d: A0 %.1f us" % (cta, t[1] - t[0]))
    i = 2
    sweeps, waits = [t[1] - t[0]], []
    prev_end = t[1]
    lev = 0
    while i + 3 < 64 and buf[cta, i + 3]:
        bs, be, cs, ce = t[i:i + 4]
        waits += [bs - prev_end, cs - be]
        sweeps += [be - bs, ce - cs]
        print("   level %d: barrier %.1f  B %.1f  barrier %.1f  C/A %.1f" % (lev, bs - prev_end, be - bs, cs - be, ce - cs))
        prev_end = ce; i += 4; lev += 1
    print("   last: barrier %.1f, last level %.1f; total %.1f us: sweeps %.1f, barriers %.1f" % (t[i] - prev_end, t[i + 1] - t[i], t[i + 1], sum(sweeps), sum(waits) + t[i] - prev_end))

allb = np.zeros((1024, 4), np.int64)
assert _lib.load().pnae_debug_am_all(allb.ctypes.data_as(ctypes.c_void_p)) == 0
plan = (ctypes.c_int * 8)()
assert _lib.load().pnae_approx_match_plan(b, n, n, torch.cuda.get_device_properties(0).multi_processor_count, plan) == 0
g = plan[0]
d = allb[:g]
print("per-CTA duration of level 3's B sweep (us): min %.1f  median %.1f  max %.1f" % (d[:, 1].min() / 1.925e3, np.median(d[:, 1]) / 1.925e3, d[:, 1].max() / 1.925e3))
print("CTA: smid B C/A (us)")
for i in list(range(0, 8)) + list(range(g // 2 - 4, g // 2 + 4)) + list(range(g - 8, g)):
    print("  %3d: sm %3d  %5.1f %5.1f" % (i, d[i, 0], d[i, 1] / 1.925e3, d[i, 2] / 1.925e3))
# do the two CTAs of an SM finish together?
from collections import defaultdict
per = defaultdict(list)
for i in range(g):
    per[int(d[i, 0])].append((i, d[i, 1] / 1.925e3))
print("CTAs per SM:", sorted(set(len(v) for v in per.values())))
slow = sorted(per.items(), key=lambda kv: -max(x[1] for x in kv[1]))[:6]
print("slowest SMs:", slow)
fast = sorted(per.items(), key=lambda kv: max(x[1] for x in kv[1]))[:6]
print("fastest SMs:", fast)
