"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): launches, mean and share per kernel.
    python tools/launch_summary.py gpurun_out/launches_r1.csv "<command that was profiled>" > profiles/rN_launches_summary.csv
"""
import csv
import re
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 14 and r[0].isdigit()]
agg = OrderedDict()
for r in rows:
    name = re.sub(r"\(.*", "", r[4]).replace("void ", "").strip()
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += float(r[-1].replace(",", ""))
tot = sum(a[1] for a in agg.values()) or 1.0
print("# ncu --metrics gpu__time_duration.sum --clock-control none, `%s`" % (sys.argv[2] if len(sys.argv) > 2 else "?"))
print("# per-launch times are cold-cache and serialised (no programmatic-launch overlap): compare SHARES, not absolutes")
print("kernel,launches,mean_ns,total_ns,share")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%s,%d,%.0f,%.0f,%.3f" % (k[-60:], n, t / n, t, t / tot))
