"""Summarise an .ncu-rep (first kernel): headline metrics, stall reasons, hottest SASS lines.
    python tools/ncu_summary.py gpurun_out/x.ncu-rep [n_lines]
"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
nl = int(sys.argv[2]) if len(sys.argv) > 2 else 25
kidx = int(sys.argv[3]) if len(sys.argv) > 3 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2 + kidx]
d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
print("kernel:", d.get("Kernel Name", ("", ""))[1][:100])
for k in ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
          "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
          "smsp__inst_executed.sum", "smsp__cycles_active.avg", "smsp__cycles_elapsed.avg", "smsp__issue_active.avg.pct_of_peak_sustained_active",
          "smsp__warps_eligible.avg.per_cycle_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
          "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
          "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
          "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active",
          "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active",
          "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_op_hmma.sum",
          "sm__inst_executed_pipe_uniform.sum", "sm__inst_executed_pipe_tmem.sum",
          "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
          "sm__cycles_elapsed.max", "gpc__cycles_elapsed.avg.per_second"]:
    if k in d:
        print("  %-70s %s %s" % (k, d[k][1], d[k][0]))
st = [(h, float(v.replace(",", ""))) for h, (u, v) in d.items() if h.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in h and v.replace(",", "").replace(".", "").isdigit()]
tot = sum(v for _, v in st) or 1
print("stall samples (all):")
for h, v in sorted(st, key=lambda x: -x[1])[:12]:
    print("  %-40s %7.0f  %5.1f%%" % (h.replace("smsp__pcsamp_warps_issue_stalled_", ""), v, 100 * v / tot))
# the source page of a multi-kernel report is one section per kernel, each introduced by a ("Kernel Name", name) row:
# take the section of THIS kernel only (round 1 attributed another kernel's instructions to the first one)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
sections, cur = [], None
for r in rows:
    if len(r) >= 2 and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        sections.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
want = d.get("Kernel Name", ("", ""))[1]
# raw-page names may be abbreviated ("nn_fwd_kernel") while the source page prints the full signature
match = [sec for sec in sections if want.split("(")[0] in sec["name"]]
sec = match[min(len(match) - 1, 0)] if match else (sections[kidx] if kidx < len(sections) else None)
if sec is None or len(sec["rows"]) < 2:
    sys.exit(0)
h2 = sec["rows"][0]; ix = {h: i for i, h in enumerate(h2)}
data = [r for r in sec["rows"][1:] if len(r) == len(h2) and r[ix["# Samples"]].isdigit()]
tot = sum(int(r[ix["# Samples"]]) for r in data) or 1
print("hottest instructions of %s (samples, %%, executed, dominant stalls):" % sec["name"][:60])
cols = [c for c in h2 if c.startswith("stall_") and "Not Issued" not in c]
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:nl]:
    stalls = sorted(((int(r[ix[c]]), c[6:]) for c in cols), reverse=True)[:2]
    print("  %5s %4.1f%% %9s  %-28s %s" % (r[ix["# Samples"]], 100 * int(r[ix["# Samples"]]) / tot, r[ix["Instructions Executed"]],
                                          ",".join("%s:%d" % (n, v) for v, n in stalls if v), r[ix["Source"]].strip()[:60]))
# instruction mix of the kernel (executed warp instructions by opcode)
from collections import Counter
mix = Counter()
for r in data:
    srcl = r[ix["Source"]].strip().split()
    op = (srcl[1] if srcl and srcl[0].startswith("@") and len(srcl) > 1 else (srcl[0] if srcl else "?")).split(".")[0]
    mix[op] += int(r[ix["Instructions Executed"]])
tot_i = sum(mix.values()) or 1
print("instruction mix (executed warp instructions): " + ", ".join("%s %.1f%%" % (k, 100.0 * v / tot_i) for k, v in mix.most_common(12)))
