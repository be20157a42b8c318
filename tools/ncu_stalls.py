"""Summary of one `ncu --set full` report of a kernel: duration, occupancy, issue utilisation, stall reasons per issue,
and the most-stalled SASS instructions.  usage: ncu_stalls.py report.ncu-rep [pixels]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
px = float(sys.argv[2]) if len(sys.argv) > 2 else 4928.0 * 3264.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, vals = rows[0], rows[-1]
d = dict(zip(hdr, vals))
for k in ("gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
          "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
          "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum"):
    print("%-70s %s" % (k, d.get(k)))
st = sorted(((float(v), k) for k, v in d.items() if "issue_stalled" in k and k.endswith("per_issue_active.ratio")), reverse=True)
print("stalls per issue:", ", ".join("%s %.2f" % (k.split("issue_stalled_")[1].split("_per_issue")[0], v) for v, k in st[:9]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]
ia, isrc, iex, isamp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
ins = [(r[ia], r[isrc].strip(), int(r[iex] or 0), int(r[isamp] or 0)) for r in rows[2:] if len(r) > iex and r[iex] != ""]
tot = sum(x[3] for x in ins)
print("executed warp-instructions %d = %.1f per pixel" % (sum(x[2] for x in ins), sum(x[2] for x in ins) * 32 / px))
for i in sorted(range(len(ins)), key=lambda i: -ins[i][3])[:8]:
    print("  #%d %s  %-60s samples %.1f%%" % (i, ins[i][0][-5:], ins[i][1][:60], 100.0 * ins[i][3] / max(tot, 1)))
