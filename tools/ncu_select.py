#!/usr/bin/env python
"""Transpose `ncu -i rep.ncu-rep --page raw --csv` into a compact metric table (one column per launch of the
kernels whose name matches argv[2]) keeping the metrics the roofline discussion uses.
usage: python tools/ncu_select.py raw.csv k1_correlate > profiles/r1_k1_level0_ncu_full.csv"""
import csv
import re
import sys

KEEP = re.compile(
    r"gpu__time_duration|dram__bytes_(read|write)\.sum$|dram__throughput|inst_executed|issue_active|"
    r"warps_active|achieved_occupancy|registers_per_thread|shared_mem_per_block|pipe_(fma|alu|lsu|fp64|xu)|"
    r"average_warps?_issue_stalled|pcsamp_warps_issue_stalled|lts__t_sector.*hit_rate|l1tex__t_sector.*hit_rate|"
    r"l1tex__data_pipe_lsu_wavefronts|bank_conflicts|l1tex__t_(requests|sectors)_pipe_lsu_mem_global|"
    r"launch__(grid_size|block_size|occupancy_limit|waves)|sm__cycles_active.avg|smsp__cycles_active.avg|"
    r"sm__throughput|gpu__compute_memory_throughput|thread_inst_executed")

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr, units, data = rows[0], rows[1], rows[2:]
ki = hdr.index("Kernel Name")
sel = [r for r in data if re.search(sys.argv[2], r[ki])]
w = csv.writer(sys.stdout)
w.writerow(["metric", "unit"] + ["launch%d" % i for i in range(len(sel))])
w.writerow(["Kernel Name", ""] + [re.sub(r"\(.*", "", r[ki]) for r in sel])
for c, name in enumerate(hdr):
    if KEEP.search(name):
        w.writerow([name, units[c]] + [r[c] for r in sel])
