"""Executed-instruction histogram from an `ncu --page source --csv` dump: per opcode, with the B200 dispatch-cost
classes of tools/sass_hist.py, plus the share of stall samples.  usage: ncu_src_hist.py src.csv [pixels]"""
import collections
import csv
import sys

FMA1 = ("FFMA", "FMUL", "FADD", "IMAD", "HFMA2", "HADD2", "HMUL2")
FMA2 = ("FFMA2", "FMUL2", "FADD2")
ALU = ("FMNMX", "FMNMX3", "FSEL", "FSETP", "IADD3", "LOP3", "ISETP", "SEL", "LEA", "SHF", "IMNMX", "VIMNMX", "VIMNMX3",
       "PRMT", "IABS", "FCHK", "I2FP", "F2FP", "PLOP3", "CSEL", "FSET", "VIADD", "IADD", "BMSK", "SGXT", "FLO", "POPC", "P2R", "R2P")


def classify(op):
    base = op.split(".")[0]
    return "fma2" if base in FMA2 else "fma1" if base in FMA1 else "alu" if base in ALU else "mov" if base == "MOV" else "other"


rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isrc, iex, isamp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
px = float(sys.argv[2]) if len(sys.argv) > 2 else 4928.0 * 3264.0
ops, cls, samples = collections.Counter(), collections.Counter(), collections.Counter()
nstatic = 0
for r in rows[2:]:
    if len(r) <= iex or not r[iex]:
        continue
    t = r[isrc].strip().split()
    if not t:
        continue
    op = t[1] if t[0].startswith("@") else t[0]
    n = int(r[iex])
    nstatic += 1
    ops[op] += n
    cls[classify(op)] += n
    samples[op] += int(r[isamp] or 0)
tot = sum(ops.values())
print("static instructions %d; executed warp-instructions %d = %.1f per pixel (x32 threads)" % (nstatic, tot, tot * 32 / px))
print("classes per pixel:", {k: round(v * 32 / px, 1) for k, v in cls.items()})
print("dispatch-cycle model per pixel: math %.1f (fma1 + 2*fma2 + 2*alu), issue slots %.1f" % (
    (cls["fma1"] + 2 * cls["fma2"] + 2 * cls["alu"]) * 32 / px, tot * 32 / px))
tsamp = sum(samples.values())
for op, n in ops.most_common(45):
    print("  %-24s %7.1f /px  %5.1f%% of samples  %s" % (op, n * 32 / px, 100.0 * samples[op] / max(tsamp, 1), classify(op)))
