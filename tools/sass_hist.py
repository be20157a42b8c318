"""Opcode histogram of a kernel's SASS, per basic region between branch targets, with a pipe-cost model measured on
B200 (tools/ubench/pipes.cu): FMA-pipe single ops 1 dispatch cycle, packed f32x2 and every ALU-pipe op (FMNMX, FSEL,
FSETP, IADD3, LOP3, ISETP, ...) 2 cycles, other pipes (LSU, XU, FP64, branch, uniform) 1 issue slot that can hide
in the shadow of a 2-cycle op.   usage: sass_hist.py <sass file> <function substring> [addr_lo addr_hi]"""
import collections
import re
import sys

FMA1 = ("FFMA", "FMUL", "FADD", "IMAD", "HFMA2", "HADD2", "HMUL2")
FMA2 = ("FFMA2", "FMUL2", "FADD2")
ALU = ("FMNMX", "FMNMX3", "FSEL", "FSETP", "IADD3", "LOP3", "ISETP", "SEL", "LEA", "SHF", "IMNMX", "VIMNMX", "VIMNMX3",
       "PRMT", "IABS", "FCHK", "I2FP", "F2FP", "PLOP3", "CSEL", "FSET", "VIADD", "IADD", "BMSK", "SGXT", "FLO", "POPC", "P2R", "R2P")


def classify(op):
    base = op.split(".")[0]
    if base in FMA2:
        return "fma2"
    if base in FMA1:
        return "fma1"
    if base in ALU:
        return "alu"
    if base == "MOV":
        return "mov"
    return "other"


def main():
    path, pat = sys.argv[1], sys.argv[2]
    lo = int(sys.argv[3], 16) if len(sys.argv) > 3 else 0
    hi = int(sys.argv[4], 16) if len(sys.argv) > 4 else 1 << 30
    cur, ins = None, []
    for line in open(path):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
        if m and cur and pat in cur:
            a = int(m.group(1), 16)
            if lo <= a < hi:
                t = m.group(2).strip().split()
                op = t[1] if t[0].startswith("@") else t[0]
                ins.append(op)
    ops = collections.Counter(ins)
    cls = collections.Counter()
    for op, n in ops.items():
        cls[classify(op)] += n
    total = sum(ops.values())
    cost = cls["fma1"] + 2 * cls["fma2"] + 2 * cls["alu"]
    print("instructions %d  classes %s" % (total, dict(cls)))
    print("dispatch-cycle model: math %d (fma1 + 2*fma2 + 2*alu); issue slots %d" % (cost, total))
    for op, n in ops.most_common(40):
        print("  %-22s %5d  %s" % (op, n, classify(op)))


main()
