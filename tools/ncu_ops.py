#!/usr/bin/env python
"""tools/ncu_ops.py -- SASS opcode histogram of a kernel in an ncu report (needs --import-source on),
weighted by executed warp instructions, grouped by the issue pipe each opcode uses.

    python tools/ncu_ops.py gpurun_out/prof.ncu-rep k_fast_cells [top_n]

Pipe assignment follows /opt/skills/guides/B300_MICROARCH.md: FFMA/FMUL/FADD/IMAD/HFMA2/IDP on the fma pipe;
IADD3/LOP3/SHF/PRMT/*MNMX/ISETP/SEL/... on the alu pipe; both issue one warp instruction per two cycles per
SM sub-partition, so the busier of the two bounds an integer kernel.
"""
import collections
import csv
import io
import re
import subprocess
import sys

FMA = ("FFMA", "FMUL", "FADD", "IMAD", "HFMA2", "HMUL2", "HADD2", "IDP", "IMUL", "FSWZADD")
LSU = ("LDS", "STS", "LDG", "STG", "LD", "ST", "LDGSTS", "ATOM", "ATOMS", "ATOMG", "RED", "LDSM", "LDGDEPBAR", "LDL", "STL", "LDC", "LDCU")
XU = ("MUFU", "F2I", "I2F", "F2F", "I2I", "POPC", "FLO", "BREV", "F2IP", "I2FP", "FRND")
CTRL = ("BRA", "EXIT", "BSSY", "BSYNC", "WARPSYNC", "BAR", "NOP", "DEPBAR", "CALL", "RET", "YIELD", "BMOV", "ERRBAR", "MEMBAR", "BRX", "JMP", "NANOSLEEP")
UNI = ("S2R", "S2UR", "CS2R", "R2UR", "UMOV", "ULDC", "UIADD3", "ULOP3", "USHF", "UIMAD", "UISETP", "USEL", "ULEA", "UPRMT", "UFLO", "UPOPC", "VOTEU", "R2P", "P2R", "UP2UR", "UR2UP", "UMOV32I")


def pipe(op):
    base = op.split(".")[0]
    if base in FMA:
        return "fma"
    if base in LSU:
        return "lsu"
    if base in XU:
        return "xu"
    if base in CTRL:
        return "ctrl"
    if base in UNI or base.startswith("U"):
        return "uniform"
    if base in ("SHFL", "VOTE", "MATCH", "REDUX"):
        return "shfl/vote"
    return "alu"


def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kernel],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hi = [i for i, r in enumerate(rows) if "Address" in r and "Instructions Executed" in r][0]
    hdr = rows[hi]
    ci, cs = hdr.index("Instructions Executed"), hdr.index("Source")
    ops = collections.Counter()
    for r in rows[hi + 1:]:
        if len(r) <= ci or r[0].startswith("Kernel Name"):
            if r and r[0].startswith("Kernel Name"):
                break       # first matching kernel instance only
            continue
        try:
            n = int(r[ci])
        except ValueError:
            continue
        s = r[cs].strip()
        s = re.sub(r"^@!?U?P\d+\s+", "", s)
        op = s.split()[0].rstrip(";") if s else "?"
        ops[op] += n
    tot = sum(ops.values()) or 1
    pipes = collections.Counter()
    for op, n in ops.items():
        pipes[pipe(op)] += n
    print("kernel %s: %d warp instructions" % (kernel, tot))
    print("by pipe: " + "  ".join("%s %.1f%%" % (p, 100.0 * n / tot) for p, n in pipes.most_common()))
    for op, n in ops.most_common(top):
        print("%6.2f%%  %-9s %s" % (100.0 * n / tot, pipe(op), op))


if __name__ == "__main__":
    main()
