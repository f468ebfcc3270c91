#!/usr/bin/env python
"""Static instruction census of a kernel object (no GPU needed): opcodes of every function of a cubin / .o,
grouped by the pipe that executes them on sm_100 (B300_MICROARCH.md: IMAD/FFMA on the fma pipe, IADD3/LOP3/SHF/
PRMT/VIMNMX on the alu pipe, both one warp instruction per two cycles per scheduler).

    python tools/sass_count.py cmsis-dsp_b200/build/ku_2_1024.o [--per N] [--match SUBSTR] [--ops]

--per N   divides by N (points per thread) so the figures read "instructions per point"
"""
import argparse
import collections
import re
import subprocess

FMA = ("IMAD", "FFMA", "FMUL", "FADD", "FFMA2", "FMUL2", "FADD2", "HFMA2", "HADD2", "HMUL2", "IDP", "DFMA", "DADD", "DMUL")
ALU = ("IADD3", "LOP3", "SHF", "PRMT", "VIMNMX", "VIADDMNMX", "VIADD", "LEA", "SEL", "ISETP", "FMNMX", "MOV", "SGXT", "IABS",
       "FSEL", "FSETP", "PLOP3", "I2IP", "BMSK", "FLO", "POPC", "IMNMX", "FCHK", "CS2R", "P2R", "R2P")
LSU = ("LDG", "STG", "LDS", "STS", "LDSM", "LD", "ST", "ATOM", "RED", "LDC", "LDCU", "UBLKCP", "UTMA", "SYNCS", "LDL", "STL")
XU = ("MUFU", "I2F", "F2I", "F2F", "I2I", "FRND")


def pipe_of(op):
    base = op.split(".")[0]
    if op.startswith("IMAD.HI"):
        return "xu"
    if base in FMA:
        return "fma"
    if base in XU:
        return "xu"
    if base in LSU:
        return "lsu"
    if base in ALU:
        return "alu"
    if base.startswith("U") or base in ("S2R", "S2UR", "BRA", "BAR", "EXIT", "NOP", "BSSY", "BSYNC", "WARPSYNC", "DEPBAR", "ERRBAR", "MEMBAR", "FENCE", "CCTL", "R2UR", "NANOSLEEP", "YIELD", "CALL", "RET", "SHFL", "VOTE", "MATCH", "ELECT"):
        return "other"
    return "other"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("obj")
    ap.add_argument("--per", type=float, default=1.0)
    ap.add_argument("--match", default="")
    ap.add_argument("--ops", action="store_true")
    a = ap.parse_args()
    sass = subprocess.run(["cuobjdump", "-sass", a.obj], capture_output=True, text=True).stdout
    fn, funcs = None, collections.OrderedDict()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            funcs[fn] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and fn:
            funcs[fn][m.group(1).rstrip(";")] += 1
    for fn, c in funcs.items():
        if a.match and a.match not in fn:
            continue
        pipes = collections.Counter()
        for op, n in c.items():
            pipes[pipe_of(op)] += n
        tot = sum(c.values())
        short = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()[:150]
        print(f"{short}\n   total {tot / a.per:.1f}  alu {pipes['alu'] / a.per:.1f}  fma {pipes['fma'] / a.per:.1f}  xu {pipes['xu'] / a.per:.1f}  "
              f"lsu {pipes['lsu'] / a.per:.1f}  other {pipes['other'] / a.per:.1f}")
        if a.ops:
            print("   " + "  ".join(f"{op}:{n}" for op, n in c.most_common(28)))


if __name__ == "__main__":
    main()
