#!/usr/bin/env python3
"""Static instruction budget of a kernel from `nvdisasm -g` output: instructions per source-line range and per issue pipe.

No GPU needed: `cuobjdump -xelf all cloudini_b200/lib/libcloudini_b200.so` gives the cubins, `nvdisasm -g <cubin>` the
annotated SASS. For straight-line (unrolled, predicated) hot paths the static count of the lines on the path equals the
dynamic count per thread, which is what ncu's `smsp__inst_executed` measures per warp.

  python tools/sass_budget.py <nvdisasm -g output> <mangled-name substring> <file.cu> name=lo-hi [name=lo-hi ...]

Pipes (B300_MICROARCH.md "Pipe rates"): fma = FFMA/FMUL/IMAD*/..., alu = IADD3/LOP3/SHF/PRMT/SEL/ISETP/FMNMX/..., both one
warp instruction per 2 cycles per SM sub-partition; xu = F2I/I2F/FLO/BREV/MUFU/POPC (quarter rate); lsu = LD*/ST*/ATOM*;
the rest (branches, barriers, shuffles, uniform datapath) is `other`.
"""
import collections
import re
import sys

FMA = {"IMAD", "FFMA", "FMUL", "FADD", "HFMA2", "HMUL2", "HADD2", "IDP", "IDP4A"}
ALU = {"IADD3", "IADD", "LOP3", "LOP", "SHF", "SHL", "SHR", "PRMT", "SEL", "ISETP", "FSETP", "FMNMX", "VIADD", "VIMNMX", "LEA", "PLOP3",
       "IMNMX", "MOV", "FSEL", "CS2R", "BMSK", "SGXT", "IABS", "P2R", "R2P", "VABSDIFF", "VABSDIFF4", "FCHK", "R2UR"}
XU = {"F2I", "I2F", "FLO", "BREV", "MUFU", "POPC", "F2F", "I2I", "F2FP", "FRND"}


def pipe(op):
    base = op.split(".")[0]
    if base in FMA:
        return "fma"
    if base in ALU:
        return "alu"
    if base in XU:
        return "xu"
    if base.startswith(("LD", "ST", "ATOM", "RED", "CCTL", "MEMBAR", "ERRBAR")):
        return "lsu"
    return "other"


def main():
    path, fn, src = sys.argv[1], sys.argv[2], sys.argv[3]
    ranges = []
    for a in sys.argv[4:]:
        name, r = a.split("=")
        lo, hi = r.split("-")
        ranges.append((name, int(lo), int(hi)))
    inside, line, cur_file = False, 0, ""
    per = collections.defaultdict(lambda: collections.Counter())
    ops = collections.defaultdict(lambda: collections.Counter())
    ins_re = re.compile(r"^\s*/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_.]*)")
    for ln in open(path, errors="replace"):
        if ln.startswith("//---------------------"):
            inside = fn in ln and ".text." in ln
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur_file, line = m.group(1), int(m.group(2))
            continue
        m = ins_re.match(ln)
        if not m:
            continue
        op = m.group(1)
        bucket = "elsewhere"
        if cur_file.endswith(src):
            for name, lo, hi in ranges:
                if lo <= line <= hi:
                    bucket = name
                    break
        else:
            bucket = "inlined:" + cur_file.rsplit("/", 1)[-1]
        per[bucket][pipe(op)] += 1
        ops[bucket][op.split(".")[0]] += 1
    print(f"{'bucket':28s} {'total':>6s} {'fma':>6s} {'alu':>6s} {'xu':>6s} {'lsu':>6s} {'other':>6s}")
    for b, c in per.items():
        print(f"{b:28s} {sum(c.values()):6d} {c['fma']:6d} {c['alu']:6d} {c['xu']:6d} {c['lsu']:6d} {c['other']:6d}")
    for b, c in ops.items():
        print(f"-- {b}: " + " ".join(f"{k}:{v}" for k, v in c.most_common(14)))


if __name__ == "__main__":
    main()
