#!/usr/bin/env python3
"""Static SASS opcode histogram of the shipped kernels (no GPU needed): `cuobjdump -sass` of cloudini_b200/lib/libcloudini_b200.so,
per kernel the instruction count by issue pipe and the most frequent opcodes. Static counts: the hot paths are unrolled,
predicated straight-line code inside the tile loops, so the proportions are those of the dynamic stream up to the rare-path
code (exact paths, error reporting) that is compiled in but does not run on plain clouds; ncu's per-pipe utilisation of the
same kernels is in profiles/r2_kernels.json.

  python tools/sass_histogram.py [substring ...] > profiles/r2_sass_histogram.json
"""
import collections, json, os, re, subprocess, sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from sass_budget import pipe  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "cloudini_b200", "lib", "libcloudini_b200.so")


def main():
    want = sys.argv[1:]
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    out, name, ops = {}, None, None
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if name:
                out[name] = ops
            name, ops = m.group(1), collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and name:
            ops[m.group(1)] += 1
    if name:
        out[name] = ops
    res = {}
    for mangled, ops in out.items():
        dem = subprocess.run(["c++filt", mangled], capture_output=True, text=True).stdout.strip().split("(")[0]
        if want and not any(w in dem for w in want):
            continue
        total = sum(ops.values())
        pipes = collections.Counter()
        base = collections.Counter()
        for op, c in ops.items():
            pipes[pipe(op)] += c
            base[op.split(".")[0]] += c
        res[dem] = {"instructions": total, "by_pipe": dict(pipes.most_common()),
                    "by_pipe_share": {k: round(v / total, 3) for k, v in pipes.most_common()},
                    "top_opcodes": dict(base.most_common(24))}
    json.dump(res, sys.stdout, indent=1)


if __name__ == "__main__":
    main()
