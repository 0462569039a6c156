"""Development aid: per-CUDA-source-line instruction and stall-sample shares of one kernel, from an ncu report
(--import-source on) or from the CSV that tools/gpu_ncu_all.sh extracts on the GPU box
(`ncu -i rep --page source --print-source cuda,sass --csv`).

  python tools/source_hot.py <rep-or-csv> [kernel-substring] [top] [launch-index]
"""
import csv, io, subprocess, sys

rep = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
which = int(sys.argv[4]) if len(sys.argv) > 4 else -1   # n-th launch of that kernel in the file (-1: all summed)
raw = open(rep).read() if rep.endswith(".csv") else subprocess.run(
    ["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
fname, func, hdr = "", "", None
acc = {}        # (file, line) -> [source, samples, instructions]
launch, seen_files = -1, set()
for r in csv.reader(io.StringIO(raw)):
    if len(r) == 2 and r[0] in ("File Path", "File Name"):
        fname = r[1].split("/")[-1]
        continue
    if len(r) == 2 and r[0] == "Function Name":
        if r[1] != func or fname in seen_files:
            func, seen_files = r[1], set()
            if want in func:
                launch += 1
        seen_files.add(fname)
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if not hdr or not r or not r[0].isdigit() or want not in func or (which >= 0 and launch != which):
        continue
    try:
        s, i = int(r[hdr.index("# Samples")]), int(r[hdr.index("Instructions Executed")])
    except (ValueError, IndexError):
        continue
    a = acc.setdefault((fname, int(r[0])), [r[1], 0, 0])
    a[1] += s
    a[2] += i
ti = sum(a[2] for a in acc.values()) or 1
ts = sum(a[1] for a in acc.values()) or 1
print(f"kernel ~ '{want}'  total warp-instr {ti}  samples {ts}")
for (f, n), (src, s, i) in sorted(acc.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{f}:{n:4d} samp {100*s/ts:5.1f}%  inst {100*i/ti:5.1f}%  {src.strip()[:120]}")
