"""Development aid: per-CUDA-source-line instruction and stall-sample shares from an ncu report (--import-source on)."""
import csv, io, subprocess, sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
fname, hdr, lines = "", None, []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr and r and r[0].isdigit():
        try:
            lines.append((fname, int(r[0]), r[1], int(r[hdr.index("# Samples")]), int(r[hdr.index("Instructions Executed")])))
        except ValueError:
            pass
ti = sum(l[4] for l in lines) or 1
ts = sum(l[3] for l in lines) or 1
print(f"total warp-instr {ti}  samples {ts}")
for f, n, src, s, i in sorted(lines, key=lambda l: -l[4])[:top]:
    print(f"{f}:{n:4d} inst {100*i/ti:5.1f}%  samp {100*s/ts:5.1f}%  {src.strip()[:110]}")
