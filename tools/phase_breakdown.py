"""Development aid: per-phase instruction / stall-sample shares of one kernel from an ncu report (--import-source on).
   python tools/phase_breakdown.py <rep> <source file name> <points per launch> name=lo-hi ..."""
import csv, io, subprocess, sys
rep, fn, npts = sys.argv[1], sys.argv[2], float(sys.argv[3])
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
fname, hdr, lines = "", None, []
for r in rows:
    if len(r) == 2 and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr and r and r[0].isdigit():
        try: lines.append((fname, int(r[0]), int(r[hdr.index("# Samples")]), int(r[hdr.index("Instructions Executed")]), r[1]))
        except ValueError: pass
ti = sum(l[3] for l in lines); ts = sum(l[2] for l in lines)
for a in sys.argv[4:]:
    name, rng = a.split("="); lo, hi = map(int, rng.split("-"))
    i = sum(l[3] for l in lines if l[0] == fn and lo <= l[1] <= hi); s = sum(l[2] for l in lines if l[0] == fn and lo <= l[1] <= hi)
    print(f"{name:10s} inst/pt {i*32/npts:6.1f}  inst% {100*i/ti:5.1f}  samp% {100*s/ts:5.1f}")
oth = {}
for l in lines:
    if l[0] != fn: oth[l[0]] = oth.get(l[0], [0, 0]); oth[l[0]][0] += l[3]; oth[l[0]][1] += l[2]
print({k: (round(v[0]*32/npts, 1), round(100*v[1]/ts, 1)) for k, v in oth.items()})
print("total inst/pt", round(ti*32/npts, 1))
print("-- top lines by stall samples")
for f, n, s, i, src in sorted(lines, key=lambda l: -l[2])[:14]:
    print(f"{f}:{n:4d} samp {100*s/ts:5.1f}% inst/pt {i*32/npts:5.1f}  {src.strip()[:90]}")
