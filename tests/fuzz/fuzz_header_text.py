import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT)
import cloudini_b200 as cb
from cloudini_b200 import synth
from oracle.client import RefOracle
ref = RefOracle()
rng = np.random.default_rng(int(sys.argv[1]))
info, cloud = synth.cloud_c2(300, seed=3)
blob = ref.encode(info, cloud)
dinfo, hdr = cb.DecodeHeader(blob)
n = cloud.size
bad = 0; shielded = 0; overlapping = 0
seen = set()
for t in range(int(sys.argv[2])):
    b = bytearray(blob)
    i = int(rng.integers(0, hdr))
    if rng.integers(0, 2): b[i] ^= 1 << int(rng.integers(0, 8))
    else: b[i] = int(rng.integers(32, 127))
    b = bytes(b)
    # the reference's decoders store at offset + i * point_step unchecked: a forged offset makes it write past `want` and
    # corrupt this process's heap, so only hand it headers whose fields fit a point (ours must still survive the others)
    safe = True
    try:
        di0, _ = cb.DecodeHeader(b)
        for f in di0.fields:
            ft = int(f.type)
            if ft < 1 or ft > 10 or f.offset + cb.SizeOf(cb.FieldType(ft)) > di0.point_step: safe = False
        if di0.width * di0.height * di0.point_step > n: safe = False
    except RuntimeError: pass
    if not safe:
        shielded += 1
        try: cb.PointcloudDecoder().decode(di0, b[_:], np.full(n, 0x33, np.uint8))
        except RuntimeError: pass
        continue
    want = np.full(n, 0x33, np.uint8); rok = True
    try: ref.decode(b, want)
    except RuntimeError as e: rok = False; rerr = str(e)
    got = np.full(n, 0x33, np.uint8); ook = True
    try:
        di, h = cb.DecodeHeader(b)
        if di.width * di.height * di.point_step > n: raise RuntimeError("too big for the test buffer")
        cb.PointcloudDecoder().decode(di, b[h:], got)
    except RuntimeError as e: ook = False; oerr = str(e)
    if rok and ook and not np.array_equal(got, want):
        # known limitation (DESIGN.md, weak spots): fields that OVERLAP inside a point are stored by the reference in field
        # order per point (last writer wins); the decode kernels do not promise a store order between fields
        spans = sorted((f.offset, f.offset + cb.SizeOf(cb.FieldType(int(f.type)))) for f in di.fields)
        if any(a[1] > b_[0] for a, b_ in zip(spans, spans[1:])):
            overlapping += 1
            continue
    if rok != ook or (rok and not np.array_equal(got, want)):
        key = (rok, ook, bytes(blob[max(0,i-12):i]).decode('latin-1'))
        bad += 1
        if len(seen) < 25 and key not in seen:
            seen.add(key)
            print("MISMATCH ref_ok", rok, "ours_ok", ook, "at", i, repr(bytes(blob[max(0,i-15):i+8])), "->", repr(b[max(0,i-15):i+8]), "|", (rerr if not rok else "")[:60], "|", (oerr if not ook else "")[:60])
print("trials", t + 1, "mismatches", bad, "kept from the reference (forged offsets / sizes)", shielded, "overlapping fields decoded differently", overlapping)
