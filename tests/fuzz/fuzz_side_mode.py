"""V5 sections ahead of the regular stream (side mode) against sections behind it and against the reference, on damaged
blobs of padded / multi-section layouts: same outcome (decodes or not) and, when the blob decodes, the same bytes.
Usage: python tests/fuzz/fuzz_side_mode.py <seed> <trials>"""
import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import cloudini_b200 as cb
from cloudini_b200 import synth
from oracle.client import RefOracle
from test_gpu_fast_paths import _int_sections_layout, _xyzirt
ref = RefOracle()
seed, trials = int(sys.argv[1]), int(sys.argv[2])
rng = np.random.default_rng(seed)
os.environ["CLDN_B200_DECODE_MODE"] = "seq"
stats = {"ok_same": 0, "both_fail": 0, "both_fail_other_message": 0, "mismatch": 0}
for trial in range(trials):
    n = int(rng.choice([700, 4097, 32768, 32769, 40_000, 70_001]))
    which = int(rng.integers(0, 4))
    info, cloud = (synth.cloud_c3(n, seed=seed + trial, version=5), _int_sections_layout(n, seed + trial), _xyzirt(n, seed + trial),
                   _int_sections_layout(n, seed + trial, with_u64=True))[which]
    step = info.point_step
    blob = ref.encode(info, cloud)
    hdr = len(cb.PointcloudEncoder(info).getHeader())
    size0 = int(np.frombuffer(blob[hdr:hdr + 4], dtype=np.uint32)[0])
    bad = bytearray(blob)
    kind = int(rng.integers(0, 7))
    if kind == 0:
        bad = bad[:hdr + int(rng.integers(5, len(blob) - hdr))]
    elif kind == 1:
        for _ in range(int(rng.integers(1, 5))):
            bad[hdr + 4 + int(rng.integers(0, len(blob) - hdr - 4))] = int(rng.integers(0, 256))
    elif kind == 2:
        for _ in range(3):
            bad[hdr + 4 + size0 - 1 - int(rng.integers(0, min(size0 - 1, 3000)))] = int(rng.integers(0, 256))
    elif kind == 3:
        bad[hdr + 4 + int(rng.integers(0, max(1, size0 - 1)))] ^= 0x80
    elif kind == 4:
        bad[hdr:hdr + 4] = np.uint32(max(0, size0 + int(rng.integers(-3, 4)))).tobytes()
    elif kind == 5:
        bad[hdr + 4 + int(rng.integers(0, max(1, size0 // 2)))] = 0
    # kind 6: undamaged
    res = []
    for flag in ("1", "0"):
        os.environ["CLDN_B200_DECODE_SIDE"] = flag
        out = np.full(n * step, 0x33, dtype=np.uint8)
        try:
            cb.PointcloudDecoder().decode(info, bytes(bad[hdr:]), out)
            res.append(("ok", out.tobytes()))
        except RuntimeError as e:
            res.append(("err", str(e)))
    w = np.full(n * step, 0x33, dtype=np.uint8)
    try:
        ref.decode_payload(info, bytes(bad[hdr:]), w)
        r = ("ok", w.tobytes())
    except RuntimeError:
        r = ("err", "")
    same = res[0] == res[1] and res[0][0] == r[0] and (r[0] == "err" or res[0][1] == r[1])
    if not same and res[0][0] == res[1][0] == r[0] == "err":
        # a blob damaged in several places: which of its defects is reported first is not defined
        stats["both_fail_other_message"] += 1
    elif not same:
        stats["mismatch"] += 1
        print("MISMATCH trial", trial, "layout", which, "n", n, "kind", kind, res[0][0], res[1][0], r[0], "|", res[0][1][:90] if res[0][0] == "err" else "", "|", res[1][1][:90] if res[1][0] == "err" else "")
    elif r[0] == "ok":
        stats["ok_same"] += 1
    else:
        stats["both_fail"] += 1
print(stats)
