import sys, os, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT)
import cloudini_b200 as cb
from cloudini_b200 import synth
from oracle.client import RefOracle, PortOracle
ref = RefOracle(); port = PortOracle()
rng = np.random.default_rng(int(sys.argv[1]))
info, cloud = synth.cloud_lossless(400, seed=3, lossless=False)
blob = ref.encode(info, cloud)
dinfo, hdr = cb.DecodeHeader(blob)
n = cloud.size
dec = cb.PointcloudDecoder()
stats = {"ok_same":0, "both_fail":0, "mismatch":0, "port_mismatch":0}
for t in range(int(sys.argv[2])):
    b = bytearray(blob)
    for _ in range(int(rng.integers(1, 3))):
        i = int(rng.integers(hdr + 4, len(b))); b[i] = int(rng.integers(0, 256))
    b = bytes(b)
    want = np.full(n, 0x33, np.uint8); rok = True
    try: ref.decode(b, want)
    except RuntimeError: rok = False
    pw = np.full(n, 0x33, np.uint8); pok = True
    try: port.decode(b, pw)
    except RuntimeError: pok = False
    if pok != rok or (rok and not np.array_equal(pw, want)): stats["port_mismatch"] += 1
    res = []
    for mode in ("par", "seq"):
        os.environ["CLDN_B200_MIXED_DECODE"] = mode
        got = np.full(n, 0x33, np.uint8); ook = True
        try: dec.decode(dinfo, b[hdr:], got)
        except RuntimeError: ook = False
        res.append((ook, got))
    bad = any(ook != rok or (rok and not np.array_equal(got, want)) for ook, got in res)
    if bad:
        stats["mismatch"] += 1
        if stats["mismatch"] <= 5: print("MISMATCH ref_ok", rok, "par_ok", res[0][0], "seq_ok", res[1][0])
    elif rok: stats["ok_same"] += 1
    else: stats["both_fail"] += 1
print(stats)
