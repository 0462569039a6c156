import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT)
import cloudini_b200 as cb
from cloudini_b200 import synth, FieldType as F
from oracle.client import RefOracle
ref = RefOracle()
rng = np.random.default_rng(int(sys.argv[1]))
def rand_floats(n, kind):
    if kind == 0: return rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32).view(np.float32)          # every bit pattern
    if kind == 1: return (rng.standard_normal(n) * 10.0 ** rng.integers(-6, 7, n)).astype(np.float32)            # every magnitude
    v = np.cumsum(rng.normal(0, 0.01, n)).astype(np.float32)                                                   # smooth + specials
    idx = rng.integers(0, n, max(1, n // 20)); v[idx] = rng.choice(np.array([np.nan, np.inf, -np.inf, -0.0, 2147483.6, -2147483.7, 2.5e-4, 5e-4, 1.5e-3, 1e-45, 3.4e38], dtype=np.float32), idx.size)
    return v
bad = 0; tot = 0
for trial in range(int(sys.argv[2])):
    n = int(rng.choice([1, 33, 2047, 2049, 5000, 32769, 40000]))
    layout = int(rng.integers(0, 4))
    res = float(rng.choice([0.001, 0.01, 0.5, 1e-6, 100.0]))
    if layout == 0:   # XYZI fast path
        info = synth.info_xyzi(n, res); cols = 4; step = 16
    elif layout == 1: # XYZ
        info = synth.info_xyz(n, res); cols = 3; step = 12
    elif layout == 2: # two lossy floats (scalar path) + f64 lossy
        info = cb.EncodingInfo(width=n, height=1, point_step=16, compression_opt=cb.CompressionOption.NONE, use_threads=False)
        info.fields = [cb.PointField("a", 0, F.FLOAT32, res), cb.PointField("b", 4, F.FLOAT32, res * 2), cb.PointField("d", 8, F.FLOAT64, res)]
        cols = 2; step = 16
    else:             # XYZI + f64 Gorilla + u8
        info = cb.EncodingInfo(width=n, height=1, point_step=26, compression_opt=cb.CompressionOption.NONE, use_threads=False)
        info.fields = [cb.PointField("x", 0, F.FLOAT32, res), cb.PointField("y", 4, F.FLOAT32, res), cb.PointField("z", 8, F.FLOAT32, res),
                       cb.PointField("i", 12, F.FLOAT32, res), cb.PointField("t", 16, F.FLOAT64, None), cb.PointField("u", 24, F.UINT8, None)]
        cols = 4; step = 26
    buf = rng.integers(0, 256, (n, step), dtype=np.uint8)
    kind = int(rng.integers(0, 3))
    for c in range(cols): buf[:, 4*c:4*c+4] = rand_floats(n, kind).view(np.uint8).reshape(n, 4)
    if layout == 2: buf[:, 8:16] = (rand_floats(n, kind).astype(np.float64) * (1e10 if kind == 1 else 1.0)).view(np.uint8).reshape(n, 8)
    cloud = buf.reshape(-1)
    try: want = ref.encode(info, cloud)
    except RuntimeError: continue
    tot += 1
    got = cb.PointcloudEncoder(info).encode(cloud)
    ok = got == want
    if ok:
        dinfo, hdr = cb.DecodeHeader(got)
        w = np.full(cloud.size, 0x17, np.uint8); ref.decode(want, w)
        g = np.full(cloud.size, 0x17, np.uint8); cb.PointcloudDecoder().decode(dinfo, got[hdr:], g)
        ok = np.array_equal(g, w)
    if not ok: bad += 1; print("MISMATCH layout", layout, "n", n, "res", res, "kind", kind, len(got), len(want))
print("cases", tot, "mismatches", bad)
