"""V5 adaptive integer sections: mode choice and section bytes against the reference for structured random integers
(constant, few / ~2048 / many distinct values, runs, piecewise linear, noise, extremes) of every integer type, at sizes
around the probe window (4096) and the chunk size. Usage: python tests/fuzz/fuzz_v5_int_fields.py <seed> <trials>"""
import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import cloudini_b200 as cb
from oracle.client import RefOracle
from make_golden import int_field_cloud
ref = RefOracle()
rng = np.random.default_rng(int(sys.argv[1]))
types = [(cb.FieldType.INT16, np.int16), (cb.FieldType.UINT16, np.uint16), (cb.FieldType.INT32, np.int32), (cb.FieldType.UINT32, np.uint32),
         (cb.FieldType.INT64, np.int64), (cb.FieldType.UINT64, np.uint64)]
bad = 0
for trial in range(int(sys.argv[2])):
    ft, dt = types[int(rng.integers(0, len(types)))]
    n = int(rng.choice([1, 100, 4095, 4096, 4097, 10_000, 32768, 32769, 40_000]))
    ii = np.iinfo(dt)
    shape = int(rng.integers(0, 9))
    def rnd(k, lo=ii.min, hi=ii.max):
        return rng.integers(lo, hi, k, dtype=np.int64 if dt != np.uint64 else np.uint64, endpoint=True).astype(dt)
    if shape == 0: v = np.full(n, rnd(1)[0], dtype=dt)
    elif shape == 1: v = rnd(int(rng.integers(2, 9)))[rng.integers(0, 8, n) % int(rng.integers(2, 9))] if False else rnd(8)[rng.integers(0, 8, n)]
    elif shape == 2: v = rnd(int(rng.choice([2040, 2048, 2049, 2100])))[rng.integers(0, 2040, n)]
    elif shape == 3: v = rnd(n)
    elif shape == 4: v = rnd(n // 7 + 1)[np.arange(n) // 7]
    elif shape == 5: v = np.cumsum(rng.integers(-5, 6, n // 9 + 1)[np.arange(n) // 9]).astype(np.int64).astype(dt)
    elif shape == 6: v = (np.arange(n) * int(rng.integers(-1000, 1000))).astype(np.int64).astype(dt)
    elif shape == 7: v = rng.choice(np.array([ii.min, ii.max, 0, 1], dtype=dt), n)
    else:            # the first 4096 look like one thing, the rest like another: the committed mode must stick (v5_codec.cpp:934-949)
        v = np.concatenate([np.full(min(n, 4096), 7, dtype=dt), rnd(max(0, n - 4096))])
    info, cloud = int_field_cloud(v, ft)
    want = ref.encode(info, cloud)
    got = cb.PointcloudEncoder(info).encode(cloud)
    ok = got == want
    if ok:  # the decoders must agree too (a delta of exactly INT64_MIN encodes to a byte the reference's own decoder rejects)
        dinfo, hdr = cb.DecodeHeader(got)
        w, rok = np.zeros(cloud.size, dtype=np.uint8), True
        try: ref.decode(want, w)
        except RuntimeError: rok = False
        out, ook = np.zeros(cloud.size, dtype=np.uint8), True
        try: cb.PointcloudDecoder().decode(dinfo, got[hdr:], out)
        except RuntimeError: ook = False
        ok = rok == ook and (not rok or np.array_equal(out, w))
    if not ok:
        bad += 1
        print("MISMATCH", ft, "n", n, "shape", shape, len(got), len(want))
print("trials", trial + 1, "mismatches", bad)
