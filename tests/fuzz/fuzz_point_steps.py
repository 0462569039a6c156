import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT)
import cloudini_b200 as cb
from cloudini_b200 import FieldType as F
from oracle.client import RefOracle
ref = RefOracle()
rng = np.random.default_rng(int(sys.argv[1]))
types = [F.INT8,F.UINT8,F.INT16,F.UINT16,F.INT32,F.UINT32,F.FLOAT32,F.FLOAT64,F.INT64,F.UINT64]
bad = 0; tot = 0
for step in (1,2,3,4,5,7,12,13,31,64,255,256,257,1000,4096,70000):
    for trial in range(4):
        n = int(rng.choice([1, 7, 300, 2049, 33000 if step < 300 else 500]))
        # random non-overlapping fields inside the step
        fields, pos = [], 0
        while pos < step and len(fields) < 6:
            t = types[rng.integers(0, len(types))]
            sz = cb.SizeOf(t)
            pos += int(rng.integers(0, 3))
            if pos + sz > step: break
            res = None
            if t in (F.FLOAT32, F.FLOAT64) and rng.random() < 0.7: res = float(rng.choice([0.001, 0.01, 0.5]))
            fields.append(cb.PointField(f"f{len(fields)}", pos, t, res)); pos += sz
        if not fields: 
            continue
        info = cb.EncodingInfo(fields=fields, width=n, height=1, point_step=step, compression_opt=cb.CompressionOption.NONE, use_threads=False,
                               encoding_opt=cb.EncodingOptions(int(rng.integers(0,3))), version=int(rng.choice([5,4,3])))
        cloud = rng.integers(0, 256, n*step, dtype=np.uint8)
        # make float fields sane-ish
        v = cloud.reshape(n, step)
        for f in fields:
            if f.type == F.FLOAT32: v[:, f.offset:f.offset+4] = np.cumsum(rng.normal(0,0.01,n)).astype(np.float32).view(np.uint8).reshape(n,4)
            if f.type == F.FLOAT64: v[:, f.offset:f.offset+8] = (1e6+np.cumsum(rng.normal(0,0.01,n))).astype(np.float64).view(np.uint8).reshape(n,8)
        try: want_blob = ref.encode(info, cloud)
        except RuntimeError as e: continue
        tot += 1
        try:
            blob = cb.PointcloudEncoder(info).encode(cloud)
            ok = blob == want_blob
            dinfo, hdr = cb.DecodeHeader(blob)
            want = np.full(n*step, 0x21, np.uint8); ref.decode(want_blob, want)
            got = np.full(n*step, 0x21, np.uint8); cb.PointcloudDecoder().decode(dinfo, blob[hdr:], got)
            ok = ok and np.array_equal(got, want)
        except RuntimeError as e:
            ok = False; print("ERR", str(e)[:100])
        if not ok:
            bad += 1; print("MISMATCH step", step, "n", n, [(int(f.type), f.offset, f.resolution) for f in fields], info.encoding_opt, info.version)
print("cases", tot, "mismatches", bad)
