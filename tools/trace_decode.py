"""Development aid: decodes 8 frames once with CLDN_B200_TRACE and prints per-phase tile latencies."""
import os, sys
os.environ["CLDN_B200_TRACE"] = "/tmp/cldn_trace.bin"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cloudini_b200 as cb
from cloudini_b200 import synth
F, N = int(os.environ.get("TRACE_FRAMES", "8")), 1_000_000
info = synth.info_xyzi(N)
s = torch.cuda.Stream(); torch.cuda.set_stream(s)
enc = cb.PointcloudEncoder(info, stream=s.cuda_stream); dec = cb.PointcloudDecoder(stream=s.cuda_stream)
clouds = [torch.from_numpy(synth.cloud_c2(N, seed=50 + k)[1]).cuda() for k in range(F)]
cap = cb.MaxCompressedSize(info, N, True)
blobs = [torch.empty(cap, dtype=torch.uint8, device="cuda") for _ in range(F)]
outs = [torch.zeros(N * 16, dtype=torch.uint8, device="cuda") for _ in range(F)]
eb = enc.make_device_batch([t.data_ptr() for t in clouds], [N * 16] * F, [t.data_ptr() for t in blobs], [cap] * F)
sizes = enc.encode_batch_device(eb, True, want_sizes=True)
hdr = len(enc.getHeader())
db = dec.make_device_batch([t.data_ptr() + hdr for t in blobs], [x - hdr for x in sizes], [t.data_ptr() for t in outs], [N * 16] * F)
for _ in range(3):
    dec.decode_batch_device(info, db, sync=False)
dec.sync()
tr = np.fromfile("/tmp/cldn_trace.bin", dtype=np.uint64).reshape(-1, 8)
tr = tr[tr[:, 0] > 0]
t0 = tr[:, 0].min()
rel = (tr.astype(np.int64) - np.int64(t0)) / 1000.0
print("tiles", len(tr), "kernel span us", rel.max())
names = ["start->scan", "scan->LB1done(w0)", "LB1->decoded", "decoded->sync", "sync->reduce", "reduce->LB2", "LB2->end"]
if os.environ.get("CLDN_B200_DECODE_MODE") == "seq":
    names = ["load+masks", "rank+compact", "decode run", "reduce+scan", "emit", "end sync", "-"]
for k in range(7):
    ok = (tr[:, k + 1] > 0) & (tr[:, k] > 0)
    if not ok.any():
        continue
    dur = (rel[:, k + 1] - rel[:, k])[ok]
    print(f"{names[k]:22s} mean {dur.mean():7.2f} us  p50 {np.median(dur):7.2f}  p95 {np.percentile(dur,95):7.2f}  max {dur.max():7.2f}")
last = 6 if os.environ.get("CLDN_B200_DECODE_MODE") == "seq" else 7
tot = rel[:, last] - rel[:, 0]
print("tile total mean", tot[tr[:,last]>0].mean(), "us")
# start time vs tile index
idx = np.arange(len(tr))

