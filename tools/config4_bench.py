"""BASELINE configs[3] as written: Velodyne-style 64-ring rolling clouds (130 048 points per frame), a 1000-frame batch,
encode + LZ4 on ONE B200 — stage 1 and the per-chunk LZ4 blocks both on the device (cldn_lz4.cu; the image has no nvCOMP).
Prints one JSON line. The reference decodes a sample of the blobs (interoperability), nothing else of it is involved.
   python tools/config4_bench.py [--frames 1000] [--reps 5]"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import cloudini_b200 as cb  # noqa: E402
from cloudini_b200 import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=1000)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--distinct", type=int, default=250, help="distinct rolling frames generated on the host (the batch cycles through them)")
    args = ap.parse_args()
    F = args.frames
    info, _ = synth.cloud_c4_frame(0)
    n = info.width
    info.compression_opt = cb.CompressionOption.LZ4
    t0 = time.time()
    host = [synth.cloud_c4_frame(k)[1] for k in range(min(args.distinct, F))]
    gen_s = time.time() - t0
    s = torch.cuda.Stream()
    torch.cuda.set_stream(s)
    enc, dec = cb.PointcloudEncoder(info, stream=s.cuda_stream), cb.PointcloudDecoder(stream=s.cuda_stream)
    d_src = [torch.from_numpy(c).cuda() for c in host]
    d_in = [d_src[k % len(d_src)] for k in range(F)]
    cap = cb.MaxCompressedSize(info, n, True)
    d_blob = torch.empty((F, (cap + 255) // 256 * 256), dtype=torch.uint8, device="cuda")
    d_out = torch.zeros((F, n * 16), dtype=torch.uint8, device="cuda")
    eb = enc.make_device_batch([t.data_ptr() for t in d_in], [n * 16] * F, [d_blob[k].data_ptr() for k in range(F)], [cap] * F)
    sizes = enc.encode_batch_device(eb, True, want_sizes=True)
    hdr = len(enc.getHeader())
    db = dec.make_device_batch([d_blob[k].data_ptr() + hdr for k in range(F)], [x - hdr for x in sizes], [d_out[k].data_ptr() for k in range(F)], [n * 16] * F)
    dec.decode_batch_device(info, db, sync=True)
    # interoperability sample: the stock reference decodes our blobs to what we decode
    interop = "unchecked"
    try:
        from oracle.client import RefOracle
        ref = RefOracle()
        ok = True
        for k in (0, F // 2, F - 1):
            blob = bytes(d_blob[k][:sizes[k]].cpu().numpy())
            want = np.zeros(n * 16, dtype=np.uint8)
            ref.decode(blob, want)
            ok = ok and np.array_equal(want, d_out[k].cpu().numpy())
        interop = "reference decodes the device LZ4 blobs to the same points" if ok else "MISMATCH"
    except Exception as e:  # noqa: BLE001
        interop = f"reference unavailable: {e}"
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    te = td = 0.0
    for _ in range(args.reps):
        ev[0].record()
        enc.encode_batch_device(eb, True)
        ev[1].record()
        dec.decode_batch_device(info, db, sync=False)   # (the LZ4 decode reads the frames' stage-1 sizes back: one sync inside)
        ev[2].record()
        torch.cuda.synchronize()
        te += ev[0].elapsed_time(ev[1])
        td += ev[1].elapsed_time(ev[2])
    te, td = te / args.reps, td / args.reps
    # stage 1 alone on the same batch, for the split
    info1, _ = synth.cloud_c4_frame(0)
    enc1 = cb.PointcloudEncoder(info1, stream=s.cuda_stream)
    cap1 = cb.MaxCompressedSize(info1, n, True)
    eb1 = enc1.make_device_batch([t.data_ptr() for t in d_in], [n * 16] * F, [d_blob[k].data_ptr() for k in range(F)], [cap1] * F)
    sizes1 = enc1.encode_batch_device(eb1, True, want_sizes=True)
    t1 = 0.0
    for _ in range(args.reps):
        ev[0].record()
        enc1.encode_batch_device(eb1, True)
        ev[1].record()
        torch.cuda.synchronize()
        t1 += ev[0].elapsed_time(ev[1])
    t1 /= args.reps
    out = {"config": "C4: 64-ring rolling clouds, %d frames x %d points XYZI float32x4, 1 mm, encode + LZ4 on the device" % (F, n),
           "frames": F, "points_per_frame": n, "encode_lz4_ms": te, "decode_lz4_ms": td, "encode_stage1_only_ms": t1,
           "encode_lz4_mpts": F * n / te / 1e3, "decode_lz4_mpts": F * n / td / 1e3, "encode_stage1_only_mpts": F * n / t1 / 1e3,
           "stage1_bytes_per_point": (float(np.mean(sizes1)) - hdr) / n, "lz4_bytes_per_point": (float(np.mean(sizes)) - hdr) / n,
           "interop": interop, "host_generation_s": gen_s, "distinct_frames": len(host)}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
