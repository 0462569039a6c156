"""Device-resident encode / decode timing of the BASELINE.json configs C1..C4 (stage 1 only), with an oracle parity
check of frame 0. Prints one JSON line per config. Development / reporting aid (bench.py stays the contract)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cloudini_b200 as cb
from cloudini_b200 import synth
from oracle.client import best_oracle


def run(name, info, clouds, reps=10):
    s = torch.cuda.Stream(); torch.cuda.set_stream(s)
    enc = cb.PointcloudEncoder(info, stream=s.cuda_stream); dec = cb.PointcloudDecoder(stream=s.cuda_stream)
    n, step, F = info.width, info.point_step, len(clouds)
    d_in = [torch.from_numpy(c).cuda() for c in clouds]
    cap = cb.MaxCompressedSize(info, n, True)
    d_blob = [torch.empty(cap, dtype=torch.uint8, device="cuda") for _ in range(F)]
    d_out = [torch.full((n * step,), 0x5A, dtype=torch.uint8, device="cuda") for _ in range(F)]
    eb = enc.make_device_batch([t.data_ptr() for t in d_in], [n * step] * F, [t.data_ptr() for t in d_blob], [cap] * F)
    sizes = enc.encode_batch_device(eb, True, want_sizes=True)
    hdr = len(enc.getHeader())
    db = dec.make_device_batch([t.data_ptr() + hdr for t in d_blob], [x - hdr for x in sizes], [t.data_ptr() for t in d_out], [n * step] * F)
    dec.decode_batch_device(info, db, sync=True)
    oracle = best_oracle()
    expect = oracle.encode(info, clouds[0])
    want = np.full(n * step, 0x5A, dtype=np.uint8); oracle.decode(expect, want)
    ok = bytes(d_blob[0][:sizes[0]].cpu().numpy()) == expect and np.array_equal(d_out[0].cpu().numpy(), want)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    for _ in range(3):
        enc.encode_batch_device(eb, True); dec.decode_batch_device(info, db, sync=False)
    te = td = 0.0
    for _ in range(reps):
        ev[0].record(); enc.encode_batch_device(eb, True); ev[1].record(); dec.decode_batch_device(info, db, sync=False); ev[2].record()
        torch.cuda.synchronize(); te += ev[0].elapsed_time(ev[1]); td += ev[1].elapsed_time(ev[2])
    te /= reps; td /= reps
    S = float(np.mean(sizes)) - hdr
    algo = F * (n * step + S)
    print(json.dumps({"config": name, "frames": F, "points": n, "point_step": step, "stage1_B_per_pt": S / n, "parity": "bit-exact" if ok else "MISMATCH",
                      "encode_ms": te, "decode_ms": td, "encode_Mpts": F * n / te / 1e3, "decode_Mpts": F * n / td / 1e3,
                      "encode_GBs": algo / te / 1e6, "decode_GBs": algo / td / 1e6}), flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["c1", "c2", "c3", "c4"]
    if "c1" in which:
        info, _ = synth.cloud_c1(10_000)
        run("C1 10k XYZ step12 x256 frames", info, [synth.cloud_c1(10_000, seed=1 + k)[1] for k in range(256)])
    if "c2" in which:
        info, _ = synth.cloud_c2(1_000_000)
        run("C2 1M XYZI step16 x32 frames", info, [synth.cloud_c2(1_000_000, seed=1000 + k)[1] for k in range(32)])
        run("C2 1M XYZI step16 x1 frame (latency)", info, [synth.cloud_c2(1_000_000, seed=1000)[1]], reps=30)
    if "c3" in which:
        info, _ = synth.cloud_c3(1_000_000)
        run("C3 1M XYZ+rgba u32+ring u16 step32 (V5) x16 frames", info, [synth.cloud_c3(1_000_000, seed=3 + k)[1] for k in range(16)])
    if "c4" in which:
        info, _ = synth.cloud_c4_frame(0)
        run("C4 130048-pt Velodyne-style XYZI frames x256 (stage 1 only; nvCOMP absent)", info, [synth.cloud_c4_frame(k)[1] for k in range(256)])
