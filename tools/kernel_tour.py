"""One pass over every shipped kernel, for `ncu --set full --profile-from-start off` (tools/gpu_ncu_all.sh).

Each workload is built and run once untimed (lazy initialisation, scratch allocation), then once between
cudaProfilerStart/Stop so that the capture holds exactly one launch of every kernel of the workload. Next to the capture
this script writes tour.json: per workload the algorithmic bytes (input + stage-1 bytes of the batch) and the
CUDA-event time of the encode / decode legs outside the profiler, i.e. the numbers the per-kernel roofline fractions in
profiles/r2_kernels.json are computed from.

  python tools/kernel_tour.py [--out tour.json] [--only xyzi,c3,...]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import cloudini_b200 as cb  # noqa: E402
from cloudini_b200 import ros, synth  # noqa: E402
from cloudini_b200 import EncodingInfo, EncodingOptions, CompressionOption, FieldType, PointField  # noqa: E402


def hbm_peak():
    try:
        return float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


class Leg:
    def __init__(self, name, info, clouds, lz4=False):
        self.name, self.info, self.F = name, info, len(clouds)
        self.n, self.step = info.width * info.height, info.point_step
        self.stream = torch.cuda.Stream()
        torch.cuda.set_stream(self.stream)
        self.enc = cb.PointcloudEncoder(info, stream=self.stream.cuda_stream)
        self.dec = cb.PointcloudDecoder(stream=self.stream.cuda_stream)
        n, step, F = self.n, self.step, self.F
        self.d_in = [torch.from_numpy(c).cuda() for c in clouds]
        cap = cb.MaxCompressedSize(info, n, True)
        self.d_blob = [torch.empty(cap, dtype=torch.uint8, device="cuda") for _ in range(F)]
        self.d_out = [torch.zeros(n * step, dtype=torch.uint8, device="cuda") for _ in range(F)]
        self.eb = self.enc.make_device_batch([t.data_ptr() for t in self.d_in], [n * step] * F, [t.data_ptr() for t in self.d_blob], [cap] * F)
        self.sizes = self.enc.encode_batch_device(self.eb, True, want_sizes=True)
        hdr = len(self.enc.getHeader())
        self.db = self.dec.make_device_batch([t.data_ptr() + hdr for t in self.d_blob], [x - hdr for x in self.sizes],
                                             [t.data_ptr() for t in self.d_out], [n * step] * F)
        self.dec.decode_batch_device(info, self.db, sync=True)
        self.stage1 = float(np.sum(self.sizes)) - hdr * F

    def encode(self):
        self.enc.encode_batch_device(self.eb, True)

    def decode(self):
        self.dec.decode_batch_device(self.info, self.db, sync=False)

    def timed(self, reps=10):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        te = td = 0.0
        for _ in range(3):
            self.encode(); self.decode()
        for _ in range(reps):
            ev[0].record(); self.encode(); ev[1].record(); self.decode(); ev[2].record()
            torch.cuda.synchronize()
            te += ev[0].elapsed_time(ev[1]); td += ev[1].elapsed_time(ev[2])
        te, td = te / reps, td / reps
        algo = self.F * self.n * self.step + self.stage1
        return {"workload": self.name, "frames": self.F, "points_per_frame": self.n, "point_step": self.step,
                "stage1_bytes_per_point": self.stage1 / (self.F * self.n), "algorithmic_bytes": algo,
                "encode_ms": te, "decode_ms": td, "encode_gbs": algo / te / 1e6, "decode_gbs": algo / td / 1e6,
                "encode_frac_of_hbm": algo / te / 1e6 / hbm_peak(), "decode_frac_of_hbm": algo / td / 1e6 / hbm_peak(),
                "encode_mpts": self.F * self.n / te / 1e3, "decode_mpts": self.F * self.n / td / 1e3}

    def profiled(self):
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        self.encode(); self.decode()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()


def lossless_info(n, version=5):
    """XYZ lossless (XOR / Gorilla coders of the wire version) + a FLOAT64 timestamp + u8 label: the generic kernels."""
    fields = [PointField("x", 0, FieldType.FLOAT32, None), PointField("y", 4, FieldType.FLOAT32, None),
              PointField("z", 8, FieldType.FLOAT32, None), PointField("t", 12, FieldType.FLOAT64, None),
              PointField("label", 20, FieldType.UINT8, None)]
    return EncodingInfo(fields=fields, width=n, height=1, point_step=21, encoding_opt=EncodingOptions.LOSSLESS,
                        compression_opt=CompressionOption.NONE, use_threads=False, version=version)


def lossless_cloud(n, seed):
    rng = np.random.default_rng(seed)
    buf = np.zeros((n, 21), dtype=np.uint8)
    xyz = np.cumsum(rng.normal(0, 0.01, size=(n, 3)), axis=0).astype(np.float32)
    buf[:, 0:12] = xyz.view(np.uint8).reshape(n, 12)
    t = (1.7e9 + np.arange(n) * 1e-5).astype(np.float64)
    buf[:, 12:20] = t.view(np.uint8).reshape(n, 8)
    buf[:, 20] = rng.integers(0, 4, size=n).astype(np.uint8)
    return buf.reshape(-1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/tour.json")
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    only = set(x for x in a.only.split(",") if x)
    res = {"hbm_peak_gbs": hbm_peak(), "workloads": []}

    def want(k):
        return not only or k in only

    if want("xyzi"):
        info, _ = synth.cloud_c2(1_000_000)
        leg = Leg("xyzi 32x1M step16 lossy 1mm (headline)", info, [synth.cloud_c2(1_000_000, seed=1000 + k)[1] for k in range(32)])
        res["workloads"].append(leg.timed()); leg.profiled(); del leg
    if want("c3"):
        info, _ = synth.cloud_c3(1_000_000)
        leg = Leg("c3 32x1M xyz+rgba+ring step32 (V5 sections)", info, [synth.cloud_c3(1_000_000, seed=3 + k)[1] for k in range(32)])
        res["workloads"].append(leg.timed()); leg.profiled(); del leg
    if want("c4"):
        info, _ = synth.cloud_c4_mixed_frame(0)
        leg = Leg("c4 256x130048 velodyne xyzirt step22", info, [synth.cloud_c4_mixed_frame(k)[1] for k in range(256)])
        res["workloads"].append(leg.timed()); leg.profiled(); del leg
    if want("lossless"):
        for ver in (5, 3):
            n = 500_000
            info = lossless_info(n, ver)
            leg = Leg(f"lossless 8x500k xyz f32 + t f64 + label u8 step21, wire v{ver} (generic kernels)", info,
                      [lossless_cloud(n, 50 + k) for k in range(8)])
            res["workloads"].append(leg.timed(reps=5)); leg.profiled(); del leg
    if want("lz4"):
        info, _ = synth.cloud_c4_frame(0)
        info.compression_opt = CompressionOption.LZ4
        F = 64
        clouds = [synth.cloud_c4_frame(k)[1] for k in range(F)]
        n, step = info.width, info.point_step
        s = torch.cuda.Stream(); torch.cuda.set_stream(s)
        enc = cb.PointcloudEncoder(info, stream=s.cuda_stream); dec = cb.PointcloudDecoder(stream=s.cuda_stream)
        d_in = [torch.from_numpy(c).cuda() for c in clouds]
        cap = cb.MaxCompressedSize(info, n, True)
        d_blob = [torch.empty(cap, dtype=torch.uint8, device="cuda") for _ in range(F)]
        d_out = [torch.zeros(n * step, dtype=torch.uint8, device="cuda") for _ in range(F)]
        eb = enc.make_device_batch([t.data_ptr() for t in d_in], [n * step] * F, [t.data_ptr() for t in d_blob], [cap] * F)
        sizes = enc.encode_batch_device(eb, True, want_sizes=True)
        hdr = len(enc.getHeader())
        db = dec.make_device_batch([t.data_ptr() + hdr for t in d_blob], [x - hdr for x in sizes], [t.data_ptr() for t in d_out], [n * step] * F)
        dec.decode_batch_device(info, db, sync=True)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        te = td = 0.0
        for _ in range(5):
            ev[0].record(); enc.encode_batch_device(eb, True); ev[1].record(); dec.decode_batch_device(info, db, sync=False); ev[2].record()
            torch.cuda.synchronize(); te += ev[0].elapsed_time(ev[1]); td += ev[1].elapsed_time(ev[2])
        te /= 5; td /= 5
        algo = F * n * step + float(np.sum(sizes))
        res["workloads"].append({"workload": "c4 64x130048 xyzi + LZ4 stage 2 on the device", "frames": F, "points_per_frame": n,
                                 "blob_bytes_per_point": float(np.sum(sizes)) / (F * n), "algorithmic_bytes": algo, "encode_ms": te,
                                 "decode_ms": td, "encode_frac_of_hbm": algo / te / 1e6 / hbm_peak(), "decode_frac_of_hbm": algo / td / 1e6 / hbm_peak()})
        torch.cuda.synchronize(); torch.cuda.profiler.start()
        enc.encode_batch_device(eb, True); dec.decode_batch_device(info, db, sync=False)
        torch.cuda.synchronize(); torch.cuda.profiler.stop()
    if want("viz"):
        n = 1_000_000
        info, cloud = synth.cloud_viz(n, seed=77)
        d_in = torch.from_numpy(cloud).cuda()
        d_out = torch.zeros(cloud.size, dtype=torch.uint8, device="cuda")
        pp = ros.VizPreprocessor()
        _, kept, _ = pp.run_device(info, d_in.data_ptr(), cloud.size, d_out.data_ptr(), cloud.size)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(10):
            pp.run_device(info, d_in.data_ptr(), cloud.size, d_out.data_ptr(), cloud.size)
        ev[1].record(); torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / 10
        algo = n * 16 + kept * 16
        res["workloads"].append({"workload": "viz preprocessing 1M xyzi", "points_per_frame": n, "kept": int(kept), "algorithmic_bytes": algo,
                                 "ms_per_call": ms, "frac_of_hbm": algo / ms / 1e6 / hbm_peak()})
        torch.cuda.synchronize(); torch.cuda.profiler.start()
        pp.run_device(info, d_in.data_ptr(), cloud.size, d_out.data_ptr(), cloud.size)
        torch.cuda.synchronize(); torch.cuda.profiler.stop()
    os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
    json.dump(res, open(a.out, "w"), indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
