#!/usr/bin/env python
"""bench.py — Mpoints/s encode+decode of 1M-point XYZI clouds (BASELINE.json configs[1] / C5 frame-sharded batch).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--frames F] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One "step" = one pass of the hot path over a batch of F distinct synthetic 1M-point XYZI clouds per GPU (1 mm
resolution, stage 1 only): one fused encode launch for the batch, then one fused decode of the blobs just produced.
The pool of F clouds (F x 16 MB >> 126 MB L2) is device resident; `value` is whole-job points/s with inputs in HBM,
`e2e` is the same metric through the host-pointer C ABI (pinned host buffers, H2D + D2H inside the timed region).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

POINTS = 1_000_000
WORKLOAD = "C2/C5: 1M-point synthetic XYZI float32x4 clouds, 1 mm, stage 1 only (compression NONE), frame-sharded batch"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--frames", type=int, default=128,
                    help="distinct 1M-point clouds per GPU per step (C5 is a batch of 10k clouds; 128 = 2 GB of input per launch. "
                         "Round 1 used 32: profiles/r2_batch_sweep.json has 32 / 64 / 100 / 128 side by side)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary measurements (tools/extras_bench.py, child process)")
    ap.add_argument("--e2e-frames", type=int, default=8, help="frames per end-to-end step (pinned host buffers)")
    ap.add_argument("--e2e-pipelines", type=int, default=2,
                    help="independent encoder/decoder handle pairs streaming side by side in the end-to-end leg (each pair = two host "
                         "threads; a second pair fills the copy-engine bubbles between the batches of the first)")
    return ap.parse_args()


class ClockSampler:
    """Samples SM clocks / throttle reasons while the timed region runs (B200_PROFILING.md's clocks line).
    NVML in-process (2 ms period: the timed region of this byte-stream workload lasts milliseconds, far below
    nvidia-smi's polling granularity); falls back to `nvidia-smi -lms` when pynvml is unavailable."""

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None
        self.nvml = None
        self.samples = []
        self.running = False

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.running = True
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        self._start_smi()

    def _poll(self):
        nv = self.nvml
        while self.running:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)
                try:
                    reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                self.samples.append((sm, reasons))
            except Exception:
                pass
            time.sleep(0.002)

    def _stop_nvml(self):
        nv = self.nvml
        self.running = False
        self.thread.join(timeout=1)
        try:
            mx = nv.nvmlDeviceGetMaxClockInfo(self.handle, nv.NVML_CLOCK_SM)
        except Exception:
            mx = None
        bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}
        reasons = sorted({k for _, r in self.samples for k, b in bits.items() if r & b})
        sm = [s for s, _ in self.samples]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": reasons, "samples": len(sm),
                "source": "nvml, 2 ms period, timed region only"}

    def _start_smi(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.nvml:
            return self._stop_nvml()
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for k, nme in enumerate(names):
                    if r[4 + k].lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


NCU_SUMMARY = "profiles/r2_kernels.json"


def measured_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed `ncu --set full` capture (profiles/r2_kernels.json, workload
    `xyzi`: tools/kernel_tour.py, one 32-frame launch), and the batch it was taken on."""
    p = os.path.join(ROOT, NCU_SUMMARY)
    try:
        w = json.load(open(p))["workloads"]["xyzi"]
        frames = int(w["event_timed_legs"][0]["frames"]) if w.get("event_timed_legs") else 32
        for k in w["kernels"]:
            if kernel in k["kernel"]:
                return float(k["dram_traffic_bytes"]), frames
    except Exception:
        pass
    return None, None


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
        except Exception:
            pass
    return 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"


def cpu_reference_numbers(oracle, info, cloud, blob, seconds, threads):
    """Times the CPU path on a bounded sample: `reps` encode + decode passes of one 1M-point cloud per thread."""
    t1, _ = oracle.time_encode(info, cloud, 1, 1)
    t2 = oracle.time_decode(blob, 1, 1)
    reps = max(1, int(seconds / max(t1 + t2, 1e-3)))
    te, _ = oracle.time_encode(info, cloud, reps, threads)
    td = oracle.time_decode(blob, reps, threads)
    pts = reps * threads * POINTS
    return {"enc_mpts": pts / te / 1e6, "dec_mpts": pts / td / 1e6, "rt_mpts": pts / (te + td) / 1e6,
            "sample": f"{reps} x {threads} encode+decode passes of one 1M-point XYZI cloud ({(te + td):.1f} s)", "seconds": te + td}


def run_extras(timeout_s=240):
    """Secondary measurements (N3 kernels, C3 / V5 sections, DDS converter step) in a CHILD process: whatever happens
    there — exception, CUDA error, crash, timeout — ends up as a string in `extras`, never in the headline numbers."""
    try:
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "extras_bench.py")], capture_output=True, text=True, timeout=timeout_s)
        lines = [ln for ln in r.stdout.strip().splitlines() if ln.startswith("{")]
        if r.returncode != 0 or not lines:
            return {"error": f"exit {r.returncode}: {(r.stderr or r.stdout)[-300:]}"}
        return json.loads(lines[-1])
    except Exception as e:  # noqa: BLE001
        return {"error": f"{type(e).__name__}: {e}"[:300]}


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation on all host threads (rank 0 only).

    Nothing of the product runs or is even loaded in this process: oracle/_ref/libcloudini_ref.so (the unmodified
    reference, compiled by oracle/build_ref.sh) builds the EncodingInfo itself (ref_bench_xyzi). One step = every host
    thread running `passes` (>= 4) encode and decode passes over its PRIVATE copy of one of 8 distinct clouds, at least
    ~2 s of wall clock, so that box-to-box differences in memory bandwidth are averaged rather than sampled."""
    if rank != 0:
        return
    import ctypes as C
    from cloudini_b200 import synth  # numpy-only cloud generator; importing it does not load the product library
    threads = os.cpu_count() or 1
    n_clouds = 8
    clouds = np.concatenate([synth.cloud_c2(POINTS, seed=1000 + k)[1] for k in range(n_clouds)])
    ref_lib = os.path.join(ROOT, "oracle", "_ref", "libcloudini_ref.so")
    kind = "reference"
    if os.path.exists(ref_lib):
        L = C.CDLL(ref_lib)
        L.ref_bench_xyzi.restype = C.c_int
        L.ref_bench_xyzi.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_float, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_size_t)]
        L.ref_last_error.restype = C.c_char_p

        def run(passes):
            te, td, nb = C.c_double(0), C.c_double(0), C.c_size_t(0)
            if L.ref_bench_xyzi(clouds.ctypes.data, POINTS * 16, n_clouds, 0.001, threads, passes, C.byref(te), C.byref(td), C.byref(nb)) != 0:
                raise SystemExit("bench.py --impl reference: " + L.ref_last_error().decode())
            return te.value, td.value
    else:  # no compiled reference on this box: the C restatement of the same path (oracle/cloudini_oracle.c), noted in `kind`
        from oracle.client import PortOracle
        oracle = PortOracle()
        kind = oracle.kind
        info = synth.info_xyzi(POINTS)
        cloud0 = clouds[:POINTS * 16]
        blob = oracle.encode(info, cloud0)

        def run(passes):
            te, _ = oracle.time_encode(info, cloud0, passes, threads)
            return te, oracle.time_decode(blob, passes, threads)

    budget = max(2.2, min(20.0, 150.0 / max(args.steps + args.warmup, 1)))   # seconds per step
    passes = 4
    te, td = run(passes)                                                      # calibration (also the first warm-up)
    for _ in range(4):                                                        # grow the sample until a step lasts >= 2 s
        if te + td >= 2.0:
            break
        passes = max(passes + 1, int(np.ceil(passes * min(budget, 2.3) / max(te + td, 1e-3))))
        te, td = run(passes)
    steps = []
    for s_ in range(args.warmup + args.steps):
        te, td = run(passes)
        if s_ >= args.warmup:
            steps.append((te, td))
    pts = passes * threads * POINTS
    te = float(np.mean([a for a, _ in steps]))
    td = float(np.mean([b for _, b in steps]))
    rt = pts / (te + td) / 1e6
    sample = f"{passes} x {threads} encode+decode passes per step, every thread on its own copy of one of {n_clouds} distinct 1M-point XYZI clouds ({te + td:.1f} s per step)"
    line = {
        "impl": "reference", "metric": "Mpoints/s encode+decode (1M-pt XYZI, 1mm res)", "value": rt, "unit": "Mpoints/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": (te + td) * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32->i32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "impl_detail": f"{kind} CPU path (built -O3 -DNDEBUG -msse4.1 like its Release default), {threads} host threads, "
                                                        "one encoder / decoder instance and private buffers per thread"},
        "cpu_baseline": {"value": rt, "unit": "Mpoints/s", "cores": threads, "kind": kind, "sample": sample,
                         "encode_mpts": pts / te / 1e6, "decode_mpts": pts / td / 1e6},
        "e2e": {"value": rt, "unit": "Mpoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import cloudini_b200 as cb
    from cloudini_b200 import synth

    if b"cusim" in cb.lib().cldn_b200_version():  # tests/cusim is a CPU emulation for kernel-logic tests, never a bench target
        raise SystemExit("bench.py: CLDN_B200_LIB points at the cusim test emulation — refusing to measure it")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — cloudini_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # host side of the end-to-end leg: this rank's threads and the pinned buffers they allocate live on the NUMA node of
    # its GPU (with 8 ranks streaming at once, remote-node staging halves the achievable PCIe rate)
    numa_cpus = cb.bind_host_thread_to_device(local_rank)
    distributed = world > 1
    if distributed:
        dist.init_process_group("nccl", device_id=dev)

    F = args.frames
    info = synth.info_xyzi(POINTS)
    stream = torch.cuda.Stream(device=dev)  # the library's kernels and the timing events share this stream
    torch.cuda.set_stream(stream)
    enc = cb.PointcloudEncoder(info, device=local_rank, stream=stream.cuda_stream)
    dec = cb.PointcloudDecoder(device=local_rank, stream=stream.cuda_stream)

    # ---- pool of F distinct clouds per rank (seeds 1000 + rank*64 + k), device resident ----
    host_clouds = [synth.cloud_c2(POINTS, seed=1000 + rank * 64 + k)[1] for k in range(F)]
    d_in = [torch.from_numpy(c).to(dev) for c in host_clouds]
    cap = cb.MaxCompressedSize(info, POINTS, True)
    d_blob = [torch.empty(cap, dtype=torch.uint8, device=dev) for _ in range(F)]
    d_out = [torch.zeros(POINTS * 16, dtype=torch.uint8, device=dev) for _ in range(F)]
    ebatch = enc.make_device_batch([t.data_ptr() for t in d_in], [POINTS * 16] * F, [t.data_ptr() for t in d_blob], [cap] * F)
    sizes = enc.encode_batch_device(ebatch, write_header=True, want_sizes=True)
    hdr = len(enc.getHeader())
    dbatch = dec.make_device_batch([t.data_ptr() + hdr for t in d_blob], [s - hdr for s in sizes], [t.data_ptr() for t in d_out], [POINTS * 16] * F)
    dec.decode_batch_device(info, dbatch, sync=True)

    # ---- parity check inside the bench (checker only): EVERY frame of rank 0's pool against the oracle ----
    parity = "unchecked"
    if rank == 0:
        try:
            from oracle.client import best_oracle
            oracle = best_oracle()
            ok = True
            for k in range(F):
                expect = oracle.encode(info, host_clouds[k])
                got = bytes(d_blob[k][:sizes[k]].cpu().numpy())
                want = np.zeros(POINTS * 16, dtype=np.uint8)
                oracle.decode(expect, want)
                ok = ok and got == expect and np.array_equal(d_out[k].cpu().numpy(), want)
            parity = f"bit-exact (all {F} frames)" if ok else "MISMATCH"
        except Exception as e:  # the oracle is a checker; its absence must not turn into a fake number
            parity = f"oracle unavailable: {e}"
        if parity == "MISMATCH":
            raise SystemExit("bench.py: GPU output differs from the oracle — refusing to report a number")
    fast_chunks, redo_chunks = dec.last_stats()

    def barrier():
        torch.cuda.synchronize()
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step(ev=None):
        if ev:
            ev[0].record()
        enc.encode_batch_device(ebatch, write_header=True)
        if ev:
            ev[1].record()
        dec.decode_batch_device(info, dbatch, sync=False)
        if ev:
            ev[2].record()

    for _ in range(args.warmup):
        device_step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = cb.kernel_launch_count()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    t_start.record()
    for s in range(args.steps):
        device_step(evs[s])
    t_end.record()
    barrier()
    launches = cb.kernel_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    enc.sync()
    dec.sync()
    elapsed_ms = t_start.elapsed_time(t_end)
    enc_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in evs]))
    dec_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in evs]))

    # ---- end-to-end through the host-pointer C ABI (pinned host memory, copies inside the timed region) ----
    e2e = None
    if not args.no_e2e:
        Fe = min(F, args.e2e_frames)
        h_in = [torch.from_numpy(host_clouds[k]).pin_memory() for k in range(Fe)]
        h_blob = [torch.empty(cap, dtype=torch.uint8).pin_memory() for _ in range(Fe)]
        h_out = [torch.zeros(POINTS * 16, dtype=torch.uint8).pin_memory() for _ in range(Fe)]
        P = max(1, args.e2e_pipelines)
        hencs = [cb.PointcloudEncoder(info, device=local_rank) for _ in range(P)]
        hdecs = [cb.PointcloudDecoder(device=local_rank) for _ in range(P)]
        henc, hdec = hencs[0], hdecs[0]

        # Two host threads, as a streaming user of the C ABI would run it: one drives the encoder handle, the other the
        # decoder handle (each handle is single-threaded, the two are independent), so frame batch i is decoded while
        # batch i+1 is being encoded and both PCIe directions are busy. Every point is still uploaded, encoded,
        # downloaded, uploaded again as a blob, decoded and downloaded inside the timed region.
        import queue, threading
        h_blob2 = [h_blob, [torch.empty(cap, dtype=torch.uint8).pin_memory() for _ in range(Fe)]]
        # pipeline 0 uses the buffers above; every further pipeline has its own blob ring and output buffers
        pipe_blobs = [h_blob2] + [[[torch.empty(cap, dtype=torch.uint8).pin_memory() for _ in range(Fe)] for _ in range(2)] for _ in range(P - 1)]
        pipe_outs = [h_out] + [[torch.zeros(POINTS * 16, dtype=torch.uint8).pin_memory() for _ in range(Fe)] for _ in range(P - 1)]
        last_w = [None]

        def run_pipeline(n_steps):
            """n_steps batches of Fe frames in total, dealt round-robin to the P handle pairs."""
            errs, ts = [], []

            def make(pi, steps_here):
                free, ready = threading.Semaphore(2), queue.Queue()
                enc_h, dec_h, blobs, outs = hencs[pi], hdecs[pi], pipe_blobs[pi], pipe_outs[pi]

                def enc_loop():
                    try:
                        cb.bind_host_thread_to_device(local_rank)
                        for s_ in range(steps_here):
                            free.acquire()
                            w_ = enc_h.encode_batch_host(h_in, blobs[s_ & 1], write_header=True)
                            ready.put((s_ & 1, w_))
                    except Exception as ex:  # noqa: BLE001 - surfaced below
                        errs.append(ex)
                        ready.put(None)

                def dec_loop():
                    try:
                        cb.bind_host_thread_to_device(local_rank)
                        for _ in range(steps_here):
                            item = ready.get()
                            if item is None:
                                return
                            b_, w_ = item
                            dec_h.decode_batch_host(info, [x[hdr:n] for x, n in zip(blobs[b_], w_)], outs)
                            if pi == 0:
                                last_w[0] = w_
                            free.release()
                    except Exception as ex:  # noqa: BLE001
                        errs.append(ex)
                        free.release()

                return [threading.Thread(target=enc_loop), threading.Thread(target=dec_loop)]

            for pi in range(P):
                steps_here = n_steps // P + (1 if pi < n_steps % P else 0)
                if steps_here:
                    ts += make(pi, steps_here)
            for t_ in ts:
                t_.start()
            for t_ in ts:
                t_.join()
            if errs:
                raise errs[0]

        run_pipeline(max(args.warmup, 3) * P)
        barrier()
        e2e_steps = max(4, min(args.steps, 12)) * P
        t0 = time.perf_counter()
        run_pipeline(e2e_steps)
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        w = last_w[0]
        # the same work as one serial round trip per step (encode call, then decode call), for reference
        t0 = time.perf_counter()
        for _ in range(3):
            ws = henc.encode_batch_host(h_in, h_blob, write_header=True)
            hdec.decode_batch_host(info, [b[hdr:n] for b, n in zip(h_blob, ws)], h_out)
        serial_s = (time.perf_counter() - t0) / 3
        if rank == 0 and parity.startswith("bit-exact"):
            assert np.array_equal(h_out[0].numpy(), d_out[0].cpu().numpy()), "host path differs from device path"
        # the same call with PAGEABLE caller buffers (what a ROS plugin hands over: std::vector), and the latency of one
        # 1M-point message through the host-pointer API (encode call, then decode call)
        p_in = [np.array(host_clouds[k], copy=True) for k in range(Fe)]
        p_blob = [np.empty(cap, dtype=np.uint8) for _ in range(Fe)]
        p_out = [np.zeros(POINTS * 16, dtype=np.uint8) for _ in range(Fe)]
        henc.encode_batch_host(p_in, p_blob, write_header=True)
        t0 = time.perf_counter()
        for _ in range(2):
            ws = henc.encode_batch_host(p_in, p_blob, write_header=True)
            hdec.decode_batch_host(info, [b[hdr:n] for b, n in zip(p_blob, ws)], p_out)
        pageable_s = (time.perf_counter() - t0) / 2
        lat_e, lat_d = [], []
        for _ in range(7):
            t0 = time.perf_counter()
            w1 = henc.encode_batch_host(h_in[:1], h_blob[:1], write_header=True)
            t1 = time.perf_counter()
            hdec.decode_batch_host(info, [h_blob[0][hdr:w1[0]]], h_out[:1])
            t2 = time.perf_counter()
            lat_e.append(t1 - t0)
            lat_d.append(t2 - t1)
        blob_bytes = int(sum(w))
        # ---- the PCIe ceiling of this box for exactly this traffic: plain pinned copies of one step's bytes (no kernels),
        #      each direction alone and both at once on two streams. The e2e figure is read against `both`. ----
        pcie = None
        try:
            s_up, s_dn = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
            up_dev = [torch.empty_like(t, device=dev) for t in h_in] + [torch.empty(int(n), dtype=torch.uint8, device=dev) for n in w]
            up_src = list(h_in) + [b[:int(n)] for b, n in zip(h_blob, w)]
            dn_dev = [torch.empty(int(n), dtype=torch.uint8, device=dev) for n in w] + [torch.empty(POINTS * 16, dtype=torch.uint8, device=dev) for _ in range(Fe)]
            dn_dst = [b[:int(n)] for b, n in zip(h_blob2[1], w)] + list(h_out)

            def copy_round(up, dn, reps=4):
                torch.cuda.synchronize()
                t0_ = time.perf_counter()
                for _ in range(reps):
                    if up:
                        with torch.cuda.stream(s_up):
                            for d_, s_ in zip(up_dev, up_src):
                                d_.copy_(s_, non_blocking=True)
                    if dn:
                        with torch.cuda.stream(s_dn):
                            for d_, s_ in zip(dn_dst, dn_dev):
                                d_.copy_(s_, non_blocking=True)
                torch.cuda.synchronize()
                return (time.perf_counter() - t0_) / reps

            copy_round(True, True, 1)
            up_b = sum(t.numel() for t in up_src)
            dn_b = sum(t.numel() for t in dn_dst)
            t_up, t_dn, t_both = copy_round(True, False), copy_round(False, True), copy_round(True, True)
            pcie = {"h2d_alone_gbs": up_b / t_up / 1e9, "d2h_alone_gbs": dn_b / t_dn / 1e9,
                    "both_gbs_each_way": 0.5 * (up_b + dn_b) / t_both / 1e9,
                    "ceiling_mpoints_s": Fe * POINTS / t_both / 1e6,
                    "note": "pinned cudaMemcpyAsync of one e2e step's bytes (clouds + blobs up, blobs + clouds down), no kernels, "
                            "two streams; ceiling = the points of one step / the time both directions need together"}
            del up_dev, dn_dev
        except Exception as ex:  # noqa: BLE001 - the probe is context for the e2e figure, never a reason to lose the line
            pcie = {"error": str(ex)}
        e2e = {"seconds": e2e_s, "pcie": pcie, "pipelines": P, "steps": e2e_steps, "frames": Fe, "h2d": Fe * POINTS * 16 + blob_bytes - Fe * hdr, "d2h": blob_bytes + Fe * POINTS * 16,
               "serial_mpts": Fe * POINTS / serial_s / 1e6, "pageable_mpts": Fe * POINTS / pageable_s / 1e6,
               "lat_enc_ms": float(np.median(lat_e)) * 1e3, "lat_dec_ms": float(np.median(lat_d)) * 1e3}

    # ---- reduce over ranks: time = max, points = sum (no data-path collective: frames are independent) ----
    from cloudini_b200 import dist as cdist
    total_points, (elapsed_ms, enc_ms_max, dec_ms_max, e2e_s_max) = cdist.aggregate(
        F * POINTS * args.steps, [elapsed_ms, enc_ms, dec_ms, e2e["seconds"] if e2e else 0.0], device=dev)
    value = total_points / (elapsed_ms * 1e-3) / 1e6

    if rank == 0:
        peak, peak_src = measured_peak()
        stage1_bytes = float(np.mean(sizes)) - hdr                      # S: stage-1 bytes incl. the u32 chunk prefixes
        algo_bytes = F * (POINTS * 16 + stage1_bytes)                    # encode: read N*point_step once + write S once
        achieved = algo_bytes / (enc_ms * 1e-3) / 1e9
        tb, tf = measured_traffic("encode_xyzi_fast_kernel")
        traffic = tb * F / tf if tb else None
        tbd, tfd = measured_traffic("decode_floatn_fast_kernel")
        dec_traffic = tbd * F / tfd if tbd else None
        dec_achieved = algo_bytes / (dec_ms * 1e-3) / 1e9
        line = {
            "metric": "Mpoints/s encode+decode (1M-pt XYZI, 1mm res)", "value": value, "unit": "Mpoints/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32->i32 (quantise) / u8 (varint stream)", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_gpu_per_step": F, "points_per_frame": POINTS, "point_step": 16,
                       "stage1_bytes_per_point": stage1_bytes / POINTS, "cache": "inputs larger than L2 (pool of %d x 16 MB per GPU)" % F,
                       "parallelism": f"frame-sharded x{world}, no data-path collective", "parity": parity,
                       "host_numa_cpus": numa_cpus,
                       "encode_mpts": world * F * POINTS / (enc_ms_max * 1e-3) / 1e6, "decode_mpts": world * F * POINTS / (dec_ms_max * 1e-3) / 1e6,
                       "encode_ms_per_step": enc_ms, "decode_ms_per_step": dec_ms},
            # dominant kernel of the step = the one with the larger share of the timed region (the FloatN decode);
            # the encode kernel (the one SURVEY 8(d)'s 60 % target is stated on) is reported next to it
            "roofline": {"bound": "hbm", "kernel": "decode_floatn_fast_kernel<4> (terminator ranking + varint reader + un-zigzag + per-field prefix sums + dequantise)",
                         "achieved": dec_achieved, "peak": peak, "unit": "GB/s", "frac": dec_achieved / peak,
                         "traffic": dec_traffic, "share_of_step": dec_ms / (enc_ms + dec_ms),
                         "traffic_source": NCU_SUMMARY + " (ncu --set full, dram__bytes_read+write of one 32-frame launch, scaled to this batch)",
                         "chunks_fast_reader": fast_chunks, "chunks_careful_reader": redo_chunks,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": algo_bytes, "launch_ms": dec_ms,
                         "encode": {"kernel": "encode_xyzi_fast_kernel (quantise + delta + zigzag varint + pack, persistent CTAs)", "achieved": achieved,
                                    "frac": achieved / peak, "traffic": traffic, "launch_ms": enc_ms, "share_of_step": enc_ms / (enc_ms + dec_ms),
                                    "algorithmic_bytes_per_launch": algo_bytes}},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if e2e:
            e2e_pts = world * e2e["frames"] * POINTS * e2e["steps"]
            line["e2e"] = {"value": e2e_pts / e2e_s_max / 1e6, "unit": "Mpoints/s", "h2d_bytes_per_step": int(e2e["h2d"]),
                           "d2h_bytes_per_step": int(e2e["d2h"]), "frames_per_step": e2e["frames"],
                           "api": "cldn_b200_encode_batch + cldn_b200_decode_batch, CLDN_MEM_HOST, pinned host buffers; %d independent "
                                  "encoder/decoder handle pair(s), each driven by two host threads (batch i decodes while batch i+1 "
                                  "encodes); every point is uploaded, encoded, downloaded, uploaded as a blob, decoded and downloaded" % e2e["pipelines"],
                           "pipelines": e2e["pipelines"],
                           "serial_roundtrip_mpoints_s": world * e2e["serial_mpts"],
                           "pageable_buffers_mpoints_s": world * e2e["pageable_mpts"],
                           "pcie_ceiling": e2e["pcie"],
                           "frac_of_pcie_ceiling": (e2e_pts / e2e_s_max / 1e6 / (world * e2e["pcie"]["ceiling_mpoints_s"])
                                                    if e2e["pcie"] and "ceiling_mpoints_s" in e2e["pcie"] else None),
                           "one_message_latency_ms": {"encode": e2e["lat_enc_ms"], "decode": e2e["lat_dec_ms"],
                                                      "note": "one 1M-point frame, pinned host buffers, host-pointer API, median of 7"}}
        # ---- CPU baseline on this box's host cores (rank 0, N=1 only): bounded sample ----
        if world == 1:
            try:
                from oracle.client import best_oracle
                oracle = best_oracle()
                blob0 = oracle.encode(info, host_clouds[0])
                r = cpu_reference_numbers(oracle, info, host_clouds[0], blob0, args.cpu_seconds, 1)
                line["cpu_baseline"] = {"value": r["rt_mpts"], "unit": "Mpoints/s", "cores": 1, "kind": oracle.kind, "sample": r["sample"],
                                        "encode_mpts": r["enc_mpts"], "decode_mpts": r["dec_mpts"], "host_cores_available": os.cpu_count()}
            except Exception as e:
                line["cpu_baseline"] = {"value": None, "unit": "Mpoints/s", "cores": 0, "kind": "unavailable", "sample": str(e)}
        if world == 1 and not args.no_extras:
            line["extras"] = run_extras()
        print(json.dumps(line), flush=True)
    if distributed:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
