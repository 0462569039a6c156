"""Seeded synthetic PointCloud2 workloads (BASELINE.json configs C1..C5).

Host-side numpy only; the same bytes are handed to the GPU path and to the oracle.
"""
from __future__ import annotations

import numpy as np

from . import CompressionOption, EncodingInfo, EncodingOptions, FieldType, PointField


def _lidar_xyz(n: int, rng: np.random.Generator, rings: int = 64, az_step: float = 2 * np.pi / 2048):
    i = np.arange(n, dtype=np.int64)
    ring = (i % rings).astype(np.float64)
    az = (i // rings).astype(np.float64) * az_step
    el = np.deg2rad(-24.8 + ring * (26.8 / max(rings - 1, 1)))
    r = 20.0 + 10.0 * np.sin(3.0 * az) + 2.0 * np.cos(7.0 * az + ring * 0.05) + rng.normal(0.0, 0.01, n)
    x = r * np.cos(el) * np.cos(az)
    y = r * np.cos(el) * np.sin(az)
    z = r * np.sin(el)
    return x.astype(np.float32), y.astype(np.float32), z.astype(np.float32)


def info_xyz(n: int, resolution: float = 0.001, version: int = 5) -> EncodingInfo:
    return EncodingInfo(
        fields=[PointField("x", 0, FieldType.FLOAT32, resolution), PointField("y", 4, FieldType.FLOAT32, resolution),
                PointField("z", 8, FieldType.FLOAT32, resolution)],
        width=n, height=1, point_step=12, encoding_opt=EncodingOptions.LOSSY, compression_opt=CompressionOption.NONE,
        use_threads=False, version=version)


def info_xyzi(n: int, resolution: float = 0.001, version: int = 5) -> EncodingInfo:
    info = info_xyz(n, resolution, version)
    info.fields.append(PointField("intensity", 12, FieldType.FLOAT32, resolution))
    info.point_step = 16
    return info


def cloud_c1(n: int = 10_000, seed: int = 1, adversarial: bool = False):
    """C1: XYZ float32, step 12. adversarial=True sprinkles NaN, +-inf, huge magnitudes and exact .5 ties."""
    rng = np.random.default_rng(seed)
    x, y, z = _lidar_xyz(n, rng)
    pts = np.stack([x, y, z], axis=1)
    if adversarial and n > 0:
        k = max(n // 50, 1)
        flat = pts.reshape(-1)
        idx = rng.choice(flat.size, size=min(4 * k, flat.size), replace=False)
        q = max(len(idx) // 4, 1)
        flat[idx[:q]] = np.nan
        flat[idx[q:2 * q]] = rng.choice(np.array([np.inf, -np.inf, 3e9, -3e9, 2.2e6, -2.2e6], dtype=np.float32), size=len(idx[q:2 * q]))
        flat[idx[2 * q:3 * q]] = (rng.integers(-5000, 5000, size=len(idx[2 * q:3 * q])).astype(np.float32) + 0.5) * np.float32(0.001)
        flat[idx[3 * q:]] = rng.normal(0, 1e-4, size=len(idx[3 * q:])).astype(np.float32)
    return info_xyz(n), np.ascontiguousarray(pts, dtype=np.float32).view(np.uint8).reshape(-1)


def cloud_c2(n: int = 1_000_000, seed: int = 2):
    """C2: XYZI float32x4, step 16, intensity = integers 0..255 as float."""
    rng = np.random.default_rng(seed)
    x, y, z = _lidar_xyz(n, rng)
    inten = rng.integers(0, 256, size=n).astype(np.float32)
    pts = np.stack([x, y, z, inten], axis=1)
    return info_xyzi(n), np.ascontiguousarray(pts, dtype=np.float32).view(np.uint8).reshape(-1)


def cloud_c3(n: int = 1_000_000, seed: int = 3, version: int = 5):
    """C3: XYZ f32 + rgba u32 @16 + ring u16 @20, ROS-style padded step 32, padding filled with 0xCD."""
    rng = np.random.default_rng(seed)
    x, y, z = _lidar_xyz(n, rng)
    buf = np.full((n, 32), 0xCD, dtype=np.uint8)
    buf[:, 0:4] = x.view(np.uint8).reshape(n, 4)
    buf[:, 4:8] = y.view(np.uint8).reshape(n, 4)
    buf[:, 8:12] = z.view(np.uint8).reshape(n, 4)
    rgba = (np.uint32(0xFF000000) | (rng.integers(0, 8, size=n).astype(np.uint32) * np.uint32(0x101010))).astype(np.uint32)
    ring = (np.arange(n) % 64).astype(np.uint16)
    buf[:, 16:20] = rgba.view(np.uint8).reshape(n, 4)
    buf[:, 20:22] = ring.view(np.uint8).reshape(n, 2)
    info = EncodingInfo(
        fields=[PointField("x", 0, FieldType.FLOAT32, 0.001), PointField("y", 4, FieldType.FLOAT32, 0.001),
                PointField("z", 8, FieldType.FLOAT32, 0.001), PointField("rgba", 16, FieldType.UINT32, None),
                PointField("ring", 20, FieldType.UINT16, None)],
        width=n, height=1, point_step=32, encoding_opt=EncodingOptions.LOSSY, compression_opt=CompressionOption.NONE,
        use_threads=False, version=version)
    return info, buf.reshape(-1)


def cloud_c4_frame(frame: int, seed: int = 4, rings: int = 64, az: int = 2032):
    """C4: one Velodyne-style rolling frame (64 rings x 2032 azimuths = 130048 points), XYZI float32, step 16."""
    n = rings * az
    rng = np.random.default_rng(seed * 100003 + frame)
    i = np.arange(n, dtype=np.int64)
    ring = (i % rings).astype(np.float64)
    a = (i // rings).astype(np.float64) * (2 * np.pi / az) + 0.002 * frame
    el = np.deg2rad(-24.8 + ring * (26.8 / (rings - 1)))
    r = 15.0 + 8.0 * np.sin(2.0 * a + 0.01 * frame) + 3.0 * np.cos(5.0 * a) + rng.normal(0.0, 0.008, n)
    x = (r * np.cos(el) * np.cos(a) + 0.05 * frame).astype(np.float32)
    y = (r * np.cos(el) * np.sin(a)).astype(np.float32)
    z = (r * np.sin(el)).astype(np.float32)
    inten = rng.integers(0, 256, size=n).astype(np.float32)
    pts = np.stack([x, y, z, inten], axis=1)
    return info_xyzi(n), np.ascontiguousarray(pts, dtype=np.float32).view(np.uint8).reshape(-1)


def cloud_c4_mixed_frame(frame: int, seed: int = 4, rings: int = 64, az: int = 2032):
    """C4 with the sensor's full layout: Velodyne PointXYZIRT-like, x y z intensity f32 + ring u16 @16 +
    time f32 @18, point_step 22 (unaligned on purpose, as ROS drivers pack it). Planner: FloatN(4) + V5 section (ring)
    + scalar lossy float (time, resolution 1e-5 s) in the regular stream."""
    info16, xyzi = cloud_c4_frame(frame, seed, rings, az)
    n = info16.width
    buf = np.zeros((n, 22), dtype=np.uint8)
    buf[:, :16] = xyzi.reshape(n, 16)
    i = np.arange(n, dtype=np.int64)
    buf[:, 16:18] = (i % rings).astype(np.uint16).view(np.uint8).reshape(n, 2)
    t = ((i // rings).astype(np.float64) * (0.1 / az)).astype(np.float32)  # one revolution = 100 ms
    buf[:, 18:22] = t.view(np.uint8).reshape(n, 4)
    F = FieldType
    info = EncodingInfo(
        fields=[PointField("x", 0, F.FLOAT32, 0.001), PointField("y", 4, F.FLOAT32, 0.001), PointField("z", 8, F.FLOAT32, 0.001),
                PointField("intensity", 12, F.FLOAT32, 0.001), PointField("ring", 16, F.UINT16, None),
                PointField("time", 18, F.FLOAT32, 1e-5)],
        width=n, height=1, point_step=22, encoding_opt=EncodingOptions.LOSSY, compression_opt=CompressionOption.NONE,
        use_threads=False, version=5)
    return info, np.ascontiguousarray(buf).reshape(-1)


def cloud_livox(n: int = 40_000, seed: int = 8, version: int = 5):
    """Livox-style layout (x y z intensity f32, tag u8, line u8, offset_time u32; point_step 22): the uint8 fields are raw
    Copy bytes in the middle of the varint stream (field_encoder.hpp:51-67), so value boundaries cannot be ranked by
    terminator bits — the layout that exercises the pointer-jumping decoder. Random tag / line bytes on purpose (every
    bit pattern, including 0x00 and 0x80, appears as a raw byte)."""
    rng = np.random.default_rng(seed)
    step = 22
    buf = rng.integers(0, 256, (n, step), dtype=np.uint8)
    xyz = np.cumsum(rng.normal(0, 0.02, (n, 3)), axis=0).astype(np.float32)
    if n:
        k = max(1, n // 100)
        xyz[rng.integers(0, n, k), rng.integers(0, 3, k)] = np.nan
        xyz[rng.integers(0, n, max(1, n // 2000)), 0] = np.float32(-3.0e6)  # 5-byte varints
    buf[:, :12] = xyz.view(np.uint8).reshape(n, 12)
    buf[:, 12:16] = rng.integers(0, 255, n).astype(np.float32).view(np.uint8).reshape(n, 4)
    buf[:, 18:22] = (np.arange(n) * 100).astype(np.uint32).view(np.uint8).reshape(n, 4)
    F = FieldType
    info = EncodingInfo(
        fields=[PointField("x", 0, F.FLOAT32, 0.001), PointField("y", 4, F.FLOAT32, 0.001), PointField("z", 8, F.FLOAT32, 0.001),
                PointField("intensity", 12, F.FLOAT32, 0.01), PointField("tag", 16, F.UINT8, None), PointField("line", 17, F.UINT8, None),
                PointField("offset_time", 18, F.UINT32, None)],
        width=n, height=1, point_step=step, encoding_opt=EncodingOptions.LOSSY, compression_opt=CompressionOption.NONE,
        use_threads=False, version=version)
    return info, np.ascontiguousarray(buf).reshape(-1)


def cloud_lossless(n: int = 40_000, seed: int = 6, lossless: bool = True, version: int = 5, stamp_res=None,
                   hostile: bool = True):
    """The reference's DDS / PCD layout (x, y, z, intensity f32, ring u16, timestamp f64; point_step 26) used to reach
    the lossless float coders: EncodingOptions.LOSSLESS turns the f32 fields into XOR residuals and the resolution-less
    FLOAT64 into Gorilla (wire version >= 4) or XOR (version 3) (codec_common.cpp:100-140). `hostile` sprinkles the
    values that move the Gorilla window: repeats (xor == 0), NaN / inf / -0.0, denormals and full-entropy doubles."""
    rng = np.random.default_rng(seed)
    step = 26
    buf = np.zeros((n, step), dtype=np.uint8)
    xyz = np.cumsum(rng.normal(0, 0.02, (n, 3)), axis=0).astype(np.float32)
    inten = rng.integers(0, 255, n).astype(np.float32)
    ts = (1.7e9 + np.arange(n) * 1e-5 + (np.arange(n) // 2048) * 0.1).astype(np.float64)
    if hostile and n > 64:
        k = max(1, n // 200)
        ts[rng.integers(1, n, k)] = ts[rng.integers(0, n, k)]                 # jumps backwards / forwards
        idx = rng.integers(1, n, k); ts[idx] = ts[idx - 1]                     # exact repeats -> the single '0' bit
        ts[rng.integers(0, n, k)] = rng.integers(0, 2**63, k, dtype=np.int64).view(np.float64)  # random bit patterns
        ts[rng.integers(0, n, 8)] = np.array([np.nan, np.inf, -np.inf, -0.0, 0.0, 5e-324, 1.0, -1.0])
        xyz[rng.integers(0, n, k), rng.integers(0, 3, k)] = np.nan
        inten[rng.integers(0, n, 4)] = np.array([np.inf, -np.inf, -0.0, 1e-42], dtype=np.float32)
    buf[:, 0:12] = xyz.view(np.uint8).reshape(n, 12)
    buf[:, 12:16] = inten.view(np.uint8).reshape(n, 4)
    buf[:, 16:18] = (np.arange(n) % 64).astype(np.uint16).view(np.uint8).reshape(n, 2)
    buf[:, 18:26] = ts.view(np.uint8).reshape(n, 8)
    res = None if lossless else 0.001
    F = FieldType
    info = EncodingInfo(
        fields=[PointField("x", 0, F.FLOAT32, res), PointField("y", 4, F.FLOAT32, res), PointField("z", 8, F.FLOAT32, res),
                PointField("intensity", 12, F.FLOAT32, res), PointField("ring", 16, F.UINT16, None),
                PointField("timestamp", 18, F.FLOAT64, stamp_res)],
        width=n, height=1, point_step=step,
        encoding_opt=EncodingOptions.LOSSLESS if lossless else EncodingOptions.LOSSY,
        compression_opt=CompressionOption.NONE, use_threads=False, version=version)
    return info, np.ascontiguousarray(buf).reshape(-1)


def random_layout_case(seed: int):
    """A random EncodingInfo (field types, offsets with padding, resolutions, encoding option, wire version 3/4/5) with
    matching data: the planner's every branch is hit over a few hundred seeds (codec_common.cpp:69-198, v5_codec.cpp:883-892).
    Used to pin the C oracle against the compiled reference; the GPU path can be swept with the same seeds."""
    F = FieldType
    types = [F.INT8, F.UINT8, F.INT16, F.UINT16, F.INT32, F.UINT32, F.FLOAT32, F.FLOAT64, F.INT64, F.UINT64]
    npt = {F.INT8: np.int8, F.UINT8: np.uint8, F.INT16: np.int16, F.UINT16: np.uint16, F.INT32: np.int32, F.UINT32: np.uint32,
           F.FLOAT32: np.float32, F.FLOAT64: np.float64, F.INT64: np.int64, F.UINT64: np.uint64}
    size = {F.INT8: 1, F.UINT8: 1, F.INT16: 2, F.UINT16: 2, F.INT32: 4, F.UINT32: 4, F.FLOAT32: 4, F.FLOAT64: 8, F.INT64: 8, F.UINT64: 8}
    rng = np.random.default_rng(seed)
    nf = int(rng.integers(1, 7))
    n = int(rng.choice([1, 2, 17, 300, 4097, 33000 if seed % 7 == 0 else 500]))
    names = [(nm, F.FLOAT32) for nm in "xyz"] if rng.random() < 0.6 else []
    names += [(f"f{k}", types[int(rng.integers(0, len(types)))]) for k in range(nf)]
    fields, cols, off = [], [], 0
    for nm, t in names:
        off += int(rng.integers(0, 3)) if rng.random() < 0.3 else 0  # padding in front of the field
        res = None
        if t in (F.FLOAT32, F.FLOAT64):
            res = float(rng.choice([0.001, 0.01, 0.5])) if rng.random() < 0.7 else None
            v = np.cumsum(rng.normal(0, 0.05, n)).astype(npt[t])
            if n > 10:
                v[rng.integers(0, n, 2)] = np.nan
                if rng.random() < 0.3:
                    v[rng.integers(0, n)] = np.inf
        else:
            lim = np.iinfo(npt[t])
            kind = int(rng.integers(0, 4))
            if kind == 0:
                v = (np.arange(n) * int(rng.integers(1, 5)) + int(rng.integers(0, 100))).astype(npt[t])
            elif kind == 1:
                v = rng.integers(0, 4, n).astype(npt[t])
            elif kind == 2:
                v = ((np.arange(n) // 64) % 5).astype(npt[t])
            else:
                v = rng.integers(max(lim.min, -2**62), min(lim.max, 2**62), n, dtype=np.int64 if lim.min < 0 else np.uint64).astype(npt[t])
        fields.append(PointField(nm, off, t, res))
        cols.append((off, size[t], v))
        off += size[t]
    step = off + (int(rng.integers(0, 4)) if rng.random() < 0.3 else 0)
    buf = np.full((n, step), 0xCD, dtype=np.uint8)
    for o, sz, v in cols:
        buf[:, o:o + sz] = np.ascontiguousarray(v).view(np.uint8).reshape(n, sz)
    enc = [EncodingOptions.LOSSY, EncodingOptions.LOSSLESS, EncodingOptions.NONE][int(rng.integers(0, 3))]
    info = EncodingInfo(fields=fields, width=n, height=1, point_step=step, encoding_opt=enc,
                        compression_opt=CompressionOption.NONE, use_threads=False, version=int(rng.choice([3, 4, 5])))
    return info, np.ascontiguousarray(buf).reshape(-1)


def pointcloud2_msg(fields, point_step: int, cloud, width=None, height: int = 1, frame_id: str = "lidar_link",
                    stamp=(1700000000, 123456789), is_dense: bool = True, row_step=None, big_endian: bool = False) -> bytes:
    """A DDS (PLAIN_CDR) serialised sensor_msgs/msg/PointCloud2: 4-byte encapsulation header, then the members in
    declaration order, every primitive aligned to its size relative to the byte after the header. `fields` is a list of
    (name, offset, FieldType). Test input for the envelope functions (ros_msg_utils.cpp:54-95 reads exactly this)."""
    import struct
    e = ">" if big_endian else "<"
    b = bytearray([0, 0 if big_endian else 1, 0, 0])

    def align(n):
        while (len(b) - 4) % n:
            b.append(0)

    def u32(v):
        align(4)
        b.extend(struct.pack(e + "I", v & 0xFFFFFFFF))

    def string(t: str):
        raw = t.encode()
        u32(len(raw) + 1)
        b.extend(raw + b"\0")

    data = np.ascontiguousarray(cloud).view(np.uint8).reshape(-1)
    n = data.size // point_step if point_step else 0
    width = n // height if width is None else width
    align(4)
    b.extend(struct.pack(e + "i", stamp[0]))
    u32(stamp[1])
    string(frame_id)
    u32(height)
    u32(width)
    u32(len(fields))
    for name, offset, ftype in fields:
        string(name)
        u32(offset)
        b.append(int(ftype))
        u32(1)
    b.append(1 if big_endian else 0)
    u32(point_step)
    u32(point_step * width if row_step is None else row_step)
    u32(data.size)
    b.extend(data.tobytes())
    b.append(1 if is_dense else 0)
    return bytes(b)


def cloud_viz(n: int = 50_000, seed: int = 7, step: int = 16, dup_rate: float = 0.3, nan_rate: float = 0.02):
    """Input for applyVizLossyPreprocessing: a lidar-like XYZ cloud in which a share of the points repeats an EARLIER
    point's voxel (same coordinates +- a sub-resolution jitter) and a share has a NaN / inf coordinate. Extra bytes of
    the point (intensity / ring / padding) are random so that 'first occurrence wins' is visible in the output."""
    rng = np.random.default_rng(seed)
    xyz = np.stack(_lidar_xyz(n, rng), axis=1) if n else np.zeros((0, 3), dtype=np.float32)
    if n > 1:
        k = int(n * dup_rate)
        dst = rng.integers(1, n, k)
        src = (dst * rng.random(k)).astype(np.int64)  # an earlier index
        jitter = (rng.random((k, 3)).astype(np.float32) - 0.5) * np.float32(2e-4)
        # snap the source to a voxel centre first so that the jitter stays inside the voxel
        xyz[src] = np.round(xyz[src] * 1000.0).astype(np.float32) / np.float32(1000.0)
        xyz[dst] = xyz[src] + jitter
        bad = rng.integers(0, n, int(n * nan_rate))
        xyz[bad, rng.integers(0, 3, bad.size)] = rng.choice(np.array([np.nan, np.inf, -np.inf], dtype=np.float32), bad.size)
    buf = rng.integers(0, 256, (n, step), dtype=np.uint8)
    buf[:, :12] = xyz.view(np.uint8).reshape(n, 12)
    info = EncodingInfo(width=n, height=1, point_step=step, encoding_opt=EncodingOptions.LOSSY,
                        compression_opt=CompressionOption.NONE, use_threads=False)
    info.fields = [PointField("x", 0, FieldType.FLOAT32, 0.001), PointField("y", 4, FieldType.FLOAT32, 0.001),
                   PointField("z", 8, FieldType.FLOAT32, 0.001)]
    if step >= 16:
        info.fields.append(PointField("intensity", 12, FieldType.FLOAT32, 0.01))
    return info, buf.reshape(-1)


def fnv1a64(data) -> int:
    """FNV-1a 64-bit (the fingerprint mcap_codec_benchmark --hash prints, tools/src/mcap_codec_benchmark.cpp:103-109)."""
    h = 0xCBF29CE484222325
    for b in bytes(data):
        h = ((h ^ b) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return h
