"""Generates tests/golden/golden_v1.npz and golden_v2.npz from the UNMODIFIED reference (oracle/_ref/libcloudini_ref.so).

Run in the build container (where /root/reference exists):  python tests/golden/make_golden.py
Each case stores the EncodingInfo (as the reference's YAML text + numeric version), the raw input bytes and the
exact blob the reference encoder produced (CompressionOption::NONE, use_threads=false). Decoded outputs are not stored:
tests re-derive them with the oracle and compare against a zero-initialised buffer.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import cloudini_b200 as cb  # noqa: E402
from cloudini_b200 import synth  # noqa: E402
from oracle.client import RefOracle, build_ref  # noqa: E402


def int_field_cloud(values: np.ndarray, ftype: cb.FieldType, version=5):
    """XYZ(constant) + one integer field, the layout of the reference's V5 mode tests (test_field_encoders.cpp:590-674)."""
    n = len(values)
    size = cb.SizeOf(ftype)
    step = 12 + size
    buf = np.zeros((n, step), dtype=np.uint8)
    xyz = np.stack([np.arange(n) * 0.001, np.full(n, 1.0), np.full(n, -3.0)], axis=1).astype(np.float32)
    buf[:, :12] = xyz.view(np.uint8).reshape(n, 12)
    buf[:, 12:] = values.view(np.uint8).reshape(n, size)
    info = cb.EncodingInfo(
        fields=[cb.PointField("x", 0, cb.FieldType.FLOAT32, 0.001), cb.PointField("y", 4, cb.FieldType.FLOAT32, 0.001),
                cb.PointField("z", 8, cb.FieldType.FLOAT32, 0.001), cb.PointField("v", 12, ftype, None)],
        width=n, height=1, point_step=step, compression_opt=cb.CompressionOption.NONE, use_threads=False, version=version)
    return info, buf.reshape(-1)


def cases():
    out = {}
    # SURVEY.md §8(c) known-answer vectors
    xyz3 = np.array([[1, -2, 0.5], [1.001, -2, 0.5], [1.001, -1, 100]], dtype=np.float32)
    out["xyz3"] = (synth.info_xyz(3), xyz3.view(np.uint8).reshape(-1))
    ties = np.array([[0.25, 0.75, 1.25]], dtype=np.float32)
    out["ties_even"] = (synth.info_xyz(1, 0.5), ties.view(np.uint8).reshape(-1))
    edge = np.array([[5e-4, 1.5e-3, 2.5e-3], [np.nan, 1, 1], [2, 1, np.inf], [2, 1, 3e9], [2, 1, -3e6]], dtype=np.float32)
    out["nan_inf"] = (synth.info_xyz(5), edge.view(np.uint8).reshape(-1))
    out["c1"] = synth.cloud_c1()
    out["c1_adv"] = synth.cloud_c1(10_000, seed=7, adversarial=True)
    out["c2_40k"] = synth.cloud_c2(40_000, seed=2)
    out["c3_40k"] = synth.cloud_c3(40_000, seed=3)
    out["c3_40k_v4"] = synth.cloud_c3(40_000, seed=3, version=4)
    n = 32768 + 19
    i = np.arange(n)
    out["mode_linear_u32"] = int_field_cloud((100000 + 3 * i).astype(np.uint32), cb.FieldType.UINT32)
    out["mode_palette_u32"] = int_field_cloud((i % 4).astype(np.uint32), cb.FieldType.UINT32)
    out["mode_rle_u16"] = int_field_cloud(((i // 256) % 8).astype(np.uint16), cb.FieldType.UINT16)
    out["mode_desc_i32"] = int_field_cloud((200000 - 5 * i).astype(np.int32), cb.FieldType.INT32)
    rng = np.random.default_rng(11)
    out["mode_random_u16"] = int_field_cloud(rng.integers(0, 65536, size=n).astype(np.uint16), cb.FieldType.UINT16)
    out["mode_i64"] = int_field_cloud((rng.integers(-2**40, 2**40, size=5000)).astype(np.int64), cb.FieldType.INT64)
    # real sensor data: an excerpt of the reference's sample cloud (binary PCD body = packed XYZI float32)
    pcd = os.path.join(os.environ.get("CLOUDINI_REFERENCE", "/root/reference"), "cloudini_lib/samples/lidar.pcd")
    if os.path.exists(pcd):
        raw = open(pcd, "rb").read()
        body = raw[raw.index(b"DATA binary\n") + len(b"DATA binary\n"):]
        npts = 36000
        out["lidar_36k"] = (synth.info_xyzi(npts), np.frombuffer(body[:npts * 16], dtype=np.uint8).copy())
    return out


def cases_v2():
    """Lossless float coders (XOR / Gorilla, field_encoder.hpp:123-312) on the DDS-like x,y,z,intensity,ring,timestamp
    layout of the reference's sample message: every wire version, LOSSLESS and LOSSY (the resolution-less FLOAT64 stays
    Gorilla / XOR either way), with the hostile values (repeats, NaN / inf / -0.0, denormals, random bit patterns)."""
    out = {}
    for version in (5, 4, 3):
        out[f"lossless_v{version}"] = synth.cloud_lossless(3000, seed=40 + version, lossless=True, version=version)
        out[f"gorilla_lossy_v{version}"] = synth.cloud_lossless(3000, seed=50 + version, lossless=False, version=version)
    out["lossless_v5_2chunks"] = synth.cloud_lossless(33_000, seed=60, lossless=True, version=5)
    return out


def write(ref, path, table):
    blob = {}
    for name, (info, cloud) in table.items():
        enc = ref.encode(info, cloud)
        blob[name + "__yaml"] = np.frombuffer(cb.EncodingInfoToYAML(info).encode(), dtype=np.uint8)
        blob[name + "__version"] = np.array([info.version], dtype=np.int32)
        blob[name + "__input"] = np.asarray(cloud, dtype=np.uint8)
        blob[name + "__blob"] = np.frombuffer(enc, dtype=np.uint8)
        print(f"{name:18s} points={info.width:6d} step={info.point_step:2d} blob={len(enc)}")
    np.savez_compressed(path, **blob)
    print("wrote", path, os.path.getsize(path), "bytes")


def main():
    build_ref()
    ref = RefOracle()
    here = os.path.dirname(os.path.abspath(__file__))
    if "--v2-only" not in sys.argv:
        write(ref, os.path.join(here, "golden_v1.npz"), cases())
    write(ref, os.path.join(here, "golden_v2.npz"), cases_v2())


if __name__ == "__main__":
    main()
