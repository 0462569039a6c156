"""Generates tests/golden/golden_ros_v1.npz from the UNMODIFIED reference (oracle/_ref/libcloudini_ref.so).

Run in the build container (where /root/reference exists):  python tests/golden/make_golden_ros.py
Each case stores a DDS-serialised PointCloud2 message, the conversion options, and what the reference's converter step
(tools/src/mcap_converter.cpp:184-204 -> ros_msg_utils.cpp) made of it: the parsed description, the
CompressedPointCloud2 message (CompressionOption::NONE so the bytes are library-version independent) and the PointCloud2
message restored from it. One case is an excerpt of the reference's own samples/dds_message.bin.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from cloudini_b200 import ros, synth  # noqa: E402
from cloudini_b200 import FieldType as FT  # noqa: E402
from oracle.client import RefOracle, build_ref  # noqa: E402

XYZI = [("x", 0, FT.FLOAT32), ("y", 4, FT.FLOAT32), ("z", 8, FT.FLOAT32), ("intensity", 12, FT.FLOAT32)]


def cases():
    out = {}
    out["xyzi_viz"] = (synth.pointcloud2_msg(XYZI, 16, synth.cloud_viz(6000, seed=5)[1]), {}, 0.001, True)
    out["xyzi_plain"] = (synth.pointcloud2_msg(XYZI, 16, synth.cloud_c2(5000, seed=2)[1], frame_id="base"), {"intensity": 0.5}, 0.002, False)
    out["xyz_organized"] = (synth.pointcloud2_msg(XYZI[:3], 12, synth.cloud_c1(640, seed=1)[1], width=64, height=10, frame_id="", is_dense=False), {}, 0.001, False)
    out["empty"] = (synth.pointcloud2_msg(XYZI, 16, np.zeros(0, dtype=np.uint8)), {}, 0.001, True)
    # the reference's own sample message, re-serialised with its first 4000 points (the file itself is 1.6 MB)
    sample = os.path.join(os.environ.get("CLOUDINI_REFERENCE", "/root/reference"), "cloudini_lib/samples/dds_message.bin")
    if os.path.exists(sample):
        raw = open(sample, "rb").read()
        pc = ros.getDeserializedPointCloudMessage(raw)
        n = 4000
        fields = [(f.name, f.offset, f.type) for f in pc.fields]
        msg = synth.pointcloud2_msg(fields, pc.point_step, np.array(pc.data[:n * pc.point_step]), frame_id=pc.frame_id,
                                    stamp=(pc.stamp_sec, pc.stamp_nsec), is_dense=pc.is_dense)
        out["dds_sample_4000"] = (msg, {}, 0.001, True)
    return out


def main():
    build_ref()
    ref = RefOracle()
    arrays = {}
    for name, (msg, profile, res, viz) in cases().items():
        comp = ref.ros_compress(msg, profile, res, viz, 1, 0, 5)
        arrays[name + "__msg"] = np.frombuffer(msg, dtype=np.uint8)
        arrays[name + "__profile"] = np.frombuffer(";".join(f"{k}={v!r}" for k, v in profile.items()).encode(), dtype=np.uint8)
        arrays[name + "__opts"] = np.array([res, 1.0 if viz else 0.0], dtype=np.float64)
        arrays[name + "__describe"] = np.frombuffer(ref.ros_describe(msg).encode(), dtype=np.uint8)
        arrays[name + "__compressed"] = np.frombuffer(comp, dtype=np.uint8)
        arrays[name + "__restored"] = np.frombuffer(ref.ros_decompress(comp, len(msg) + 4096), dtype=np.uint8)
        print(f"{name}: msg {len(msg)} B -> compressed {len(comp)} B")
    # N3 alone: raw clouds through applyVizLossyPreprocessing (input, EncodingInfo as YAML, survivors, info afterwards)
    import cloudini_b200 as cb
    for name, (info, cloud) in {"viz_xyzi_8k": synth.cloud_viz(8000, seed=31), "viz_step32_3k": synth.cloud_viz(3000, seed=32, step=32),
                                "viz_xyz_500": synth.cloud_viz(500, seed=33, step=12)}.items():
        after, kept = ref.viz_preprocess(info, cloud)
        arrays["VIZ_" + name + "__input"] = cloud
        arrays["VIZ_" + name + "__yaml"] = np.frombuffer(cb.EncodingInfoToYAML(info).encode(), dtype=np.uint8)
        arrays["VIZ_" + name + "__yaml_after"] = np.frombuffer(cb.EncodingInfoToYAML(after).encode(), dtype=np.uint8)
        arrays["VIZ_" + name + "__output"] = kept
        print(f"{name}: {info.width} -> {after.width} points")
    path = os.path.join(ROOT, "tests", "golden", "golden_ros_v1.npz")
    np.savez_compressed(path, **arrays)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
