"""SURVEY.md §8(f) N2, host side (no GPU): the CDR parser / profile logic of the DDS envelope against the compiled
reference (cloudini_lib/src/ros_msg_utils.cpp:54-131,217-238) and against the committed golden messages."""
import os

import numpy as np
import pytest

import cloudini_b200 as cb
from cloudini_b200 import ros, synth
from cloudini_b200 import FieldType as FT

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

XYZI = [("x", 0, FT.FLOAT32), ("y", 4, FT.FLOAT32), ("z", 8, FT.FLOAT32), ("intensity", 12, FT.FLOAT32)]
VELO = [("x", 0, FT.FLOAT32), ("y", 4, FT.FLOAT32), ("z", 8, FT.FLOAT32), ("intensity", 12, FT.FLOAT32),
        ("ring", 16, FT.UINT16), ("time", 18, FT.FLOAT32)]


def _describe(pc: ros.RosPointCloud2) -> str:  # the same text oracle/ref_wrapper.cpp::ref_ros_describe prints
    t = f"stamp {pc.stamp_sec} {pc.stamp_nsec}\nframe_id {pc.frame_id}\nheight {pc.height} width {pc.width}\n"
    for f in pc.fields:
        t += f"field {f.name} {f.offset} {int(f.type)}\n"
    t += f"point_step {pc.point_step} row_step {pc.row_step}\n"
    t += f"data {pc.data_offset} {pc.data.size}\nis_dense {1 if pc.is_dense else 0}\n"
    return t


def _messages():
    rng = np.random.default_rng(5)
    yield synth.pointcloud2_msg(XYZI, 16, synth.cloud_c2(1000, seed=3)[1])
    yield synth.pointcloud2_msg(VELO, 22, rng.integers(0, 255, 22 * 777, dtype=np.uint8), frame_id="velodyne", is_dense=False)
    yield synth.pointcloud2_msg(XYZI[:3], 12, synth.cloud_c1(64 * 10, seed=1)[1], width=64, height=10, frame_id="")
    yield synth.pointcloud2_msg(XYZI, 16, np.zeros(0, dtype=np.uint8), frame_id="a_rather_long_frame_identifier/with/slashes")
    yield synth.pointcloud2_msg(XYZI, 16, synth.cloud_c2(10, seed=3)[1], big_endian=True)


def test_parse_matches_reference(ref):
    for msg in _messages():
        assert _describe(ros.getDeserializedPointCloudMessage(msg)) == ref.ros_describe(msg)


def test_parse_errors_match_reference(ref):
    good = synth.pointcloud2_msg(XYZI, 16, synth.cloud_c2(100, seed=3)[1])
    for cut in (0, 3, 4, 11, 20, 27, 40, 90, len(good) - 1601, len(good) - 1):
        bad = good[:cut]
        with pytest.raises(RuntimeError):
            ref.ros_describe(bad) if cut >= 4 else (_ for _ in ()).throw(RuntimeError())  # nanocdr reads 4 header bytes unchecked
        with pytest.raises(RuntimeError):
            ros.getDeserializedPointCloudMessage(bad)
    for patch in ((0, 1), (1, 2), (1, 4), (2, 1), (3, 9)):  # first byte != 0, PL_CDR, PLAIN_CDR2, extended header
        bad = bytearray(good)
        bad[patch[0]] = patch[1]
        with pytest.raises(RuntimeError):
            ref.ros_describe(bytes(bad))
        with pytest.raises(RuntimeError):
            ros.getDeserializedPointCloudMessage(bytes(bad))


def test_non_canonical_inputs_found_by_the_fuzzer():
    good = synth.pointcloud2_msg(XYZI, 16, synth.cloud_c2(100, seed=3)[1])
    # is_dense byte that is neither 0 nor 1: kept as it came in (the reference writes it back unchanged), still truthy
    odd = bytearray(good)
    odd[-1] = 0x7C
    pc = ros.getDeserializedPointCloudMessage(bytes(odd))
    assert pc.is_dense == 0x7C and pc.is_dense and pc._c_view()[0].is_dense == 0x7C
    assert ros.getDeserializedPointCloudMessage(good).is_dense is True
    # a field name whose CDR string has a NUL inside ("x\0" declared 4 bytes long): refused, not silently shortened
    at = good.index(b"\x02\x00\x00\x00x\x00")
    forged = bytearray(good)
    forged[at] = 4
    with pytest.raises(RuntimeError, match="embedded NUL"):
        ros.getDeserializedPointCloudMessage(bytes(forged))


def test_to_encoding_info_and_profile():
    pc = ros.getDeserializedPointCloudMessage(synth.pointcloud2_msg(VELO, 22, np.zeros(22 * 5, dtype=np.uint8)))
    ros.applyResolutionProfile({"intensity": 0.0, "time": 0.5, "ring": 2.0}, pc.fields, 0.001)
    assert [(f.name, f.resolution) for f in pc.fields] == [("x", pytest.approx(0.001)), ("y", pytest.approx(0.001)),
                                                             ("z", pytest.approx(0.001)), ("ring", 2.0), ("time", 0.5)]
    info = ros.toEncodingInfo(pc)
    assert info.encoding_opt == cb.EncodingOptions.LOSSY and info.compression_opt == cb.CompressionOption.ZSTD  # ros_msg_utils.cpp:127-128
    assert (info.width, info.height, info.point_step) == (5, 1, 22) and [f.name for f in info.fields] == ["x", "y", "z", "ring", "time"]
    pc2 = ros.getDeserializedPointCloudMessage(synth.pointcloud2_msg(VELO, 22, np.zeros(22 * 5, dtype=np.uint8)))
    ros.applyResolutionProfile({}, pc2.fields, None)
    assert all(f.resolution is None for f in pc2.fields)


def test_golden_messages_parse(golden_ros):
    for name, g in golden_ros.items():
        pc = ros.getDeserializedPointCloudMessage(g["msg"])
        assert _describe(pc) == g["describe"], name


def test_wasm_shaped_host_functions(ref, golden_ros):
    # cldn_GetHeaderAsYAML / cldn_GetHeaderAsYAMLFromDDS / cldn_GetDecompressedSize (wasm_functions.cpp:24-56,95-106): host only
    import ctypes as C
    L = cb.lib()
    for f in (L.cldn_b200_GetHeaderAsYAML, L.cldn_b200_GetHeaderAsYAMLFromDDS):
        f.restype, f.argtypes = C.c_uint32, [C.c_void_p, C.c_uint32, C.c_char_p, C.c_uint32]
    L.cldn_b200_GetDecompressedSize.restype, L.cldn_b200_GetDecompressedSize.argtypes = C.c_uint32, [C.c_void_p, C.c_uint32]
    for name, g in golden_ros.items():
        comp = np.frombuffer(g["compressed"], dtype=np.uint8)
        pc = ros.getDeserializedPointCloudMessage(g["compressed"])
        assert L.cldn_b200_GetDecompressedSize(comp.ctypes.data, comp.size) == pc.width * pc.height * pc.point_step
        buf = C.create_string_buffer(1 << 16)
        n = L.cldn_b200_GetHeaderAsYAMLFromDDS(comp.ctypes.data, comp.size, buf, len(buf))
        if pc.data.size == 0:   # empty cloud: no blob, hence no header -> 0 like the reference (DecodeHeader throws)
            assert n == 0
            continue
        blob = np.ascontiguousarray(pc.data)
        assert n > 0 and n == L.cldn_b200_GetHeaderAsYAML(blob.ctypes.data, blob.size, buf, len(buf))
        text = buf.raw[:n].decode()
        assert text == cb.EncodingInfoToYAML(cb.DecodeHeader(bytes(blob))[0])
        assert L.cldn_b200_GetHeaderAsYAML(blob.ctypes.data, blob.size, buf, 10) == 0          # capacity too small
    junk = np.zeros(64, dtype=np.uint8)
    assert L.cldn_b200_GetHeaderAsYAML(junk.ctypes.data, junk.size, C.create_string_buffer(64), 64) == 0
    assert L.cldn_b200_GetDecompressedSize(junk.ctypes.data, 3) == 0
