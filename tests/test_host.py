"""CPU tests of the host-side logic of the C-ABI library (no kernels are launched):
the library loads, exports every symbol include/cloudini_b200.h declares, and its header / YAML / sizing / planning
helpers agree with the oracle. Compute entry points must fail loudly without a GPU (there is no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import cloudini_b200 as cb
from cloudini_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(lib_built):
    text = "".join(open(os.path.join(ROOT, "include", h)).read() for h in ("cloudini_b200.h", "cloudini_b200_ros.h"))
    declared = sorted(set(re.findall(r"\b(cldn_b200_\w+)\s*\(", text)))
    assert len(declared) >= 28
    L = C.CDLL(lib_built)
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, missing


def test_info_struct_layout_matches_header(lib_built):
    # the ctypes mirror must match the C struct: round-trip through the library's own YAML writer/parser
    info, _ = synth.cloud_c3(10)
    info.encoding_config = "abc"
    y = cb.EncodingInfoToYAML(info)
    back = cb.EncodingInfoFromYAML(y)
    assert cb.EncodingInfoToYAML(back) == y
    assert [f.name for f in back.fields] == ["x", "y", "z", "rgba", "ring"]
    assert back.fields[0].resolution == pytest.approx(0.001) and back.fields[3].resolution is None


def test_header_yaml_text(lib_built, port):
    # test_header.cpp:107-140 (YAML round trip) and :142-163 (default V05 vs explicit V04)
    for version in (5, 4, 3):
        info, _ = synth.cloud_c3(123, version=version)
        h = cb.EncodeHeader(info)
        assert h.startswith(b"CLOUDINI_V0%d\n" % version) and h.endswith(b"\0")
        assert h == port.header(info)
        parsed, used = cb.DecodeHeader(h + b"payload")
        assert used == len(h) and parsed.version == version
        assert cb.EncodingInfoToYAML(parsed) == cb.EncodingInfoToYAML(info)
    y = cb.EncodingInfoToYAML(synth.info_xyz(7))
    assert "resolution: 0.001\n" in y and y.startswith("version: 5\nwidth: 7\nheight: 1\npoint_step: 12\nencoding_opt: LOSSY\n")


def test_header_errors(lib_built):
    with pytest.raises(RuntimeError, match="too small"):
        cb.DecodeHeader(b"CLOUD")
    with pytest.raises(RuntimeError, match="magic"):
        cb.DecodeHeader(b"NOTCLOUDINI_V05\nversion: 5\n\0")
    with pytest.raises(RuntimeError, match="[Uu]nsupported encoding version"):
        cb.DecodeHeader(b"CLOUDINI_V09\nversion: 9\n\0")
    with pytest.raises(RuntimeError, match="null terminator"):  # test_header.cpp:243-262
        cb.DecodeHeader(cb.EncodeHeader(synth.info_xyz(3))[:-1])


def test_legacy_binary_header(lib_built):
    # cloudini.cpp:319-343 / 395-427: binary header written by old encoders is still readable
    import struct
    name = b"x"
    blob = b"CLOUDINI_V03" + struct.pack("<IIIBBH", 11, 1, 4, 1, 0, 1) + struct.pack("<H", len(name)) + name + struct.pack("<IBf", 0, 7, 0.01)
    info, used = cb.DecodeHeader(blob + b"xx")
    assert used == len(blob) and info.version == 3 and info.width == 11 and info.point_step == 4
    assert info.fields[0].name == "x" and info.fields[0].type == cb.FieldType.FLOAT32
    assert info.fields[0].resolution == pytest.approx(0.01)


@pytest.mark.parametrize("comp", [cb.CompressionOption.NONE, cb.CompressionOption.LZ4, cb.CompressionOption.ZSTD])
def test_max_compressed_size_matches_oracle(lib_built, port, comp):
    for info, _ in (synth.cloud_c1(1), synth.cloud_c2(1), synth.cloud_c3(1), synth.cloud_c3(1, version=4)):
        info.compression_opt = comp
        for n in (0, 1, 1000, 32768, 32769, 1_000_000):
            for hdr in (True, False):
                assert cb.MaxCompressedSize(info, n, hdr) == port.max_compressed_size(info, n, hdr)


def test_max_compressed_size_matches_reference(lib_built, ref):
    for comp in cb.CompressionOption:
        for info, _ in (synth.cloud_c2(1), synth.cloud_c3(1)):
            info.compression_opt = comp
            for n in (0, 5, 40_000, 1_000_000):
                assert cb.MaxCompressedSize(info, n, True) == ref.max_compressed_size(info, n, True)


@pytest.mark.parametrize("version", [3, 4, 5])
@pytest.mark.parametrize("lossless", [True, False])
def test_lossless_layout_header_and_sizing_match_reference(lib_built, ref, version, lossless):
    # the DDS-like layout that reaches the XOR / Gorilla coders: header text and capacity rule are host-side logic
    info, cloud = synth.cloud_lossless(1000, seed=1, lossless=lossless, version=version)
    blob = ref.encode(info, cloud)
    assert blob.startswith(cb.EncodeHeader(info))  # the reference writes exactly this header in front of the chunks
    back, used = cb.DecodeHeader(cb.EncodeHeader(info) + b"xyz")
    assert used == len(cb.EncodeHeader(info)) and back.encoding_opt == info.encoding_opt
    for a, b in zip(back.fields, info.fields):  # resolutions are float32 in the header
        assert (a.resolution is None) == (b.resolution is None)
        assert a.resolution is None or np.float32(a.resolution) == np.float32(b.resolution)
    for n in (0, 1, 5, 4133, 100_000):
        for hdr in (False, True):
            assert cb.MaxCompressedSize(info, n, hdr) == ref.max_compressed_size(info, n, hdr)
    # the stage-1 blob of the reference fits the capacity rule (sanity of the rule itself for these coders)
    assert len(blob) <= cb.MaxCompressedSize(info, 1000, True)


def test_binary_header_writer_matches_reference(lib_built, ref):
    # EncodeHeader(..., HeaderEncoding::BINARY) (cloudini.cpp:319-344): byte-identical, and readable by both readers
    cases = [synth.cloud_c1(11)[0], synth.cloud_c3(12)[0], synth.cloud_c3(13, version=4)[0],
             synth.cloud_lossless(14, lossless=True, version=3)[0], synth.random_layout_case(5)[0]]
    for info in cases:
        for comp in cb.CompressionOption:
            info.compression_opt = comp
            h = cb.EncodeHeader(info, binary=True)
            assert h == ref.header(info, binary=True)
            assert cb.EncodeHeader(info) == ref.header(info, binary=False)
            mine, used = cb.DecodeHeader(h + b"tail")
            theirs, used_ref = ref.decode_header(h + b"tail")
            assert used == used_ref == len(h)
            assert cb.EncodingInfoToYAML(mine) == cb.EncodingInfoToYAML(theirs)
            assert (mine.width, mine.point_step, mine.version, len(mine.fields)) == (info.width, info.point_step, info.version, len(info.fields))


def test_binary_header_width_10_is_ambiguous_like_the_reference(lib_built, ref):
    # a binary header whose width's low byte is 0x0A ('\n') is taken for a YAML header by DecodeHeader
    # (cloudini.cpp:375-377): the reference fails to read back what it wrote, and so do we (the YAML parsers word the
    # complaint differently)
    info = synth.cloud_c1(10)[0]
    h = cb.EncodeHeader(info, binary=True)
    assert h == ref.header(info, binary=True)
    with pytest.raises(RuntimeError):
        ref.decode_header(h)
    with pytest.raises(RuntimeError):
        cb.DecodeHeader(h)


def test_point_step_zero_rejected(lib_built):
    info = synth.info_xyz(1)
    info.point_step = 0
    with pytest.raises(RuntimeError, match="point_step cannot be 0"):
        cb.MaxCompressedSize(info, 10)


def test_no_cpu_fallback_without_gpu(lib_built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    info, cloud = synth.cloud_c1(100)
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        cb.PointcloudEncoder(info)
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        cb.PointcloudDecoder()
    from cloudini_b200 import ros
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        ros.VizPreprocessor()


def test_damaged_field_list_is_refused_not_reinterpreted(lib_built):
    # found by tests/fuzz/fuzz_header_text.py: a damaged line inside "fields:" used to be skipped, the keys behind it landed
    # in the previous field and the blob decoded — silently — with a layout nobody wrote
    info, _ = synth.cloud_c2(100, seed=1)
    blob = cb.EncodeHeader(info)
    assert cb.DecodeHeader(blob)[0].fields[1].name == "y"
    for old, new in ((b"  - name: y", b"  Q name: y"), (b"  - name: y", b"E - name: y"), (b"fields:\n", b"fields:*\n"),
                     (b"  - name: x", b"  - nam8: x"), (b"    offset: 4", b"    offset: 4\n    offset: 8"), (b"offset: 12", b"offset: 1:"),
                     (b"  - name: z", b"  - name:  "), (b"    type: FLOAT32\n    resolution: 0.001\n  - name: y", b"    type FLOAT32\n    resolution: 0.001\n  - name: y")):
        bad = blob.replace(old, new, 1)
        assert bad != blob, old
        with pytest.raises(RuntimeError):
            cb.DecodeHeader(bad)


def test_forged_headers_are_refused_not_executed(lib_built):
    # a field that does not fit inside the point would make the decoders write past the output buffer (the reference
    # does exactly that: field_decoder.cpp:74-78 has no bound). Planning refuses such an EncodingInfo on both sides.
    info, cloud = synth.cloud_c2(100, seed=1)
    blob = cb.EncodeHeader(info)
    for bad in (blob.replace(b"offset: 12", b"offset: 92"), blob.replace(b"point_step: 16", b"point_step: 11"),
                blob.replace(b"offset: 8", b"offset: 4294967290")):
        dinfo, _ = cb.DecodeHeader(bad + b"\0" * 8)   # the text itself parses
        with pytest.raises(RuntimeError, match="does not fit a point"):
            cb.MaxCompressedSize(dinfo, 10) and cb.PointcloudEncoder(dinfo)
    # required keys (the reference's as<>() of a missing node throws "Node is not a string", cloudini.cpp:199-221)
    for drop in (b"    offset: 12\n", b"    resolution: 0.001\n", b"compression_opt: NONE\n", b"point_step: 16\n"):
        text = blob[13:].rstrip(b"\0").replace(drop, b"", 1)
        with pytest.raises(RuntimeError, match="Node is not a string|missing"):
            cb.EncodingInfoFromYAML(text.decode())
    # numbers: digits then anything (iss >> uint32_t), but at least one digit
    assert cb.EncodingInfoFromYAML(blob[13:].rstrip(b"\0").replace(b"width: 100", b"width: 10O").decode()).width == 10
    with pytest.raises(RuntimeError, match="Failed to convert scalar"):
        cb.EncodingInfoFromYAML(blob[13:].rstrip(b"\0").replace(b"width: 100", b"width: x100").decode())
