"""TEST INFRASTRUCTURE ONLY: ctypes clients for the two checkers.

* ``RefOracle``  — oracle/_ref/libcloudini_ref.so: the UNMODIFIED reference compiled in place (oracle/build_ref.sh).
* ``PortOracle`` — oracle/_build/libcloudini_oracle.so: the plain-C restatement (oracle/cloudini_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

import cloudini_b200 as cb

HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB = os.path.join(HERE, "_ref", "libcloudini_ref.so")
PORT_LIB = os.path.join(HERE, "_build", "libcloudini_oracle.so")
REFERENCE_TREE = os.environ.get("CLOUDINI_REFERENCE", "/root/reference")


def build_port(force: bool = False) -> str:
    src = os.path.join(HERE, "cloudini_oracle.c")
    if force or not os.path.exists(PORT_LIB) or os.path.getmtime(PORT_LIB) < os.path.getmtime(src):
        subprocess.check_call([os.path.join(HERE, "build_oracle.sh")], stdout=subprocess.DEVNULL)
    return PORT_LIB


def build_ref(force: bool = False):
    """Builds the reference-backed oracle when /root/reference is present (never on the GPU box)."""
    have_tree = os.path.isdir(os.path.join(REFERENCE_TREE, "cloudini_lib", "src"))
    wrapper = os.path.join(HERE, "ref_wrapper.cpp")
    stale = os.path.exists(REF_LIB) and have_tree and os.path.getmtime(REF_LIB) < os.path.getmtime(wrapper)
    if os.path.exists(REF_LIB) and not force and not stale:
        return REF_LIB
    if not have_tree:
        return REF_LIB if os.path.exists(REF_LIB) else None
    subprocess.check_call([os.path.join(HERE, "build_ref.sh")], stdout=subprocess.DEVNULL)
    return REF_LIB


def have_ref() -> bool:
    return os.path.exists(REF_LIB)


def _u8(buf) -> np.ndarray:
    return np.frombuffer(buf, dtype=np.uint8) if isinstance(buf, (bytes, bytearray, memoryview)) else np.ascontiguousarray(buf).view(np.uint8).reshape(-1)


class RefOracle:
    kind = "reference"

    def __init__(self):
        if not os.path.exists(REF_LIB):
            raise RuntimeError(f"{REF_LIB} not built (run oracle/build_ref.sh where /root/reference exists)")
        L = C.CDLL(REF_LIB)
        L.ref_last_error.restype = C.c_char_p
        L.ref_encode.restype = C.c_longlong
        L.ref_encode.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]
        L.ref_max_compressed_size.restype = C.c_size_t
        L.ref_max_compressed_size.argtypes = [C.c_char_p, C.c_int, C.c_size_t, C.c_int]
        if hasattr(L, "ref_encode_header"):  # older prebuilt wrappers do not have it
            L.ref_encode_header.restype = C.c_longlong
            L.ref_encode_header.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
        L.ref_decode_header.restype = C.c_longlong
        L.ref_decode_header.argtypes = [C.c_void_p, C.c_size_t, C.c_char_p, C.c_size_t, C.POINTER(C.c_int)]
        L.ref_decode.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.ref_decode_payload.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.ref_time_encode.restype = C.c_double
        L.ref_time_encode.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.POINTER(C.c_size_t)]
        L.ref_time_decode.restype = C.c_double
        L.ref_time_decode.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int]
        if hasattr(L, "ref_viz_preprocess"):  # N2 / N3 (ros_msg_utils.cpp); older prebuilt wrappers do not have them
            L.ref_viz_preprocess.restype = C.c_longlong
            L.ref_viz_preprocess.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_char_p, C.c_size_t]
            L.ref_ros_describe.restype = C.c_longlong
            L.ref_ros_describe.argtypes = [C.c_void_p, C.c_size_t, C.c_char_p, C.c_size_t]
            L.ref_ros_compress.restype = C.c_longlong
            L.ref_ros_compress.argtypes = [C.c_void_p, C.c_size_t, C.c_char_p, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t]
            L.ref_ros_decompress.restype = C.c_longlong
            L.ref_ros_decompress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        self.L = L

    def _err(self):
        return self.L.ref_last_error().decode("utf-8", "replace")

    # ---- N3 / N2: cloudini_ros (ros_msg_utils.cpp) ----
    def viz_preprocess(self, info: cb.EncodingInfo, cloud):
        """applyVizLossyPreprocessing on a raw cloud. Returns (EncodingInfo after, surviving bytes)."""
        a = _u8(cloud)
        out = np.zeros(max(a.size, 1), dtype=np.uint8)
        ybuf = C.create_string_buffer(1 << 16)
        n = self.L.ref_viz_preprocess(_yaml(info), int(info.version), a.ctypes.data, a.size, out.ctypes.data, out.size, ybuf, len(ybuf))
        if n < 0:
            raise RuntimeError(self._err())
        after = cb.EncodingInfoFromYAML(ybuf.value.decode())
        after.version = info.version
        return after, out[:n * info.point_step]

    def ros_describe(self, msg: bytes) -> str:
        a = _u8(msg)
        buf = C.create_string_buffer(1 << 16)
        if self.L.ref_ros_describe(a.ctypes.data, a.size, buf, len(buf)) < 0:
            raise RuntimeError(self._err())
        return buf.value.decode("utf-8", "replace")

    def ros_compress(self, msg: bytes, profile=None, default_resolution=None, viz=False, encoding_opt=1, compression_opt=0, version=5) -> bytes:
        """The converter's per-message step (tools/src/mcap_converter.cpp:184-204)."""
        a = _u8(msg)
        text = ";".join(f"{k}={float(v)!r}" for k, v in (profile or {}).items()).encode()
        out = np.zeros(2 * a.size + (1 << 16), dtype=np.uint8)
        n = self.L.ref_ros_compress(a.ctypes.data, a.size, text, -1.0 if default_resolution is None else float(default_resolution),
                                    1 if viz else 0, int(encoding_opt), int(compression_opt), int(version), out.ctypes.data, out.size)
        if n < 0:
            raise RuntimeError(self._err())
        return bytes(out[:n])

    def ros_decompress(self, msg: bytes, capacity: int) -> bytes:
        a = _u8(msg)
        out = np.zeros(capacity, dtype=np.uint8)
        n = self.L.ref_ros_decompress(a.ctypes.data, a.size, out.ctypes.data, out.size)
        if n < 0:
            raise RuntimeError(self._err())
        return bytes(out[:n])

    def max_compressed_size(self, info: cb.EncodingInfo, points: int, include_header: bool = True) -> int:
        return int(self.L.ref_max_compressed_size(_yaml(info), info.version, points, int(include_header)))

    def encode(self, info: cb.EncodingInfo, cloud, write_header: bool = True, cap: int = None) -> bytes:
        data = _u8(cloud)
        n = data.nbytes // info.point_step if info.point_step else 0
        cap = self.max_compressed_size(info, n, True) + 64 if cap is None else cap
        out = np.empty(cap, dtype=np.uint8)
        w = self.L.ref_encode(_yaml(info), info.version, 0, data.ctypes.data, data.nbytes, out.ctypes.data, cap, int(write_header))
        if w < 0:
            raise RuntimeError(self._err())
        return out[:w].tobytes()

    def header(self, info: cb.EncodingInfo, binary: bool = False) -> bytes:
        out = np.empty(1 << 16, dtype=np.uint8)
        n = self.L.ref_encode_header(_yaml(info), info.version, int(binary), out.ctypes.data, out.nbytes)
        if n < 0:
            raise RuntimeError(self._err())
        return out[:n].tobytes()

    def decode_header(self, blob: bytes):
        buf = C.create_string_buffer(1 << 16)
        ver = C.c_int(0)
        raw = _u8(blob)
        n = self.L.ref_decode_header(raw.ctypes.data, raw.nbytes, buf, len(buf), C.byref(ver))
        if n < 0:
            raise RuntimeError(self._err())
        info = cb.EncodingInfoFromYAML(buf.value.decode())
        info.version = ver.value
        return info, int(n)

    def decode(self, blob: bytes, out: np.ndarray) -> np.ndarray:
        raw = _u8(blob)
        if self.L.ref_decode(raw.ctypes.data, raw.nbytes, out.ctypes.data, out.nbytes) != 0:
            raise RuntimeError(self._err())
        return out

    def decode_payload(self, info: cb.EncodingInfo, payload: bytes, out: np.ndarray) -> np.ndarray:
        raw = _u8(payload)
        if self.L.ref_decode_payload(_yaml(info), info.version, raw.ctypes.data, raw.nbytes, out.ctypes.data, out.nbytes) != 0:
            raise RuntimeError(self._err())
        return out

    def time_encode(self, info, cloud, reps: int, threads: int = 1):
        data = _u8(cloud)
        sz = C.c_size_t(0)
        t = self.L.ref_time_encode(_yaml(info), info.version, data.ctypes.data, data.nbytes, reps, threads, C.byref(sz))
        if t < 0:
            raise RuntimeError(self._err())
        return t, sz.value

    def time_decode(self, blob, reps: int, threads: int = 1):
        raw = _u8(blob)
        t = self.L.ref_time_decode(raw.ctypes.data, raw.nbytes, reps, threads)
        if t < 0:
            raise RuntimeError(self._err())
        return t


def _yaml(info: cb.EncodingInfo) -> bytes:
    return cb.EncodingInfoToYAML(info).encode()


class PortOracle:
    kind = "port"

    def __init__(self):
        build_port()
        L = C.CDLL(PORT_LIB)
        L.orc_last_error.restype = C.c_char_p
        L.orc_header.restype = C.c_size_t
        L.orc_header.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.orc_info_to_yaml.restype = C.c_size_t
        L.orc_info_to_yaml.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
        L.orc_max_compressed_size.restype = C.c_size_t
        L.orc_max_compressed_size.argtypes = [C.c_void_p, C.c_size_t, C.c_int]
        L.orc_encode.restype = C.c_longlong
        L.orc_encode.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]
        L.orc_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.orc_time_encode.restype = C.c_double
        L.orc_time_encode.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int, C.POINTER(C.c_longlong)]
        L.orc_time_decode.restype = C.c_double
        L.orc_time_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]
        L.orc_viz_preprocess.restype = C.c_longlong
        L.orc_viz_preprocess.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        self.L = L

    def _err(self):
        return self.L.orc_last_error().decode("utf-8", "replace")

    def viz_preprocess(self, info, cloud):
        """applyVizLossyPreprocessing restated (ros_msg_utils.cpp:249-341). Returns (EncodingInfo after, surviving bytes)."""
        a = _u8(cloud)
        c = cb._to_c(info)
        out = np.zeros(max(a.size, 1), dtype=np.uint8)
        n = self.L.orc_viz_preprocess(C.byref(c), a.ctypes.data, a.size, out.ctypes.data, out.size)
        if n == -1:  # the reference's no-op conditions
            return info, a
        if n < 0:
            raise RuntimeError("orc_viz_preprocess failed")
        after = cb._from_c(c)
        after.version = info.version
        return after, out[:n * info.point_step]

    def header(self, info) -> bytes:
        c = cb._to_c(info)
        n = self.L.orc_header(C.byref(c), None, 0)
        buf = (C.c_uint8 * n)()
        self.L.orc_header(C.byref(c), buf, n)
        return bytes(buf)

    def max_compressed_size(self, info, points: int, include_header: bool = True) -> int:
        c = cb._to_c(info)
        return int(self.L.orc_max_compressed_size(C.byref(c), points, int(include_header)))

    def encode(self, info, cloud, write_header: bool = True, cap: int = None) -> bytes:
        data = _u8(cloud)
        c = cb._to_c(info)
        n = data.nbytes // info.point_step if info.point_step else 0
        cap = self.max_compressed_size(info, n, True) + 64 if cap is None else cap
        out = np.empty(cap, dtype=np.uint8)
        w = self.L.orc_encode(C.byref(c), data.ctypes.data, data.nbytes, out.ctypes.data, cap, int(write_header))
        if w < 0:
            raise RuntimeError(self._err())
        return out[:w].tobytes()

    def decode_payload(self, info, payload, out: np.ndarray) -> np.ndarray:
        raw = _u8(payload)
        c = cb._to_c(info)
        if self.L.orc_decode(C.byref(c), raw.ctypes.data, raw.nbytes, out.ctypes.data, out.nbytes) != 0:
            raise RuntimeError(self._err())
        return out

    def decode(self, blob, out: np.ndarray) -> np.ndarray:
        info, hdr = cb.DecodeHeader(blob)
        return self.decode_payload(info, bytes(blob)[hdr:], out)

    def time_encode(self, info, cloud, reps: int, threads: int = 1):
        data = _u8(cloud)
        c = cb._to_c(info)
        cap = self.max_compressed_size(info, data.nbytes // info.point_step, True) + 64
        out = np.empty(cap, dtype=np.uint8)
        sz = C.c_longlong(0)
        t = self.L.orc_time_encode(C.byref(c), data.ctypes.data, data.nbytes, out.ctypes.data, cap, reps, C.byref(sz))
        if t < 0:
            raise RuntimeError(self._err())
        return t, sz.value

    def time_decode(self, blob, reps: int, threads: int = 1):
        info, hdr = cb.DecodeHeader(blob)
        raw = _u8(blob)
        c = cb._to_c(info)
        out = np.zeros(info.width * info.height * info.point_step, dtype=np.uint8)
        t = self.L.orc_time_decode(C.byref(c), raw.ctypes.data + hdr, raw.nbytes - hdr, out.ctypes.data, out.nbytes, reps)
        if t < 0:
            raise RuntimeError(self._err())
        return t


def best_oracle():
    """The reference itself when its .so is available, else the C port."""
    return RefOracle() if have_ref() else PortOracle()
