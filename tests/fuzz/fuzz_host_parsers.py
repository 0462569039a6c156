import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT)
import cloudini_b200 as cb
from cloudini_b200 import synth, ros
rng = np.random.default_rng(int(sys.argv[1]))
info, cloud = synth.cloud_c3(50, seed=1)
yaml_hdr = cb.EncodeHeader(info)
bin_hdr = cb.EncodeHeader(info, binary=True)
n_ok = n_err = 0
for base in (yaml_hdr, bin_hdr):
    for t in range(int(sys.argv[2])):
        b = bytearray(base + bytes(64))
        for _ in range(int(rng.integers(1, 8))):
            i = int(rng.integers(0, len(base))); b[i] = int(rng.integers(0, 256))
        if rng.random() < 0.3: b = b[:int(rng.integers(0, len(b)))]
        try:
            di, h = cb.DecodeHeader(bytes(b)); n_ok += 1
            try: cb.MaxCompressedSize(di, 1000)
            except RuntimeError: pass
            try: cb.EncodingInfoToYAML(di)
            except RuntimeError: pass
        except RuntimeError: n_err += 1
# random garbage through the DDS parser and the YAML reader
for t in range(int(sys.argv[2])):
    g = rng.integers(0, 256, int(rng.integers(0, 400)), dtype=np.uint8).tobytes()
    try: ros.getDeserializedPointCloudMessage(bytes([0, 1, 0, 0]) + g)
    except RuntimeError: pass
    try: cb.EncodingInfoFromYAML(g.decode("latin-1"))
    except RuntimeError: pass
print("decoded ok", n_ok, "rejected", n_err)
