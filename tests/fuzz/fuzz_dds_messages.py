import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT)
import cloudini_b200 as cb
from cloudini_b200 import ros, synth
from cloudini_b200 import FieldType as FT
from oracle.client import RefOracle
ref = RefOracle()
rng = np.random.default_rng(int(sys.argv[1]))
XYZI = [("x",0,FT.FLOAT32),("y",4,FT.FLOAT32),("z",8,FT.FLOAT32),("intensity",12,FT.FLOAT32)]
good = synth.pointcloud2_msg(XYZI, 16, synth.cloud_viz(400, seed=3)[1])
def convert(msg, viz):
    pc = ros.getDeserializedPointCloudMessage(msg)
    ros.applyResolutionProfile({}, pc.fields, 0.001)
    if viz: ros.applyVizLossyPreprocessing(pc)
    info = ros.toEncodingInfo(pc)
    info.compression_opt, info.use_threads = cb.CompressionOption.NONE, False
    return ros.convertPointCloud2ToCompressedCloud(pc, info)
mism = 0; n = 0; both_ok = 0
for t in range(int(sys.argv[2])):
    b = bytearray(good)
    for _ in range(int(rng.integers(1, 3))): b[int(rng.integers(4, 150))] = int(rng.integers(0, 256))
    b = bytes(b); viz = bool(rng.integers(0, 2))
    n += 1
    # the reference reads / writes out of bounds on forged field offsets (it segfaults): only hand it messages it survives
    safe = True
    try:
        pc0 = ros.getDeserializedPointCloudMessage(b)
        for f in pc0.fields:
            t = int(f.type)
            if t < 1 or t > 10 or f.offset + cb.SizeOf(cb.FieldType(t)) > pc0.point_step: safe = False
        if pc0.point_step == 0 or pc0.data.size % max(pc0.point_step, 1): safe = False
    except RuntimeError: pass
    if not safe:
        try: convert(b, viz)
        except RuntimeError: pass
        continue
    try: want = ref.ros_compress(b, {}, 0.001, viz, 1, 0, 5)
    except RuntimeError as e: want = None; werr = str(e)
    try: got = convert(b, viz)
    except RuntimeError as e: got = None; gerr = str(e)
    if want is not None and got is not None and want == got: both_ok += 1
    if (want is None) != (got is None) or (want is not None and want != got):
        # tolerated: inputs the reference handles with undefined behaviour / that this build refuses loudly by design
        msgs = ("does not fit a point", "at most", "field name longer", "Unsupported field type", "too many fields", "embedded NUL")
        if got is None and any(m in gerr for m in msgs): continue
        mism += 1
        if mism <= 12: print("MISMATCH ref", want is not None, "ours", got is not None, "viz", viz, (werr if want is None else "")[:70], "|", (gerr if got is None else "")[:90])
print("n", n, "identical", both_ok, "unexplained mismatches", mism)
