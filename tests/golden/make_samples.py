"""Writes tests/golden/samples_v1.npz: full-length copies of the reference's two sample DATA files (cloudini_lib/samples:
the point payload of lidar.pcd and the CDR bytes of dds_message.bin), so that the SURVEY 8(c) known-answer hashes of
test_reference_sample_files are checked on the GPU box too, where /root/reference does not exist.
Run once where the reference tree is present:  python tests/golden/make_samples.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BASE = os.path.join(os.environ.get("CLOUDINI_REFERENCE", "/root/reference"), "cloudini_lib", "samples")


def main():
    raw = open(os.path.join(BASE, "lidar.pcd"), "rb").read()
    n = int(raw[raw.index(b"POINTS ") + 7:raw.index(b"\n", raw.index(b"POINTS "))])
    body = raw[raw.index(b"DATA binary\n") + 12:]
    lidar = np.frombuffer(body[:n * 16], dtype=np.uint8)
    dds = np.frombuffer(open(os.path.join(BASE, "dds_message.bin"), "rb").read(), dtype=np.uint8)
    out = os.path.join(HERE, "samples_v1.npz")
    np.savez_compressed(out, lidar_pcd_points=lidar, lidar_pcd_n=np.array([n]), dds_message_bin=dds)
    print(out, os.path.getsize(out), "bytes; lidar points", n, "dds bytes", dds.size)


if __name__ == "__main__":
    sys.exit(main())
