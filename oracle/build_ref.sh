#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY (oracle). Compiles the UNMODIFIED reference codec sources where they lie under
# /root/reference (never copied into this repo) plus oracle/ref_wrapper.cpp into oracle/_ref/libcloudini_ref.so.
# The reference's own CMake build cannot run offline (CPM downloads lz4/zstd/mcap; gtest/PCL absent), so the
# codec translation units (+ ros_msg_utils.cpp, the DDS envelope / viz preprocessing around them) are compiled directly. -msse4.1 is mandatory: without it cast_vector4f_to_vector4i
# (cloudini_lib/include/cloudini_lib/intrinsics.hpp:288-300) silently switches from round-to-nearest-even to std::round.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${CLOUDINI_REFERENCE:-/root/reference}/cloudini_lib"
OUT="$HERE/_ref"
if [ ! -d "$REF/src" ]; then
  echo "build_ref.sh: reference tree not found at $REF (expected on the GPU box: the prebuilt .so travels)" >&2
  exit 3
fi
mkdir -p "$OUT"
# -O3 -DNDEBUG: the reference's own default build type is Release (cloudini_lib/CMakeLists.txt:5-8)
g++ -std=c++20 -O3 -DNDEBUG -msse4.1 -fPIC -shared \
    -I"$REF/include" -I"$REF/src" -I"$HERE/shim" \
    "$REF"/src/{chunk_writer,cloudini,codec_common,field_encoder,field_decoder,v4_codec,v5_codec,ros_msg_utils}.cpp \
    "$HERE/ref_wrapper.cpp" \
    -l:liblz4.so.1 -l:libzstd.so.1 -lpthread \
    -o "$OUT/libcloudini_ref.so"
echo "built $OUT/libcloudini_ref.so"
