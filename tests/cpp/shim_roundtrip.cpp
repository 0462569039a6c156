// Exercises the header-only C++ shim the way the reference's own tests use its API
// (cloudini_lib/test/test_field_encoders.cpp:695-769, test_header.cpp:142-163): build an EncodingInfo, encode a cloud
// into a std::vector, DecodeHeader + decode, compare within resolution. Compiled on CPU (syntax/link check), run on GPU.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "cloudini_b200/cloudini.hpp"

using namespace Cloudini;

struct PointXYZI { float x, y, z, intensity; };

int main() {
  const size_t n = 4133;  // same size as the reference's v5==v4 test
  std::vector<PointXYZI> pts(n);
  for (size_t i = 0; i < n; ++i) {
    pts[i] = {0.001f * float(i), 1.0f + 0.002f * float(i % 97), -3.0f + std::sin(0.01f * float(i)), float(i % 256)};
  }
  pts[7].y = std::nanf("");
  EncodingInfo info;
  info.width = n; info.height = 1; info.point_step = sizeof(PointXYZI);
  info.encoding_opt = EncodingOptions::LOSSY;
  info.compression_opt = CompressionOption::NONE;
  info.use_threads = false;
  info.fields = {{"x", 0, FieldType::FLOAT32, 0.001f}, {"y", 4, FieldType::FLOAT32, 0.001f},
                 {"z", 8, FieldType::FLOAT32, 0.001f}, {"intensity", 12, FieldType::FLOAT32, 0.001f}};
  try {
    PointcloudEncoder encoder(info);
    std::vector<uint8_t> blob;
    ConstBufferView cloud(reinterpret_cast<const uint8_t*>(pts.data()), pts.size() * sizeof(PointXYZI));
    const size_t size = encoder.encode(cloud, blob);
    if (size != blob.size() || std::memcmp(blob.data(), "CLOUDINI_V05\n", 13) != 0) { std::puts("bad blob"); return 2; }

    // the BufferView overload must produce the same bytes
    std::vector<uint8_t> blob2(MaxCompressedSize(info, n, true));
    BufferView view(blob2.data(), blob2.size());
    const size_t size2 = encoder.encode(cloud, view, true);
    if (size2 != size || std::memcmp(blob.data(), blob2.data(), size) != 0) { std::puts("overloads differ"); return 3; }

    ConstBufferView in(blob.data(), blob.size());
    EncodingInfo decoded_info = DecodeHeader(in);
    if (decoded_info.width != n || decoded_info.fields.size() != 4 || decoded_info.version != 5) { std::puts("bad header"); return 4; }
    PointcloudDecoder decoder;
    std::vector<uint8_t> out;
    decoder.decode(decoded_info, in, out);
    const PointXYZI* got = reinterpret_cast<const PointXYZI*>(out.data());
    for (size_t i = 0; i < n; ++i) {
      if (i == 7) { if (!std::isnan(got[i].y)) { std::puts("NaN lost"); return 5; } continue; }
      if (std::fabs(got[i].x - pts[i].x) > 0.0011f || std::fabs(got[i].y - pts[i].y) > 0.0011f ||
          std::fabs(got[i].z - pts[i].z) > 0.0011f || std::fabs(got[i].intensity - pts[i].intensity) > 0.0011f) {
        std::printf("point %zu out of tolerance\n", i);
        return 6;
      }
    }
    // error behaviour: payload that still carries the header -> std::runtime_error (cloudini.cpp:640-643)
    bool threw = false;
    try { decoder.decode(decoded_info, ConstBufferView(blob.data(), blob.size()), out); } catch (const std::runtime_error&) { threw = true; }
    if (!threw) { std::puts("missing exception"); return 7; }
  } catch (const std::exception& e) {
    std::printf("exception: %s\n", e.what());
    return 1;
  }
  std::puts("shim_roundtrip: ok");
  return 0;
}
