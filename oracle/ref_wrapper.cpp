// TEST INFRASTRUCTURE ONLY (oracle). Not part of the product path.
//
// extern "C" façade over the UNMODIFIED reference library, compiled in place from
// /root/reference/cloudini_lib by oracle/build_ref.sh into oracle/_ref/libcloudini_ref.so.
// It lets the tests (ctypes) and bench.py's cpu_baseline / --impl reference leg call
//   Cloudini::PointcloudEncoder::encode      (cloudini_lib/src/cloudini.cpp:501-623)
//   Cloudini::PointcloudDecoder::decode      (cloudini_lib/src/cloudini.cpp:635-668)
//   Cloudini::MaxCompressedSize              (cloudini_lib/src/cloudini.cpp:249-292)
//   Cloudini::DecodeHeader / EncodingInfoToYAML (cloudini_lib/src/cloudini.cpp:165-190,353-428)
// The configuration crosses the boundary as the reference's own YAML header text
// (the same convention as cldn_EncodePointcloudData, wasm_functions.h:88-93).
//
// and, for SURVEY.md §8(f) N2 / N3 (the callers either side of the codec),
//   cloudini_ros::getDeserializedPointCloudMessage / applyResolutionProfile / toEncodingInfo /
//   convertPointCloud2ToCompressedCloud / convertCompressedCloudToPointCloud2 / applyVizLossyPreprocessing
//                                            (cloudini_lib/src/ros_msg_utils.cpp:54-341)
// No reference source is copied: this file only #includes the reference's public headers.
#include <chrono>
#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "cloudini_lib/cloudini.hpp"
#include "cloudini_lib/ros_msg_utils.hpp"

namespace {
thread_local std::string g_err;

Cloudini::EncodingInfo infoFromYaml(const char* yaml, int version_override, int use_threads) {
  Cloudini::EncodingInfo info = Cloudini::EncodingInfoFromYAML(yaml);
  // EncodingInfoFromYAML parses "version" as a single character (see the comment at
  // cloudini.cpp:389-392), so the caller passes the numeric version explicitly.
  info.version = static_cast<uint8_t>(version_override);
  info.use_threads = use_threads != 0;
  return info;
}
}  // namespace

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }

// Returns total bytes written (header included when write_header != 0); -1 on exception.
long long ref_encode(const char* yaml, int version, int use_threads, const uint8_t* cloud,
                     size_t cloud_bytes, uint8_t* out, size_t out_capacity, int write_header) {
  try {
    Cloudini::EncodingInfo info = infoFromYaml(yaml, version, use_threads);
    Cloudini::PointcloudEncoder enc(info);
    Cloudini::ConstBufferView in(cloud, cloud_bytes);
    Cloudini::BufferView view(out, out_capacity);
    return static_cast<long long>(enc.encode(in, view, write_header != 0));
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}

// Cloudini::EncodeHeader(info, out, binary ? BINARY : YAML). Returns the header size; -1 on exception / small buffer.
long long ref_encode_header(const char* yaml, int version, int binary, uint8_t* out, size_t out_capacity) {
  try {
    Cloudini::EncodingInfo info = infoFromYaml(yaml, version, 0);
    std::vector<uint8_t> h;
    Cloudini::EncodeHeader(info, h, binary ? Cloudini::HeaderEncoding::BINARY : Cloudini::HeaderEncoding::YAML);
    if (h.size() > out_capacity) { g_err = "header buffer too small"; return -1; }
    memcpy(out, h.data(), h.size());
    return static_cast<long long>(h.size());
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}

// Sizing helper. Returns 0 on exception.
size_t ref_max_compressed_size(const char* yaml, int version, size_t points, int include_header) {
  try {
    Cloudini::EncodingInfo info = infoFromYaml(yaml, version, 0);
    return Cloudini::MaxCompressedSize(info, points, include_header != 0);
  } catch (const std::exception& e) {
    g_err = e.what();
    return 0;
  }
}

// Parses the header of a full blob. Writes the YAML text (NUL-terminated) of the decoded info to
// yaml_out, the numeric version to *version_out, and returns the header length in bytes (-1 on error).
long long ref_decode_header(const uint8_t* blob, size_t blob_bytes, char* yaml_out, size_t yaml_capacity,
                            int* version_out) {
  try {
    Cloudini::ConstBufferView view(blob, blob_bytes);
    Cloudini::EncodingInfo info = Cloudini::DecodeHeader(view);
    const std::string yaml = Cloudini::EncodingInfoToYAML(info);
    if (yaml.size() + 1 > yaml_capacity) {
      g_err = "yaml_out too small";
      return -1;
    }
    std::memcpy(yaml_out, yaml.c_str(), yaml.size() + 1);
    *version_out = info.version;
    return static_cast<long long>(blob_bytes - view.size());
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}

// Decodes a full blob (header + payload) into out (caller pre-fills it; the reference only
// writes declared field bytes). Returns 0 on success, -1 on exception.
int ref_decode(const uint8_t* blob, size_t blob_bytes, uint8_t* out, size_t out_capacity) {
  try {
    Cloudini::ConstBufferView view(blob, blob_bytes);
    Cloudini::EncodingInfo info = Cloudini::DecodeHeader(view);
    Cloudini::PointcloudDecoder dec;
    dec.decode(info, view, Cloudini::BufferView(out, out_capacity));
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}

// Decode a header-less payload with an explicit configuration.
int ref_decode_payload(const char* yaml, int version, const uint8_t* payload, size_t payload_bytes,
                       uint8_t* out, size_t out_capacity) {
  try {
    Cloudini::EncodingInfo info = infoFromYaml(yaml, version, 0);
    Cloudini::PointcloudDecoder dec;
    dec.decode(info, Cloudini::ConstBufferView(payload, payload_bytes), Cloudini::BufferView(out, out_capacity));
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}

// ---- timing helpers for bench.py (cpu_baseline / --impl reference) -------------------------
// Times exactly the region mcap_codec_benchmark.cpp:454-457 / 509-512 times: encoder.encode(in, view, true)
// and decoder.decode(...), with a pre-sized output. `threads` independent encoder instances each
// process `reps` frames (one instance per thread: the reference's stage 1 is single-threaded).
// Returns elapsed seconds (wall clock over all threads), or -1 on error.
double ref_time_encode(const char* yaml, int version, const uint8_t* cloud, size_t cloud_bytes, int reps,
                       int threads, size_t* encoded_bytes_out) {
  try {
    Cloudini::EncodingInfo info = infoFromYaml(yaml, version, 0);
    const size_t points = cloud_bytes / info.point_step;
    const size_t cap = Cloudini::MaxCompressedSize(info, points, true);
    std::vector<std::vector<uint8_t>> outs(threads, std::vector<uint8_t>(cap));
    std::vector<size_t> sizes(threads, 0);
    auto work = [&](int t) {
      Cloudini::PointcloudEncoder enc(info);
      for (int r = 0; r < reps; ++r) {
        Cloudini::BufferView view(outs[t].data(), outs[t].size());
        sizes[t] = enc.encode(Cloudini::ConstBufferView(cloud, cloud_bytes), view, true);
      }
    };
    const auto t0 = std::chrono::steady_clock::now();
    if (threads == 1) {
      work(0);
    } else {
      std::vector<std::thread> pool;
      for (int t = 0; t < threads; ++t) pool.emplace_back(work, t);
      for (auto& th : pool) th.join();
    }
    const auto t1 = std::chrono::steady_clock::now();
    if (encoded_bytes_out) *encoded_bytes_out = sizes[0];
    return std::chrono::duration<double>(t1 - t0).count();
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1.0;
  }
}

double ref_time_decode(const uint8_t* blob, size_t blob_bytes, int reps, int threads) {
  try {
    Cloudini::ConstBufferView probe(blob, blob_bytes);
    Cloudini::EncodingInfo info = Cloudini::DecodeHeader(probe);
    const size_t header_bytes = blob_bytes - probe.size();
    const size_t out_bytes = static_cast<size_t>(info.width) * info.height * info.point_step;
    std::vector<std::vector<uint8_t>> outs(threads, std::vector<uint8_t>(out_bytes));
    auto work = [&](int t) {
      Cloudini::PointcloudDecoder dec;
      for (int r = 0; r < reps; ++r) {
        dec.decode(info, Cloudini::ConstBufferView(blob + header_bytes, blob_bytes - header_bytes),
                   Cloudini::BufferView(outs[t].data(), outs[t].size()));
      }
    };
    const auto t0 = std::chrono::steady_clock::now();
    if (threads == 1) {
      work(0);
    } else {
      std::vector<std::thread> pool;
      for (int t = 0; t < threads; ++t) pool.emplace_back(work, t);
      for (auto& th : pool) th.join();
    }
    const auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1.0;
  }
}

// ---- bench.py --impl reference: the whole reference arm without any product code in the process -------------------------
// The EncodingInfo of BASELINE configs[1] (x y z intensity FLOAT32 at 1 mm, LOSSY, stage 1 only) is built HERE, so the
// reference process never needs the product library to render a YAML header. `threads` workers, each with a PRIVATE copy
// of one of the `n_clouds` distinct input clouds (worker t takes cloud t % n_clouds), its own encoder / decoder instance
// and its own output buffers, run `passes` encode passes, then `passes` decode passes of the blob they produced.
// Timed regions as in mcap_codec_benchmark.cpp:454-457 / 509-512 (encode / decode call only, buffers pre-sized).
// Returns 0 and fills enc_s / dec_s (wall clock over all workers), -1 on exception.
int ref_bench_xyzi(const uint8_t* clouds, size_t cloud_bytes, int n_clouds, float resolution, int threads, int passes,
                   double* enc_s, double* dec_s, size_t* blob_bytes_out) {
  try {
    Cloudini::EncodingInfo info;
    const char* names[4] = {"x", "y", "z", "intensity"};
    for (int k = 0; k < 4; ++k) {
      Cloudini::PointField f;
      f.name = names[k];
      f.offset = 4u * k;
      f.type = Cloudini::FieldType::FLOAT32;
      f.resolution = resolution;
      info.fields.push_back(f);
    }
    info.point_step = 16;
    info.width = static_cast<uint32_t>(cloud_bytes / 16);
    info.height = 1;
    info.encoding_opt = Cloudini::EncodingOptions::LOSSY;
    info.compression_opt = Cloudini::CompressionOption::NONE;
    info.use_threads = false;
    const size_t cap = Cloudini::MaxCompressedSize(info, info.width, true);
    std::vector<std::vector<uint8_t>> ins(threads), blobs(threads), outs(threads);
    std::vector<size_t> sizes(threads, 0);
    for (int t = 0; t < threads; ++t) {
      const uint8_t* src = clouds + static_cast<size_t>(t % n_clouds) * cloud_bytes;
      ins[t].assign(src, src + cloud_bytes);
      blobs[t].resize(cap);
      outs[t].resize(cloud_bytes);
    }
    auto run = [&](auto&& work) {
      const auto t0 = std::chrono::steady_clock::now();
      std::vector<std::thread> pool;
      for (int t = 1; t < threads; ++t) pool.emplace_back(work, t);
      work(0);
      for (auto& th : pool) th.join();
      return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    };
    std::string first_error;
    *enc_s = run([&](int t) {
      try {
        Cloudini::PointcloudEncoder enc(info);
        for (int r = 0; r < passes; ++r) {
          Cloudini::BufferView view(blobs[t].data(), blobs[t].size());
          sizes[t] = enc.encode(Cloudini::ConstBufferView(ins[t].data(), ins[t].size()), view, true);
        }
      } catch (const std::exception& e) { if (t == 0) first_error = e.what(); }
    });
    *dec_s = run([&](int t) {
      try {
        Cloudini::ConstBufferView probe(blobs[t].data(), sizes[t]);
        Cloudini::EncodingInfo hinfo = Cloudini::DecodeHeader(probe);
        const size_t header_bytes = sizes[t] - probe.size();
        Cloudini::PointcloudDecoder dec;
        for (int r = 0; r < passes; ++r) {
          dec.decode(hinfo, Cloudini::ConstBufferView(blobs[t].data() + header_bytes, sizes[t] - header_bytes),
                     Cloudini::BufferView(outs[t].data(), outs[t].size()));
        }
      } catch (const std::exception& e) { if (t == 0) first_error = e.what(); }
    });
    if (!first_error.empty()) { g_err = first_error; return -1; }
    if (blob_bytes_out) *blob_bytes_out = sizes[0];
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}

// ---- N3: applyVizLossyPreprocessing on a raw cloud described by the YAML header text ------------------------------------
// Returns the number of surviving points (-1 on exception); the rewritten cloud goes to `out`, the updated EncodingInfo
// (width, height, FLOAT64 resolutions) to yaml_out.
long long ref_viz_preprocess(const char* yaml, int version, const uint8_t* cloud, size_t cloud_bytes, uint8_t* out,
                             size_t out_capacity, char* yaml_out, size_t yaml_capacity) {
  try {
    Cloudini::EncodingInfo info = infoFromYaml(yaml, version, 0);
    cloudini_ros::RosPointCloud2 pc;
    pc.height = info.height;
    pc.width = info.width;
    pc.fields = info.fields;
    pc.point_step = info.point_step;
    pc.row_step = info.point_step * info.width;
    pc.data = Cloudini::ConstBufferView(cloud, cloud_bytes);
    cloudini_ros::applyVizLossyPreprocessing(pc);
    if (pc.data.size() > out_capacity) { g_err = "output too small"; return -1; }
    if (pc.data.size()) memcpy(out, pc.data.data(), pc.data.size());
    Cloudini::EncodingInfo after = cloudini_ros::toEncodingInfo(pc);
    after.encoding_opt = info.encoding_opt;
    after.compression_opt = info.compression_opt;
    after.version = info.version;
    const std::string y = Cloudini::EncodingInfoToYAML(after);
    if (y.size() + 1 > yaml_capacity) { g_err = "yaml buffer too small"; return -1; }
    memcpy(yaml_out, y.c_str(), y.size() + 1);
    return pc.point_step ? static_cast<long long>(pc.data.size() / pc.point_step) : 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}

// ---- N2: DDS envelope ------------------------------------------------------------------------------------------------------
static cloudini_ros::ResolutionProfile parseProfile(const char* text) {  // "name=res;name=res"
  cloudini_ros::ResolutionProfile p;
  std::string t = text ? text : "";
  size_t pos = 0;
  while (pos < t.size()) {
    size_t end = t.find(';', pos);
    if (end == std::string::npos) end = t.size();
    const std::string item = t.substr(pos, end - pos);
    const size_t eq = item.find('=');
    if (eq != std::string::npos) p[item.substr(0, eq)] = std::stof(item.substr(eq + 1));
    pos = end + 1;
  }
  return p;
}

// Canonical one-line-per-item description of a parsed message (compared verbatim with the product's parser).
long long ref_ros_describe(const uint8_t* msg, size_t msg_bytes, char* out, size_t capacity) {
  try {
    auto pc = cloudini_ros::getDeserializedPointCloudMessage(Cloudini::ConstBufferView(msg, msg_bytes));
    std::string t;
    t += "stamp " + std::to_string(pc.ros_header.stamp_sec) + " " + std::to_string(pc.ros_header.stamp_nsec) + "\n";
    t += "frame_id " + pc.ros_header.frame_id + "\n";
    t += "height " + std::to_string(pc.height) + " width " + std::to_string(pc.width) + "\n";
    for (const auto& f : pc.fields) {
      t += "field " + f.name + " " + std::to_string(f.offset) + " " + std::to_string(static_cast<int>(f.type)) + "\n";
    }
    t += "point_step " + std::to_string(pc.point_step) + " row_step " + std::to_string(pc.row_step) + "\n";
    t += "data " + std::to_string(pc.data.data() - msg) + " " + std::to_string(pc.data.size()) + "\n";
    t += "is_dense " + std::to_string(pc.is_dense ? 1 : 0) + "\n";
    if (t.size() + 1 > capacity) { g_err = "text buffer too small"; return -1; }
    memcpy(out, t.c_str(), t.size() + 1);
    return static_cast<long long>(t.size());
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}

// The per-message step of the reference's converter (tools/src/mcap_converter.cpp:184-204): parse, apply the resolution
// profile, optionally the viz preprocessing, toEncodingInfo, convertPointCloud2ToCompressedCloud.
// default_resolution < 0 = none. Returns the size of the CompressedPointCloud2 message, -1 on exception.
long long ref_ros_compress(const uint8_t* msg, size_t msg_bytes, const char* profile, float default_resolution, int viz,
                           int encoding_opt, int compression_opt, int version, uint8_t* out, size_t out_capacity) {
  try {
    auto pc = cloudini_ros::getDeserializedPointCloudMessage(Cloudini::ConstBufferView(msg, msg_bytes));
    std::optional<float> def;
    if (default_resolution >= 0.0f) def = default_resolution;
    cloudini_ros::applyResolutionProfile(parseProfile(profile), pc.fields, def);
    if (viz) cloudini_ros::applyVizLossyPreprocessing(pc);
    auto info = cloudini_ros::toEncodingInfo(pc);
    info.encoding_opt = static_cast<Cloudini::EncodingOptions>(encoding_opt);
    info.compression_opt = static_cast<Cloudini::CompressionOption>(compression_opt);
    info.version = static_cast<uint8_t>(version);
    info.use_threads = false;
    std::vector<uint8_t> result;
    cloudini_ros::convertPointCloud2ToCompressedCloud(pc, info, result);
    if (result.size() > out_capacity) { g_err = "output too small"; return -1; }
    memcpy(out, result.data(), result.size());
    return static_cast<long long>(result.size());
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}

long long ref_ros_decompress(const uint8_t* msg, size_t msg_bytes, uint8_t* out, size_t out_capacity) {
  try {
    auto pc = cloudini_ros::getDeserializedPointCloudMessage(Cloudini::ConstBufferView(msg, msg_bytes));
    std::vector<uint8_t> result;
    cloudini_ros::convertCompressedCloudToPointCloud2(pc, result);
    if (result.size() > out_capacity) { g_err = "output too small"; return -1; }
    memcpy(out, result.data(), result.size());
    return static_cast<long long>(result.size());
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}

}  // extern "C"
