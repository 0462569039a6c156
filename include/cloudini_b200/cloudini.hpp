/*
 * Header-only C++ shim: re-creates the reference's public codec API (namespace Cloudini, cloudini_lib/include/
 * cloudini_lib/cloudini.hpp:31-244 and basic_types.hpp:29-96) on top of the C ABI in include/cloudini_b200.h, so that
 * callers such as cloudini_ros/src/cloudini_publisher_plugin.cpp:53-79 and cloudini_subscriber_plugin.cpp:29-79
 * compile against the B200 library with no source change other than the include path:
 *
 *     #include <cloudini_b200/cloudini.hpp>     // instead of <cloudini_lib/cloudini.hpp>
 *     Cloudini::PointcloudEncoder encoder(info);
 *     encoder.encode(cloud_view, output_vector);
 *
 * Same names, argument meaning and error behaviour (std::runtime_error with the reference's messages). ConstBufferView /
 * BufferView are minimal pointer+size views with the members the call sites use (data(), size(), empty(), trim_front()).
 * Buffers are HOST memory here (that is what the reference API passes); device-resident and batched entry points are
 * available through the C ABI directly.
 */
#pragma once

#include <cstdint>
#include <cstring>
#include <limits>
#include <optional>
#include <stdexcept>
#include <string>
#include <string_view>
#include <vector>

#include "../cloudini_b200.h"

namespace Cloudini {

// ---- basic_types.hpp ------------------------------------------------------------------------------------------------
enum class FieldType : uint8_t {
  UNKNOWN = 0, INT8 = 1, UINT8 = 2, INT16 = 3, UINT16 = 4, INT32 = 5, UINT32 = 6, FLOAT32 = 7, FLOAT64 = 8, INT64 = 9, UINT64 = 10,
};

struct PointField {
  std::string name;
  uint32_t offset = 0;
  FieldType type = FieldType::UNKNOWN;
  std::optional<float> resolution;
  bool operator==(const PointField& o) const { return name == o.name && offset == o.offset && type == o.type && resolution == o.resolution; }
  bool operator!=(const PointField& o) const { return !(*this == o); }
};

constexpr static uint32_t kDecodeButSkipStore = std::numeric_limits<uint32_t>::max();

inline int constexpr SizeOf(const FieldType& type) {
  switch (type) {
    case FieldType::INT8: case FieldType::UINT8: return 1;
    case FieldType::INT16: case FieldType::UINT16: return 2;
    case FieldType::INT32: case FieldType::UINT32: case FieldType::FLOAT32: return 4;
    case FieldType::FLOAT64: case FieldType::INT64: case FieldType::UINT64: return 8;
    default: return 0;
  }
}

// ---- contrib/span.hpp (the subset the codec API exposes) -------------------------------------------------------------
template <typename T>
class Span {
 public:
  Span() = default;
  Span(T* data, size_t size) : data_(data), size_(size) {}
  template <typename Container>
  Span(Container& c) : data_(c.data()), size_(c.size()) {}
  T* data() const { return data_; }
  size_t size() const { return size_; }
  bool empty() const { return size_ == 0; }
  void trim_front(size_t n) {
    if (n > size_) throw std::runtime_error("Span::trim_front: out of range");
    data_ += n;
    size_ -= n;
  }
 private:
  T* data_ = nullptr;
  size_t size_ = 0;
};
using ConstBufferView = Span<const uint8_t>;
using BufferView = Span<uint8_t>;

// ---- cloudini.hpp ----------------------------------------------------------------------------------------------------
enum class EncodingOptions : uint8_t { NONE = 0, LOSSY = 1, LOSSLESS = 2 };
enum class CompressionOption : uint8_t { NONE = 0, LZ4 = 1, ZSTD = 2 };
constexpr const uint8_t kEncodingVersion = CLDN_ENCODING_VERSION;

struct EncodingInfo {
  std::vector<PointField> fields;
  uint32_t width = 0;
  uint32_t height = 1;
  uint32_t point_step = 0;
  EncodingOptions encoding_opt = EncodingOptions::LOSSY;
  std::string encoding_config;
  CompressionOption compression_opt = CompressionOption::ZSTD;
  bool use_threads = true;
  uint8_t version = kEncodingVersion;
};

namespace detail {
inline void check(int rc) {
  if (rc != CLDN_OK) throw std::runtime_error(cldn_b200_last_error());
}
inline cldn_info_t to_c(const EncodingInfo& in) {
  // limits of the C ABI's fixed-size record: refuse instead of silently producing a different header
  if (in.fields.size() > CLDN_MAX_FIELDS) throw std::runtime_error("cloudini_b200: more than " + std::to_string(CLDN_MAX_FIELDS) + " fields are not supported");
  if (in.encoding_config.size() >= sizeof(cldn_info_t{}.encoding_config)) throw std::runtime_error("cloudini_b200: encoding_config is too long for the C ABI record");
  for (const auto& f : in.fields) {
    if (f.name.size() >= CLDN_MAX_NAME) throw std::runtime_error("cloudini_b200: field name '" + f.name + "' is longer than " + std::to_string(CLDN_MAX_NAME - 1) + " characters");
  }
  cldn_info_t c;
  cldn_b200_info_init(&c);
  c.width = in.width; c.height = in.height; c.point_step = in.point_step;
  c.encoding_opt = static_cast<uint8_t>(in.encoding_opt);
  c.compression_opt = static_cast<uint8_t>(in.compression_opt);
  c.version = in.version; c.use_threads = in.use_threads ? 1 : 0;
  c.n_fields = static_cast<uint32_t>(in.fields.size());
  std::strncpy(c.encoding_config, in.encoding_config.c_str(), sizeof(c.encoding_config) - 1);
  for (size_t i = 0; i < in.fields.size(); ++i) {
    std::memset(&c.fields[i], 0, sizeof(cldn_field_t));
    std::strncpy(c.fields[i].name, in.fields[i].name.c_str(), CLDN_MAX_NAME - 1);
    c.fields[i].offset = in.fields[i].offset;
    c.fields[i].type = static_cast<uint8_t>(in.fields[i].type);
    c.fields[i].has_resolution = in.fields[i].resolution.has_value() ? 1 : 0;
    c.fields[i].resolution = in.fields[i].resolution.value_or(0.0f);
  }
  return c;
}
inline EncodingInfo from_c(const cldn_info_t& c) {
  EncodingInfo out;
  out.width = c.width; out.height = c.height; out.point_step = c.point_step;
  out.encoding_opt = static_cast<EncodingOptions>(c.encoding_opt);
  out.compression_opt = static_cast<CompressionOption>(c.compression_opt);
  out.version = c.version; out.use_threads = c.use_threads != 0;
  out.encoding_config = c.encoding_config;
  for (uint32_t i = 0; i < c.n_fields; ++i) {
    PointField f;
    f.name = c.fields[i].name;
    f.offset = c.fields[i].offset;
    f.type = static_cast<FieldType>(c.fields[i].type);
    if (c.fields[i].has_resolution) f.resolution = c.fields[i].resolution;
    out.fields.push_back(std::move(f));
  }
  return out;
}
}  // namespace detail

enum class HeaderEncoding { BINARY, YAML };

inline std::string EncodingInfoToYAML(const EncodingInfo& info) {
  const cldn_info_t c = detail::to_c(info);
  size_t need = 0;
  cldn_b200_info_to_yaml(&c, nullptr, 0, &need);
  std::string out(need, '\0');
  detail::check(cldn_b200_info_to_yaml(&c, out.data(), out.size(), nullptr));
  out.resize(need ? need - 1 : 0);
  return out;
}

inline EncodingInfo EncodingInfoFromYAML(std::string_view yaml) {
  cldn_info_t c;
  detail::check(cldn_b200_info_from_yaml(yaml.data(), yaml.size(), &c));
  return detail::from_c(c);
}

inline void EncodeHeader(const EncodingInfo& header, std::vector<uint8_t>& output, HeaderEncoding encoding = HeaderEncoding::YAML) {
  const cldn_info_t c = detail::to_c(header);
  auto write = encoding == HeaderEncoding::YAML ? cldn_b200_encode_header : cldn_b200_encode_header_binary;
  size_t need = 0;
  write(&c, nullptr, 0, &need);
  output.resize(need);
  detail::check(write(&c, output.data(), output.size(), nullptr));
}

// Advances `input` past the header, like the reference.
inline EncodingInfo DecodeHeader(ConstBufferView& input) {
  cldn_info_t c;
  size_t used = 0;
  detail::check(cldn_b200_decode_header(input.data(), input.size(), &c, &used));
  input.trim_front(used);
  return detail::from_c(c);
}

inline size_t MaxCompressedSize(const EncodingInfo& info, size_t points_count, bool include_header = true) {
  const cldn_info_t c = detail::to_c(info);
  const size_t n = cldn_b200_max_compressed_size(&c, points_count, include_header ? 1 : 0);
  if (n == 0 && (info.point_step == 0 || include_header || points_count > 0)) throw std::runtime_error(cldn_b200_last_error());
  return n;
}

class PointcloudEncoder {
 public:
  explicit PointcloudEncoder(const EncodingInfo& info) : info_(info) {
    const cldn_info_t c = detail::to_c(info);
    detail::check(cldn_b200_encoder_create(&c, -1, nullptr, &handle_));
    const uint8_t* h = nullptr;
    size_t n = 0;
    cldn_b200_encoder_header(handle_, &h, &n);
    header_.assign(h, h + n);
  }
  PointcloudEncoder(const PointcloudEncoder&) = delete;
  PointcloudEncoder& operator=(const PointcloudEncoder&) = delete;
  ~PointcloudEncoder() { cldn_b200_encoder_destroy(handle_); }

  // cloudini.cpp:501-520
  size_t encode(ConstBufferView cloud_data, std::vector<uint8_t>& output) {
    if (info_.point_step == 0) throw std::runtime_error("point_step cannot be 0");
    if (cloud_data.size() % info_.point_step != 0) throw std::runtime_error("Input cloud_data size is not a multiple of point_step");
    output.resize(MaxCompressedSize(info_, cloud_data.size() / info_.point_step, true));
    size_t written = 0;
    detail::check(cldn_b200_encode(handle_, cloud_data.data(), cloud_data.size(), output.data(), output.size(), 1, &written, CLDN_MEM_HOST));
    output.resize(written);
    return written;
  }
  // cloudini.cpp:522-623 — the caller's view is not advanced; returns the bytes written (header included if requested)
  size_t encode(ConstBufferView cloud_data, BufferView& output, bool write_header) {
    size_t written = 0;
    detail::check(cldn_b200_encode(handle_, cloud_data.data(), cloud_data.size(), output.data(), output.size(), write_header ? 1 : 0,
                                   &written, CLDN_MEM_HOST));
    return written;
  }
  const EncodingInfo& getEncodingInfo() const { return info_; }
  const std::vector<uint8_t>& getHeader() const { return header_; }

 private:
  EncodingInfo info_;
  std::vector<uint8_t> header_;
  cldn_encoder_t* handle_ = nullptr;
};

class PointcloudDecoder {
 public:
  PointcloudDecoder() { detail::check(cldn_b200_decoder_create(-1, nullptr, &handle_)); }
  PointcloudDecoder(const PointcloudDecoder&) = delete;
  PointcloudDecoder& operator=(const PointcloudDecoder&) = delete;
  ~PointcloudDecoder() { cldn_b200_decoder_destroy(handle_); }

  // cloudini.cpp:635-668: compressed_data must NOT contain the header
  void decode(const EncodingInfo& info, ConstBufferView compressed_data, BufferView output) {
    const cldn_info_t c = detail::to_c(info);
    detail::check(cldn_b200_decode(handle_, &c, compressed_data.data(), compressed_data.size(), output.data(), output.size(), CLDN_MEM_HOST));
  }
  void decode(const EncodingInfo& info, ConstBufferView compressed_data, std::vector<uint8_t>& output) {
    output.resize(static_cast<size_t>(info.width) * info.height * info.point_step);
    decode(info, compressed_data, BufferView(output.data(), output.size()));
  }

 private:
  cldn_decoder_t* handle_ = nullptr;
};

}  // namespace Cloudini
