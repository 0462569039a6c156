/*
 * Header-only C++ shim: re-creates the reference's `namespace cloudini_ros` (cloudini_lib/include/cloudini_lib/
 * ros_msg_utils.hpp:32-221) on top of the C ABI in include/cloudini_b200_ros.h, so that the callers of the DDS envelope
 * — cloudini_ros/src/topic_converter.cpp:160-185, cloudini_ros/src/conversion_utils.cpp:70-95,
 * cloudini_lib/tools/src/mcap_converter.cpp:184-204 — compile against the B200 library by switching the include:
 *
 *     #include <cloudini_b200/ros_msg_utils.hpp>    // instead of <cloudini_lib/ros_msg_utils.hpp>
 *     auto pc_info = cloudini_ros::getDeserializedPointCloudMessage(raw_dds_msg);
 *     cloudini_ros::applyResolutionProfile(profile, pc_info.fields, default_resolution);
 *     cloudini_ros::applyVizLossyPreprocessing(pc_info);                       // sm_100a kernels
 *     auto encoding_info = cloudini_ros::toEncodingInfo(pc_info);
 *     cloudini_ros::convertPointCloud2ToCompressedCloud(pc_info, encoding_info, compressed_dds_msg);   // GPU codec
 *
 * Same names, argument meaning and error behaviour (std::runtime_error). The CDR header is parsed / written by host code
 * inside the library; point data only ever goes through the GPU kernels.
 */
#pragma once
#include <map>
#include <optional>
#include <string>
#include <utility>
#include <vector>

#include "../cloudini_b200_ros.h"
#include "cloudini.hpp"

namespace cloudini_ros {

struct RosHeader {  // ros_msg_utils.hpp:26-30
  int32_t stamp_sec = 0;
  uint32_t stamp_nsec = 0;
  std::string frame_id;
};

// Mirrors cloudini_ros::RosPointCloud2 (ros_msg_utils.hpp:32-148). `data` views either the DDS buffer it was parsed from
// or `owned_data` (after applyVizLossyPreprocessing); `cdr_header` is the 4-byte encapsulation header as found on the wire.
struct RosPointCloud2 {
  uint8_t cdr_header[4] = {0, 1, 0, 0};
  RosHeader ros_header;
  uint32_t height = 1;
  uint32_t width = 0;
  std::vector<Cloudini::PointField> fields;
  uint32_t point_step = 0;
  uint32_t row_step = 0;
  bool is_bigendian = false;
  Cloudini::ConstBufferView data;
  bool is_dense = true;
  std::vector<uint8_t> owned_data;

  RosPointCloud2() = default;
  RosPointCloud2(const RosPointCloud2& o) { *this = o; }
  RosPointCloud2& operator=(const RosPointCloud2& o) {
    if (this == &o) return *this;
    const bool owned_view = !o.owned_data.empty() && o.data.data() == o.owned_data.data();
    std::memcpy(cdr_header, o.cdr_header, 4);
    ros_header = o.ros_header; height = o.height; width = o.width; fields = o.fields; point_step = o.point_step;
    row_step = o.row_step; is_bigendian = o.is_bigendian; data = o.data; is_dense = o.is_dense; owned_data = o.owned_data;
    if (owned_view) data = Cloudini::ConstBufferView(owned_data.data(), owned_data.size());  // ros_msg_utils.hpp:155-159
    return *this;
  }
  RosPointCloud2(RosPointCloud2&& o) noexcept { *this = std::move(o); }
  RosPointCloud2& operator=(RosPointCloud2&& o) noexcept {
    if (this == &o) return *this;
    const bool owned_view = !o.owned_data.empty() && o.data.data() == o.owned_data.data();
    std::memcpy(cdr_header, o.cdr_header, 4);
    ros_header = std::move(o.ros_header); height = o.height; width = o.width; fields = std::move(o.fields);
    point_step = o.point_step; row_step = o.row_step; is_bigendian = o.is_bigendian; data = o.data; is_dense = o.is_dense;
    owned_data = std::move(o.owned_data);
    if (owned_view) data = Cloudini::ConstBufferView(owned_data.data(), owned_data.size());
    return *this;
  }
};

using ResolutionProfile = std::map<std::string, float>;  // ros_msg_utils.hpp:162-164

namespace detail {
inline void fields_to_c(const std::vector<Cloudini::PointField>& in, cldn_field_t* out, uint32_t* n) {
  if (in.size() > CLDN_MAX_FIELDS) throw std::runtime_error("too many fields");
  *n = static_cast<uint32_t>(in.size());
  for (size_t i = 0; i < in.size(); ++i) {
    std::memset(&out[i], 0, sizeof(cldn_field_t));
    std::strncpy(out[i].name, in[i].name.c_str(), CLDN_MAX_NAME - 1);
    out[i].offset = in[i].offset;
    out[i].type = static_cast<uint8_t>(in[i].type);
    out[i].has_resolution = in[i].resolution.has_value() ? 1 : 0;
    out[i].resolution = in[i].resolution.value_or(0.0f);
  }
}
inline void fields_from_c(const cldn_field_t* in, uint32_t n, std::vector<Cloudini::PointField>& out) {
  out.clear();
  for (uint32_t i = 0; i < n; ++i) {
    Cloudini::PointField f;
    f.name = in[i].name;
    f.offset = in[i].offset;
    f.type = static_cast<Cloudini::FieldType>(in[i].type);
    if (in[i].has_resolution) f.resolution = in[i].resolution;
    out.push_back(std::move(f));
  }
}
// The C view of a RosPointCloud2 (pointers into pc: valid while pc is alive and unchanged).
inline cldn_ros_msg_t to_c(const RosPointCloud2& pc) {
  cldn_ros_msg_t m;
  std::memset(&m, 0, sizeof(m));
  std::memcpy(m.cdr_header, pc.cdr_header, 4);
  m.stamp_sec = pc.ros_header.stamp_sec;
  m.stamp_nsec = pc.ros_header.stamp_nsec;
  m.frame_id = pc.ros_header.frame_id.data();
  m.frame_id_len = static_cast<uint32_t>(pc.ros_header.frame_id.size());
  m.height = pc.height; m.width = pc.width; m.point_step = pc.point_step; m.row_step = pc.row_step;
  m.is_bigendian = pc.is_bigendian ? 1 : 0; m.is_dense = pc.is_dense ? 1 : 0;
  fields_to_c(pc.fields, m.fields, &m.n_fields);
  m.data = pc.data.data();
  m.data_bytes = pc.data.size();
  return m;
}
}  // namespace detail

// ros_msg_utils.cpp:90-95
inline RosPointCloud2 getDeserializedPointCloudMessage(Cloudini::ConstBufferView pc2_dds_msg) {
  cldn_ros_msg_t m;
  Cloudini::detail::check(cldn_b200_ros_parse(pc2_dds_msg.data(), pc2_dds_msg.size(), &m));
  RosPointCloud2 pc;
  std::memcpy(pc.cdr_header, m.cdr_header, 4);
  pc.ros_header.stamp_sec = m.stamp_sec;
  pc.ros_header.stamp_nsec = m.stamp_nsec;
  pc.ros_header.frame_id.assign(m.frame_id, m.frame_id_len);
  pc.height = m.height; pc.width = m.width; pc.point_step = m.point_step; pc.row_step = m.row_step;
  // a non-canonical CDR bool (byte > 1) is normalised to true here; the reference memcpy's the byte into its bool and
  // back (undefined, in practice the byte survives) — the C ABI (cldn_ros_msg_t::is_dense) and the Python mirror carry it
  pc.is_dense = m.is_dense != 0;
  detail::fields_from_c(m.fields, m.n_fields, pc.fields);
  pc.data = Cloudini::ConstBufferView(m.data, m.data_bytes);
  return pc;
}

// ros_msg_utils.cpp:217-238
inline void applyResolutionProfile(const ResolutionProfile& profile, std::vector<Cloudini::PointField>& field,
                                   std::optional<float> default_resolution = std::nullopt) {
  cldn_field_t c[CLDN_MAX_FIELDS];
  uint32_t n = 0;
  detail::fields_to_c(field, c, &n);
  std::vector<const char*> names;
  std::vector<float> res;
  for (const auto& kv : profile) { names.push_back(kv.first.c_str()); res.push_back(kv.second); }
  const float dflt = default_resolution.value_or(0.0f);
  Cloudini::detail::check(cldn_b200_ros_apply_resolution_profile(c, &n, names.data(), res.data(), names.size(),
                                                                 default_resolution ? &dflt : nullptr));
  detail::fields_from_c(c, n, field);
}

// ros_msg_utils.cpp:122-131
inline Cloudini::EncodingInfo toEncodingInfo(const RosPointCloud2& pc_info) {
  Cloudini::EncodingInfo info;
  info.height = pc_info.height;
  info.width = pc_info.width;
  info.point_step = pc_info.point_step;
  info.encoding_opt = Cloudini::EncodingOptions::LOSSY;
  info.compression_opt = Cloudini::CompressionOption::ZSTD;
  info.fields = pc_info.fields;
  return info;
}

// ros_msg_utils.cpp:167-213 — the point payload is encoded on the GPU
inline void convertPointCloud2ToCompressedCloud(const RosPointCloud2& pc_info, const Cloudini::EncodingInfo& encoding_info,
                                                std::vector<uint8_t>& compressed_dds_msg) {
  const cldn_info_t c = Cloudini::detail::to_c(encoding_info);
  cldn_encoder_t* enc = nullptr;  // from the library's per-thread pool (the reference builds one per message, :198)
  Cloudini::detail::check(cldn_b200_pool_encoder(&c, &enc));
  const cldn_ros_msg_t m = detail::to_c(pc_info);
  size_t need = 0, written = 0;
  int rc = cldn_b200_ros_compress_msg(enc, &m, nullptr, 0, &need);
  if (rc == CLDN_OK) {
    compressed_dds_msg.resize(need);
    rc = cldn_b200_ros_compress_msg(enc, &m, compressed_dds_msg.data(), compressed_dds_msg.size(), &written);
  }
  Cloudini::detail::check(rc);
  compressed_dds_msg.resize(written);
}

// ros_msg_utils.cpp:134-165 — the blob is decoded on the GPU straight into the output message
inline void convertCompressedCloudToPointCloud2(const RosPointCloud2& pc_info, std::vector<uint8_t>& pc2_dds_msg) {
  cldn_decoder_t* dec = nullptr;
  Cloudini::detail::check(cldn_b200_pool_decoder(&dec));
  const cldn_ros_msg_t m = detail::to_c(pc_info);
  size_t need = 0, written = 0;
  int rc = cldn_b200_ros_decompress_msg(dec, &m, nullptr, 0, &need);
  if (rc == CLDN_OK) {
    pc2_dds_msg.resize(need);
    rc = cldn_b200_ros_decompress_msg(dec, &m, pc2_dds_msg.data(), pc2_dds_msg.size(), &written);
  }
  Cloudini::detail::check(rc);
  pc2_dds_msg.resize(written);
}

// ros_msg_utils.cpp:249-341 — NaN drop + voxel de-duplication on the GPU; pc_info is rewritten in place
inline void applyVizLossyPreprocessing(RosPointCloud2& pc_info) {
  Cloudini::EncodingInfo info = toEncodingInfo(pc_info);
  cldn_info_t c = Cloudini::detail::to_c(info);
  cldn_preproc_t* pp = nullptr;
  Cloudini::detail::check(cldn_b200_pool_preproc(&pp));
  std::vector<uint8_t> out(pc_info.data.size());
  size_t kept = 0;
  int applied = 0;
  const int rc = cldn_b200_viz_lossy_preprocess(pp, &c, pc_info.data.data(), pc_info.data.size(), out.data(), out.size(), &kept,
                                                &applied, CLDN_MEM_HOST);
  Cloudini::detail::check(rc);
  if (!applied) return;  // the reference's early returns leave pc_info untouched
  out.resize(kept * pc_info.point_step);
  pc_info.owned_data = std::move(out);
  pc_info.data = Cloudini::ConstBufferView(pc_info.owned_data.data(), pc_info.owned_data.size());
  pc_info.width = static_cast<uint32_t>(kept);
  pc_info.height = 1;
  pc_info.row_step = pc_info.point_step * pc_info.width;
  pc_info.fields = Cloudini::detail::from_c(c).fields;
}

}  // namespace cloudini_ros
