// Exercises include/cloudini_b200/ros_msg_utils.hpp the way the reference's converter uses cloudini_ros
// (cloudini_lib/tools/src/mcap_converter.cpp:184-204): parse a DDS PointCloud2 message, apply the resolution profile,
// optionally the viz preprocessing, toEncodingInfo, convertPointCloud2ToCompressedCloud; then convert back.
// usage: ros_shim_convert <in.msg> <out.compressed> <out.restored> <default_resolution> <viz 0|1>
// Compiled on CPU (syntax / link check of the shim), run on the GPU box.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iterator>
#include <vector>

#include "cloudini_b200/ros_msg_utils.hpp"

static std::vector<uint8_t> slurp(const char* path) {
  std::ifstream f(path, std::ios::binary);
  return std::vector<uint8_t>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}
static void dump(const char* path, const std::vector<uint8_t>& v) {
  std::ofstream f(path, std::ios::binary);
  f.write(reinterpret_cast<const char*>(v.data()), static_cast<std::streamsize>(v.size()));
}

int main(int argc, char** argv) {
  if (argc != 6) { std::puts("usage: ros_shim_convert in out_compressed out_restored default_resolution viz"); return 2; }
  try {
    const std::vector<uint8_t> msg = slurp(argv[1]);
    Cloudini::ConstBufferView raw_dds_msg(msg.data(), msg.size());
    auto pc_info = cloudini_ros::getDeserializedPointCloudMessage(raw_dds_msg);
    cloudini_ros::applyResolutionProfile(cloudini_ros::ResolutionProfile{}, pc_info.fields, static_cast<float>(std::atof(argv[4])));
    if (std::atoi(argv[5])) cloudini_ros::applyVizLossyPreprocessing(pc_info);
    auto copy = pc_info;  // the copy must re-bind its data view to its own owned_data (ros_msg_utils.hpp:155-159,
                          // cloudini_lib/test/test_ros_msg.cpp:146-156 RosPointCloud2CopyRebindsOwnedDataView)
    if (!pc_info.owned_data.empty() && (copy.data.data() != copy.owned_data.data() || copy.data.size() != copy.owned_data.size())) {
      std::puts("ros_shim_convert: copy did not re-bind its data view");
      return 3;
    }
    auto encoding_info = cloudini_ros::toEncodingInfo(copy);
    encoding_info.compression_opt = Cloudini::CompressionOption::NONE;
    encoding_info.use_threads = false;
    std::vector<uint8_t> compressed, restored;
    cloudini_ros::convertPointCloud2ToCompressedCloud(copy, encoding_info, compressed);
    dump(argv[2], compressed);
    auto comp_info = cloudini_ros::getDeserializedPointCloudMessage(Cloudini::ConstBufferView(compressed.data(), compressed.size()));
    cloudini_ros::convertCompressedCloudToPointCloud2(comp_info, restored);
    dump(argv[3], restored);
    std::printf("ros_shim_convert: ok %zu -> %zu -> %zu bytes, %u points\n", msg.size(), compressed.size(), restored.size(), copy.width);
    return 0;
  } catch (const std::exception& e) {
    std::printf("ros_shim_convert: exception: %s\n", e.what());
    return 1;
  }
}
