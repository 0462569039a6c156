/*
 * cloudini_b200 — C ABI of the callers either side of the codec:
 *   N3  applyVizLossyPreprocessing  (cloudini_lib/src/ros_msg_utils.cpp:249-341): NaN/inf drop + order-preserving voxel
 *       de-duplication as sm_100a kernels directly in front of the encode launch;
 *   N2  the DDS / ROS 2 envelope    (cloudini_lib/src/ros_msg_utils.cpp:54-238): CDR (de)serialisation of
 *       sensor_msgs/msg/PointCloud2 and point_cloud_interfaces/msg/CompressedPointCloud2 around the codec.
 * Same conventions as cloudini_b200.h (status codes, cldn_b200_last_error(), caller-owned buffers, no CPU fallback for
 * anything that touches point data: the envelope functions parse / write the small CDR header on the host and hand the
 * point payload to the GPU codec).
 */
#ifndef CLOUDINI_B200_ROS_H_
#define CLOUDINI_B200_ROS_H_

#include "cloudini_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- N3: visualisation-oriented lossy preprocessing ------------------------------------------------------------- */
typedef struct cldn_preproc cldn_preproc_t; /* owns the voxel hash table, tile status words and staging buffers */

int cldn_b200_preproc_create(int device, void* stream, cldn_preproc_t** out);
void cldn_b200_preproc_destroy(cldn_preproc_t* pp);

/* applyVizLossyPreprocessing(RosPointCloud2&) (ros_msg_utils.cpp:249-341, contract ros_msg_utils.hpp:198-221).
 *  - the geometry triple is detected structurally: fields[0..2] FLOAT32 with the same resolution at offsets b, b+4, b+8
 *    (names are never read); without it, with a non-positive / non-finite resolution or with an empty cloud the call is
 *    a no-op exactly like the reference: *applied = 0, `info` and `out` untouched, *kept_points = input point count;
 *  - points whose x, y or z is NaN / +-inf are dropped; of the points that quantise (lround(v * (1/res)), 21 bits per
 *    axis, packVoxelKey21 ros_msg_utils.cpp:42-49) to the same voxel only the FIRST survives; survivors keep their order
 *    and all `point_step` bytes;
 *  - on success `info` is updated like pc_info: width = kept, height = 1, every FLOAT64 field without a resolution gets
 *    1e-6 (so the planner routes it through the lossy coder).
 * `cloud` / `out` live in `mem` memory; `out_capacity` >= cloud_bytes is always enough. CLDN_MEM_DEVICE: the kernels run
 * on the handle's stream and the call waits for the 4-byte survivor count. */
int cldn_b200_viz_lossy_preprocess(cldn_preproc_t* pp, cldn_info_t* info, const void* cloud, size_t cloud_bytes,
                                   void* out, size_t out_capacity, size_t* kept_points, int* applied, int mem);

/* ---- N2: DDS envelope ------------------------------------------------------------------------------------------- */
/* Parsed view of a CDR-serialised sensor_msgs/msg/PointCloud2 or point_cloud_interfaces/msg/CompressedPointCloud2
 * (cloudini_ros::RosPointCloud2, ros_msg_utils.hpp:32-148). `frame_id` and `data` are views (into the message that was
 * parsed, or into whatever buffers the caller sets them to before re-serialising). */
typedef struct cldn_ros_msg_t {
  uint8_t cdr_header[4];     /* nanocdr::CdrHeader as found on the wire: {0, encapsulation, 0, 0} */
  int32_t stamp_sec;
  uint32_t stamp_nsec;
  const char* frame_id;      /* header.frame_id, NOT NUL-terminated */
  uint32_t frame_id_len;     /* without the trailing NUL */
  uint32_t height, width;
  uint32_t n_fields;
  cldn_field_t fields[CLDN_MAX_FIELDS]; /* has_resolution = 0 after parsing (the message carries none) */
  uint8_t is_bigendian, is_dense;
  uint32_t point_step, row_step;
  const uint8_t* data;       /* PointCloud2::data / CompressedPointCloud2::compressed_data */
  size_t data_bytes;
} cldn_ros_msg_t;

/* getDeserializedPointCloudMessage (ros_msg_utils.cpp:54-95). Host memory. Little-endian PLAIN_CDR like the reference's
 * callers produce; CLDN_ERR_CORRUPT_DATA where nanocdr throws ("not enough data", bad header). */
int cldn_b200_ros_parse(const void* dds_msg, size_t msg_bytes, cldn_ros_msg_t* out);

/* toEncodingInfo (ros_msg_utils.cpp:122-131): LOSSY + ZSTD defaults, fields copied. */
int cldn_b200_ros_to_encoding_info(const cldn_ros_msg_t* msg, cldn_info_t* info);

/* applyResolutionProfile (ros_msg_utils.cpp:217-238) on a field list (msg->fields / &msg->n_fields — the reference's
 * converter applies it to the message's own fields, tools/src/mcap_converter.cpp:189, so removed fields also leave the
 * re-written message header): a profile entry of 0 removes the field, other entries set the field's resolution,
 * FLOAT32 fields without an entry get *default_resolution when it is non-NULL. */
int cldn_b200_ros_apply_resolution_profile(cldn_field_t* fields, uint32_t* n_fields, const char* const* names,
                                           const float* resolutions, size_t n_profile, const float* default_resolution);

/* convertPointCloud2ToCompressedCloud (ros_msg_utils.cpp:167-213): PointCloud2 message -> CompressedPointCloud2
 * message (same CDR header, width/height/fields/point_step, compressed_data = Cloudini blob with header, is_dense,
 * format "cloudini"). The point payload is encoded on the GPU through `enc` (created from `encoding_info`; its
 * compression_opt may be LZ4/ZSTD: host-pointer API). The payload is msg->data / msg->data_bytes — after
 * cldn_b200_viz_lossy_preprocess point them (and msg->width / height) at the preprocessed cloud, like pc_info.data.
 * Query the worst-case size with out == NULL (*written receives it). */
int cldn_b200_ros_compress_msg(cldn_encoder_t* enc, const cldn_ros_msg_t* msg, void* out, size_t out_capacity,
                               size_t* written);

/* convertCompressedCloudToPointCloud2 (ros_msg_utils.cpp:134-165): CompressedPointCloud2 message -> PointCloud2 message,
 * the blob decoded on the GPU through `dec` straight into the output message. out == NULL queries the exact size. */
int cldn_b200_ros_decompress_msg(cldn_decoder_t* dec, const cldn_ros_msg_t* msg, void* out, size_t out_capacity,
                                 size_t* written);

/* The converter's per-message step (tools/src/mcap_converter.cpp:184-204: getDeserializedPointCloudMessage ->
 * applyResolutionProfile -> [applyVizLossyPreprocessing] -> toEncodingInfo + options -> convertPointCloud2ToCompressedCloud)
 * as ONE call, host message in, host message out. Same bytes as the five calls; what differs is where the payload lives:
 * it is uploaded once, stays in HBM between the preprocessing and the encode kernels (only the 4-byte survivor count and
 * the 8-byte blob size come back), and encoder / preprocessor / staging buffers come from a per-thread pool keyed by the
 * layout instead of being created per message. compression_opt != NONE: stage 1 as above, then the host libraries.
 * out == NULL queries a worst-case size. */
int cldn_b200_ros_convert_msg(const void* dds_msg, size_t msg_bytes, const char* const* names, const float* resolutions,
                              size_t n_profile, const float* default_resolution, int viz, int encoding_opt,
                              int compression_opt, int version, void* out, size_t out_capacity, size_t* written);

/* Handles of the calling thread's pool (created on first use, owned by the library, never to be destroyed by the caller;
 * valid for this thread only). The reference constructs encoder / decoder objects per message, which costs it nothing;
 * a GPU handle owns streams, device buffers and pinned memory, so the per-message callers (the shims of
 * convertPointCloud2ToCompressedCloud & co.) take theirs from here. The encoder is re-dimensioned to info->width / height. */
int cldn_b200_pool_encoder(const cldn_info_t* info, cldn_encoder_t** out);
int cldn_b200_pool_decoder(cldn_decoder_t** out);
int cldn_b200_pool_preproc(cldn_preproc_t** out);

/* ---- the rest of the reference's own C ABI (include/cloudini_lib/wasm_functions.h:30-93, src/wasm_functions.cpp), the
 * functions that take DDS messages. Same shape and "return 0 on any failure" convention; every output pointer is
 * followed by its capacity (the WASM module trusts the caller's allocation instead). Host memory. ------------------- */
/* cldn_GetHeaderAsYAML (wasm_functions.cpp:24-44): YAML text of the blob's header (not NUL-terminated); returns its size. */
uint32_t cldn_b200_GetHeaderAsYAML(const void* encoded_data, uint32_t encoded_data_size, char* output_yaml, uint32_t capacity);
/* cldn_GetHeaderAsYAMLFromDDS (:46-56): same, the blob is the compressed_data of a CompressedPointCloud2 message. */
uint32_t cldn_b200_GetHeaderAsYAMLFromDDS(const void* raw_dds_msg, uint32_t dds_msg_size, char* output_yaml, uint32_t capacity);
/* cldn_ComputeCompressedSize (:58-93): full GPU encode of a PointCloud2 message (every FLOAT32 field at `resolution`,
 * toEncodingInfo defaults: LOSSY + ZSTD), only the size is returned. */
uint32_t cldn_b200_ComputeCompressedSize(const void* dds_msg, uint32_t dds_msg_size, float resolution);
/* cldn_GetDecompressedSize (:95-106): height * width * point_step of the message; nothing is decoded. */
uint32_t cldn_b200_GetDecompressedSize(const void* encoded_dds_msg, uint32_t encoded_dds_size);
/* cldn_ConvertCompressedMsgToPointCloud2Msg (:108-125): CompressedPointCloud2 message -> PointCloud2 message. */
uint32_t cldn_b200_ConvertCompressedMsgToPointCloud2Msg(const void* compressed_msg, uint32_t msg_size, void* output_msg,
                                                        uint32_t capacity);
/* cldn_DecodeCompressedMessage (:127-148): CompressedPointCloud2 message -> raw point data. */
uint32_t cldn_b200_DecodeCompressedMessage(const void* compressed_msg, uint32_t msg_size, void* output_data, uint32_t capacity);
/* cldn_EncodePointcloudMessage (:178-226): PointCloud2 message -> Cloudini blob with header (every FLOAT32 field at
 * `resolution`); 0 when data size != width * height * point_step or the blob does not fit `capacity`. */
uint32_t cldn_b200_EncodePointcloudMessage(const void* pointcloud_msg, uint32_t msg_size, float resolution, void* output_data,
                                           uint32_t capacity);

#ifdef __cplusplus
}
#endif
#endif /* CLOUDINI_B200_ROS_H_ */
