// Host-side (no CUDA) part of libcloudini_b200: configuration <-> YAML text, blob header, worst-case sizing and
// the planner that turns an EncodingInfo into the flat op table the kernels consume.
// Behavioural contract = the reference functions cited at each definition (paths relative to cloudini_lib/).
#include <errno.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "cldn_plan.h"

namespace cldn {

static thread_local std::string g_last_error;

void set_error(const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
}

const char* last_error_cstr() { return g_last_error.c_str(); }

// basic_types.hpp:73-96 SizeOf
int size_of_type(uint8_t type) {
  switch (type) {
    case CLDN_INT8: case CLDN_UINT8: return 1;
    case CLDN_INT16: case CLDN_UINT16: return 2;
    case CLDN_INT32: case CLDN_UINT32: case CLDN_FLOAT32: return 4;
    case CLDN_FLOAT64: case CLDN_INT64: case CLDN_UINT64: return 8;
    default: return 0;
  }
}

static const char* type_name(uint8_t t) {  // cloudini.cpp:48-76
  static const char* names[] = {"UNKNOWN", "INT8", "UINT8", "INT16", "UINT16", "INT32",
                                "UINT32",  "FLOAT32", "FLOAT64", "INT64", "UINT64"};
  return t <= CLDN_UINT64 ? names[t] : "UNKNOWN";
}
static const char* enc_name(uint8_t e) {  // cloudini.cpp:78-89
  return e == CLDN_ENC_NONE ? "NONE" : e == CLDN_ENC_LOSSY ? "LOSSY" : e == CLDN_ENC_LOSSLESS ? "LOSSLESS" : "UNKNOWN";
}
static const char* comp_name(uint8_t c) {  // cloudini.cpp:91-102
  return c == CLDN_COMP_NONE ? "NONE" : c == CLDN_COMP_LZ4 ? "LZ4" : c == CLDN_COMP_ZSTD ? "ZSTD" : "UNKNOWN";
}

static bool parse_int_in_range(const std::string& s, int lo, int hi, int* out) {
  char* end = nullptr;
  long v = strtol(s.c_str(), &end, 10);
  if (end == s.c_str() || v < lo || v > hi) return false;
  *out = static_cast<int>(v);
  return true;
}
// cloudini.cpp:104-163: names first, then a decimal enum value.
static bool type_from_string(const std::string& s, uint8_t* out) {
  for (uint8_t t = CLDN_INT8; t <= CLDN_UINT64; ++t) {
    if (s == type_name(t)) { *out = t; return true; }
  }
  int v;
  if (parse_int_in_range(s, CLDN_UNKNOWN, CLDN_UINT64, &v)) { *out = static_cast<uint8_t>(v); return true; }
  return false;
}
static bool enc_from_string(const std::string& s, uint8_t* out) {
  if (s == "NONE") { *out = CLDN_ENC_NONE; return true; }
  if (s == "LOSSY") { *out = CLDN_ENC_LOSSY; return true; }
  if (s == "LOSSLESS") { *out = CLDN_ENC_LOSSLESS; return true; }
  int v;
  if (parse_int_in_range(s, 0, 2, &v)) { *out = static_cast<uint8_t>(v); return true; }
  return false;
}
static bool comp_from_string(const std::string& s, uint8_t* out) {
  if (s == "NONE") { *out = CLDN_COMP_NONE; return true; }
  if (s == "LZ4") { *out = CLDN_COMP_LZ4; return true; }
  if (s == "ZSTD") { *out = CLDN_COMP_ZSTD; return true; }
  int v;
  if (parse_int_in_range(s, 0, 2, &v)) { *out = static_cast<uint8_t>(v); return true; }
  return false;
}

// EncodingInfoToYAML, cloudini.cpp:165-190. Resolution is printed like `ostream << float` (== "%g").
std::string info_to_yaml(const cldn_info_t& info) {
  std::string y;
  char line[256];
  snprintf(line, sizeof(line), "version: %d\n", static_cast<int>(info.version)); y += line;
  snprintf(line, sizeof(line), "width: %u\n", info.width); y += line;
  snprintf(line, sizeof(line), "height: %u\n", info.height); y += line;
  snprintf(line, sizeof(line), "point_step: %u\n", info.point_step); y += line;
  y += "encoding_opt: "; y += enc_name(info.encoding_opt); y += "\n";
  y += "compression_opt: "; y += comp_name(info.compression_opt); y += "\n";
  if (info.encoding_config[0] != '\0') {
    y += "encoding_config: "; y += info.encoding_config; y += "\n";
  }
  y += "fields:\n";
  for (uint32_t i = 0; i < info.n_fields && i < CLDN_MAX_FIELDS; ++i) {
    const cldn_field_t& f = info.fields[i];
    y += "  - name: "; y += f.name; y += "\n";
    snprintf(line, sizeof(line), "    offset: %u\n", f.offset); y += line;
    y += "    type: "; y += type_name(f.type); y += "\n";
    if (f.has_resolution) {
      snprintf(line, sizeof(line), "    resolution: %g\n", static_cast<double>(f.resolution)); y += line;
    } else {
      y += "    resolution: null\n";
    }
  }
  return y;
}

static std::string trim(const std::string& s) {
  size_t a = 0, b = s.size();
  while (a < b && (s[a] == ' ' || s[a] == '\t' || s[a] == '\r')) ++a;
  while (b > a && (s[b - 1] == ' ' || s[b - 1] == '\t' || s[b - 1] == '\r')) --b;
  return s.substr(a, b - a);
}
static std::string unquote(const std::string& s) {
  if (s.size() >= 2 && ((s.front() == '"' && s.back() == '"') || (s.front() == '\'' && s.back() == '\''))) {
    return s.substr(1, s.size() - 2);
  }
  return s;
}

// `iss >> uint32_t` of the reference's parseScalar (yaml_parser.hpp:99-108): leading blanks, then at least one digit;
// whatever follows the digits is ignored ("30O" reads as 30), no digit at all is an error.
static bool parse_u32_prefix(const std::string& text, uint32_t* out) {
  const char* p = text.c_str();
  while (*p == ' ' || *p == '\t') ++p;
  if (*p == '+') ++p;
  if (*p < '0' || *p > '9') return false;
  unsigned long long v = 0;
  while (*p >= '0' && *p <= '9') { v = v * 10 + static_cast<unsigned>(*p - '0'); if (v > 0xFFFFFFFFull) return false; ++p; }
  *out = static_cast<uint32_t>(v);
  return true;
}

// EncodingInfoFromYAML (cloudini.cpp:192-230). Like the reference's parser (yaml_parser.hpp) this understands the
// exact shape EncodingInfoToYAML emits: top-level "key: value" lines and a "fields:" sequence of 4-key maps.
int info_from_yaml(const char* yaml, size_t len, cldn_info_t* info) {
  cldn_b200_info_init(info);
  info->n_fields = 0;
  bool in_fields = false;
  bool have[4] = {false, false, false, false};  // width height point_step version
  bool have_opts[2] = {false, false};            // encoding_opt compression_opt (as<> of a missing node throws, cloudini.cpp:203-204)
  uint8_t field_keys[CLDN_MAX_FIELDS] = {};      // per field: name | offset | type | resolution seen (all four are read, :216-221)
  int cur = -1;
  size_t pos = 0;
  auto fail = [&](const char* what, const std::string& v) {
    set_error("EncodingInfoFromYAML: %s '%s'", what, v.c_str());
    return CLDN_ERR_BAD_HEADER;
  };
  while (pos < len) {
    size_t eol = pos;
    while (eol < len && yaml[eol] != '\n') ++eol;
    std::string raw(yaml + pos, eol - pos);
    pos = eol + 1;
    std::string line = trim(raw);
    if (line.empty() || line[0] == '#') continue;
    bool new_item = false;
    if (line.rfind("- ", 0) == 0) {
      new_item = true;
      line = trim(line.substr(2));
    }
    size_t colon = line.find(':');
    // Inside the "fields:" sequence only the exact shape EncodingInfoToYAML writes is accepted. A damaged line there
    // ("  Q name: y", a duplicated or unknown key, text after "fields:") would otherwise be skipped and the keys that follow
    // would land in the wrong field: the blob would decode, silently, with a layout nobody wrote.
    if (colon == std::string::npos) {
      if (in_fields) return fail("malformed line in the field list:", line);
      continue;
    }
    const std::string key = trim(line.substr(0, colon));
    const std::string val = unquote(trim(line.substr(colon + 1)));
    if (!in_fields) {
      if (key == "fields") {
        if (!val.empty() || new_item) return fail("unexpected text after 'fields:'", val);
        in_fields = true;
        continue;
      }
      if (key == "version") {
        // the reference reads this scalar into a uint8_t, i.e. as ONE CHARACTER (yaml_parser.hpp:99-108), and DecodeHeader
        // then overrides it with the two digits of the magic (cloudini.cpp:389-392): any non-empty text is accepted;
        // a number is used as such (that is what EncodingInfoToYAML writes), anything else leaves the default
        int v;
        if (val.empty()) return fail("Failed to convert scalar:", val);
        if (parse_int_in_range(val, 0, 255, &v)) info->version = static_cast<uint8_t>(v);
        have[3] = true;
      }
      else if (key == "width") { if (!parse_u32_prefix(val, &info->width)) return fail("Failed to convert scalar:", val); have[0] = true; }
      else if (key == "height") { if (!parse_u32_prefix(val, &info->height)) return fail("Failed to convert scalar:", val); have[1] = true; }
      else if (key == "point_step") { if (!parse_u32_prefix(val, &info->point_step)) return fail("Failed to convert scalar:", val); have[2] = true; }
      else if (key == "encoding_opt") { if (!enc_from_string(val, &info->encoding_opt)) return fail("Invalid EncodingOptions string:", val); have_opts[0] = true; }
      else if (key == "compression_opt") { if (!comp_from_string(val, &info->compression_opt)) return fail("Invalid CompressionOption string:", val); have_opts[1] = true; }
      else if (key == "encoding_config") { snprintf(info->encoding_config, sizeof(info->encoding_config), "%s", val.c_str()); }
      continue;
    }
    if (new_item) {
      if (info->n_fields >= CLDN_MAX_FIELDS) { set_error("too many fields (max %d)", CLDN_MAX_FIELDS); return CLDN_ERR_UNSUPPORTED; }
      cur = static_cast<int>(info->n_fields++);
      memset(&info->fields[cur], 0, sizeof(cldn_field_t));
    }
    if (cur < 0) return fail("field key before the first list item:", key);
    cldn_field_t& f = info->fields[cur];
    const uint8_t bit = key == "name" ? 1 : key == "offset" ? 2 : key == "type" ? 4 : key == "resolution" ? 8 : 0;
    if (bit == 0) return fail("unexpected key in the field list:", key);
    if (field_keys[cur] & bit) return fail("duplicate key in a field:", key);
    if (new_item != (bit == 1)) return fail("a field starts with '- name:' and nothing else does:", key);
    if (bit != 1 && val.find(':') != std::string::npos) return fail("a second ':' in a field scalar:", val);  // YAML would see a nested mapping
    if (key == "name") {
      if (val.empty()) return fail("empty field name (Node is not a string)", val);
      snprintf(f.name, sizeof(f.name), "%s", val.c_str());
      field_keys[cur] |= 1;
    }
    else if (key == "offset") { if (!parse_u32_prefix(val, &f.offset)) return fail("Failed to convert scalar:", val); field_keys[cur] |= 2; }
    else if (key == "type") { if (!type_from_string(val, &f.type)) return fail("Invalid FieldType string:", val); field_keys[cur] |= 4; }
    else if (key == "resolution") {
      field_keys[cur] |= 8;
      if (val != "null") {  // cloudini.cpp:220-223: std::stof of the text
        char* end = nullptr;
        errno = 0;
        f.resolution = strtof(val.c_str(), &end);
        if (end == val.c_str()) return fail("bad resolution", val);
        if (errno == ERANGE) return fail("stof: resolution out of range", val);  // std::stof throws std::out_of_range (overflow and subnormal results)
        f.has_resolution = 1;
      }
    }
  }
  if (!have[0] || !have[1] || !have[2] || !have[3] || !have_opts[0] || !have_opts[1]) {
    set_error("EncodingInfoFromYAML: missing version/width/height/point_step/encoding_opt/compression_opt (Node is not a string)");
    return CLDN_ERR_BAD_HEADER;
  }
  for (uint32_t i = 0; i < info->n_fields; ++i) {
    if (field_keys[i] != 15) {
      set_error("EncodingInfoFromYAML: field %u lacks name / offset / type / resolution (Node is not a string)", i);
      return CLDN_ERR_BAD_HEADER;
    }
  }
  return CLDN_OK;
}

// ---- sizing --------------------------------------------------------------------------------------------------
// MaxSerializedFieldSize, codec_common.cpp:29-59
size_t max_serialized_field_size(const cldn_field_t& f, uint8_t encoding_opt, bool* ok) {
  *ok = true;
  switch (f.type) {
    case CLDN_INT16: case CLDN_UINT16: case CLDN_INT32: case CLDN_UINT32: case CLDN_INT64: case CLDN_UINT64:
      return 10;
    case CLDN_FLOAT32:
      return (encoding_opt == CLDN_ENC_LOSSY && f.has_resolution) ? 10 : 7;
    case CLDN_FLOAT64:
      return (encoding_opt == CLDN_ENC_LOSSY && f.has_resolution) ? 10 : 11;
    case CLDN_INT8: case CLDN_UINT8:
      return 1;
    default:
      *ok = false;
      return 0;
  }
}

// LeadingLossyFloatFieldCount, codec_common.cpp:69-82
size_t leading_lossy_float_count(const cldn_info_t& info) {
  if (info.encoding_opt != CLDN_ENC_LOSSY) return 0;
  size_t n = 0;
  for (uint32_t i = 0; i < info.n_fields; ++i) {
    if (info.fields[i].type != CLDN_FLOAT32 || !info.fields[i].has_resolution) break;
    ++n;
  }
  return (n == 3 || n == 4) ? n : 0;
}

static bool is_adaptive_int(uint8_t t) {  // isV5AdaptiveIntType, v5_codec.cpp:83-95
  return t == CLDN_INT16 || t == CLDN_UINT16 || t == CLDN_INT32 || t == CLDN_UINT32 || t == CLDN_INT64 || t == CLDN_UINT64;
}

// UsesV5Codec, v5_codec.cpp:883-892
bool uses_v5_codec(const cldn_info_t& info) {
  if (info.version < 5 || info.encoding_opt != CLDN_ENC_LOSSY) return false;
  for (uint32_t i = static_cast<uint32_t>(leading_lossy_float_count(info)); i < info.n_fields; ++i) {
    if (is_adaptive_int(info.fields[i].type)) return true;
  }
  return false;
}

static size_t lz4_bound(size_t n) { return n + n / 255 + 16; }  // LZ4_COMPRESSBOUND (lz4 v1.9/v1.10)
static size_t zstd_bound(size_t n) {                            // ZSTD_COMPRESSBOUND (zstd v1.5.x)
  return n + (n >> 8) + ((n < (128u << 10)) ? (((128u << 10) - n) >> 11) : 0);
}

// MaxCompressedSize, cloudini.cpp:249-292
size_t max_compressed_size(const cldn_info_t& info, size_t points, bool include_header, bool* ok) {
  *ok = false;
  if (info.point_step == 0) { set_error("point_step cannot be 0"); return 0; }
  size_t per_point = 0;
  for (uint32_t i = 0; i < info.n_fields; ++i) {
    bool fok;
    per_point += max_serialized_field_size(info.fields[i], info.encoding_opt, &fok);
    if (!fok) {
      set_error("Unsupported field type '%s' (type=%d) in MaxSerializedFieldSize", info.fields[i].name, info.fields[i].type);
      return 0;
    }
  }
  size_t total = include_header ? (10 + 2 + 1 + info_to_yaml(info).size() + 1) : 0;
  const bool v5 = uses_v5_codec(info);
  size_t left = points;
  const size_t chunks = points / kChunkPoints + ((points % kChunkPoints) ? 1 : 0);
  for (size_t c = 0; c < chunks; ++c) {
    const size_t n = std::min<size_t>(left, kChunkPoints);
    left -= n;
    size_t chunk_in = n * per_point;
    if (v5) chunk_in += info.n_fields * 32u + 1024u;
    total += 4;
    switch (info.compression_opt) {
      case CLDN_COMP_NONE: total += chunk_in; break;
      case CLDN_COMP_LZ4:
        if (chunk_in > 0x7FFFFFFFu) { set_error("Chunk size too large for LZ4"); return 0; }
        total += lz4_bound(chunk_in);
        break;
      case CLDN_COMP_ZSTD: total += zstd_bound(chunk_in); break;
      default: set_error("Unsupported compression option in MaxCompressedSize"); return 0;
    }
  }
  *ok = true;
  return total;
}

// ---- planner ---------------------------------------------------------------------------------------------------
static void init_plan(const cldn_info_t& info, Plan* p) {
  memset(p, 0, sizeof(*p));
  p->point_step = info.point_step;
  p->supported = 1;
}

static void push_floatn(const cldn_info_t& info, size_t n, Plan* p) {
  RegOp& op = p->ops[p->n_ops++];
  memset(&op, 0, sizeof(op));
  op.kind = OP_FLOATN;
  op.lanes = static_cast<uint8_t>(n);
  op.type = CLDN_FLOAT32;
  op.size = 4;
  for (size_t i = 0; i < n; ++i) {
    op.offset[i] = info.fields[i].offset;
    op.enc_mul_f[i] = 1.0F / info.fields[i].resolution;  // field_encoder.cpp:34 (float division)
    op.dec_mul_f[i] = info.fields[i].resolution;         // field_decoder.cpp:33
  }
}

static int check_resolution(const cldn_field_t& f, const char* who) {
  if (!(f.resolution > 0.0f)) {
    set_error("%s requires a resolution with value > 0.0 (field '%s')", who, f.name);
    return CLDN_ERR_INVALID_ARGUMENT;
  }
  return CLDN_OK;
}

static void fill_scalar(RegOp& op, uint8_t kind, const cldn_field_t& f) {
  memset(&op, 0, sizeof(op));
  op.kind = kind;
  op.lanes = 1;
  op.type = f.type;
  op.size = static_cast<uint8_t>(size_of_type(f.type));
  op.offset[0] = f.offset;
  if (kind == OP_F32_LOSSY) {
    // FieldEncoderFloat_Lossy<float>(offset, float res): multiplier_(1.0 / resolution) -> double divide, narrowed.
    op.enc_mul_f[0] = static_cast<float>(1.0 / static_cast<double>(f.resolution));
    op.dec_mul_f[0] = f.resolution;
  } else if (kind == OP_F64_LOSSY) {
    op.enc_mul_d = 1.0 / static_cast<double>(f.resolution);
    op.dec_mul_d = static_cast<double>(f.resolution);
  }
}

// CreateCompatibleEncoder (codec_common.cpp:116-155) / CreateCompatibleDecoder (:157-198)
static int push_compatible(const cldn_info_t& info, const cldn_field_t& f, bool decoder, Plan* p) {
  RegOp& op = p->ops[p->n_ops];
  const bool lossy = info.encoding_opt == CLDN_ENC_LOSSY;
  switch (f.type) {
    case CLDN_FLOAT32:
      if (lossy && f.has_resolution) {
        if (int rc = check_resolution(f, "FieldEncoder(Float/Lossy)")) return rc;
        fill_scalar(op, OP_F32_LOSSY, f);
      } else if (info.encoding_opt == CLDN_ENC_LOSSLESS) {
        fill_scalar(op, OP_XOR32, f);
      } else if (decoder && f.has_resolution) {  // legacy branch, decoder only (codec_common.cpp:165-168)
        if (int rc = check_resolution(f, "FieldDecoder(Float/Lossy)")) return rc;
        fill_scalar(op, OP_F32_LOSSY, f);
      } else {
        fill_scalar(op, OP_COPY, f);
      }
      break;
    case CLDN_FLOAT64:
      if (lossy && f.has_resolution) {
        if (int rc = check_resolution(f, "FieldEncoder(Float/Lossy)")) return rc;
        fill_scalar(op, OP_F64_LOSSY, f);
      } else if (decoder && f.has_resolution && info.encoding_opt != CLDN_ENC_LOSSLESS) {
        if (int rc = check_resolution(f, "FieldDecoder(Float/Lossy)")) return rc;
        fill_scalar(op, OP_F64_LOSSY, f);
      } else if (!f.has_resolution && info.version >= 4) {
        fill_scalar(op, OP_GORILLA64, f);
      } else {
        fill_scalar(op, OP_XOR64, f);
      }
      break;
    case CLDN_INT16: case CLDN_UINT16: case CLDN_INT32: case CLDN_UINT32: case CLDN_INT64: case CLDN_UINT64:
      fill_scalar(op, OP_INT, f);
      break;
    case CLDN_INT8: case CLDN_UINT8:
      fill_scalar(op, OP_COPY, f);
      break;
    default:
      set_error("Unsupported field type:%d", static_cast<int>(f.type));
      return CLDN_ERR_INVALID_ARGUMENT;
  }
  ++p->n_ops;
  return CLDN_OK;
}

static void finish_plan(Plan* p) {
  p->max_point_bytes = 0;
  p->min_point_bytes = 0;
  p->values_per_point = 0;
  p->n_gorilla = 0;
  p->all_varint = 1;
  p->all_fixed = 1;
  for (uint32_t i = 0; i < p->n_ops; ++i) {
    const RegOp& op = p->ops[i];
    switch (op.kind) {
      case OP_FLOATN:  // int32 delta -> zigzag+1 <= 2^32 -> at most 5 bytes per lane
        p->max_point_bytes += 5u * op.lanes; p->min_point_bytes += op.lanes; p->values_per_point += op.lanes; p->all_fixed = 0; break;
      case OP_F32_LOSSY: case OP_F64_LOSSY: case OP_INT:
        p->max_point_bytes += 10; p->min_point_bytes += 1; p->values_per_point += 1; p->all_fixed = 0; break;
      case OP_COPY:
        p->max_point_bytes += op.size; p->min_point_bytes += op.size; p->all_varint = 0; break;
      case OP_XOR32: case OP_XOR64:  // raw residual of the previous point's bits (field_encoder.hpp:360-370)
        p->max_point_bytes += op.size; p->min_point_bytes += op.size; p->all_varint = 0; p->all_fixed = 0; break;
      default:  // Gorilla: at most 1+1+5+6+64 = 77 bits = 10 bytes per value; minInputBytes() = 0 (field_decoder.hpp:163-166)
        p->max_point_bytes += 10; p->all_varint = 0; p->all_fixed = 0; ++p->n_gorilla; break;
    }
  }
  p->floatn_only = (p->n_ops == 1 && p->ops[0].kind == OP_FLOATN) ? 1 : 0;
  if (p->n_ops == 0) { p->all_varint = 0; }
  // stored byte ranges of the regular stream's fields: do any two intersect? (forged offsets, or e.g. `rgb` and `rgba`
  // declared at the same offset)
  p->regular_overlap = 0;
  uint64_t lo[kMaxOps * 4], hi[kMaxOps * 4];
  uint32_t n = 0;
  for (uint32_t i = 0; i < p->n_ops; ++i) {
    const RegOp& op = p->ops[i];
    const uint32_t width = (op.kind == OP_FLOATN) ? 4u : op.size;
    for (int l = 0; l < (op.kind == OP_FLOATN ? op.lanes : 1); ++l) {
      if (op.offset[l] == CLDN_SKIP_STORE_OFFSET) continue;
      lo[n] = op.offset[l]; hi[n] = static_cast<uint64_t>(op.offset[l]) + width; ++n;
    }
  }
  for (uint32_t a = 0; a < n; ++a) {
    for (uint32_t b = a + 1; b < n; ++b) {
      if (lo[a] < hi[b] && lo[b] < hi[a]) p->regular_overlap = 1;
    }
  }
}

static int build_plan(const cldn_info_t& info, bool decoder, Plan* p) {
  init_plan(info, p);
  if (info.n_fields > CLDN_MAX_FIELDS) { set_error("too many fields"); return CLDN_ERR_UNSUPPORTED; }
  // Memory safety: every field must lie inside the point. The reference never checks this (a forged header makes its
  // decoders write past the output buffer: field_decoder.cpp:74-78 stores at offset + i * point_step unconditionally);
  // the kernels would do the same to device memory, so such an EncodingInfo is refused here, loudly.
  for (uint32_t i = 0; i < info.n_fields; ++i) {
    const cldn_field_t& f = info.fields[i];
    const uint64_t end = static_cast<uint64_t>(f.offset) + static_cast<uint64_t>(size_of_type(f.type));
    if (!(decoder && f.offset == CLDN_SKIP_STORE_OFFSET) && end > info.point_step) {
      set_error("field '%s' (offset %u, %d bytes) does not fit a point of %u bytes", f.name, f.offset, size_of_type(f.type), info.point_step);
      return CLDN_ERR_INVALID_ARGUMENT;
    }
  }
  const bool v5 = uses_v5_codec(info);
  p->uses_v5 = v5 ? 1 : 0;
  if (!v5 && info.encoding_opt == CLDN_ENC_NONE) {  // BuildV4Encoders, v4_codec.cpp:29-34: everything is a raw copy
    for (uint32_t i = 0; i < info.n_fields; ++i) {
      if (size_of_type(info.fields[i].type) == 0) { set_error("Unsupported field type"); return CLDN_ERR_INVALID_ARGUMENT; }
      fill_scalar(p->ops[p->n_ops++], OP_COPY, info.fields[i]);
    }
    finish_plan(p);
    return CLDN_OK;
  }
  const size_t lead = leading_lossy_float_count(info);
  if (lead) {
    for (size_t i = 0; i < lead; ++i) {
      // FieldEncoderFloatN_Lossy ctor: multiplier = 1.0F/res must be > 0 (field_encoder.cpp:34-37)
      const float m = decoder ? info.fields[i].resolution : 1.0F / info.fields[i].resolution;
      if (!(m > 0.0f)) { set_error("FieldEncoderFloatN_Lossy requires a resolution with value > 0.0"); return CLDN_ERR_INVALID_ARGUMENT; }
    }
    push_floatn(info, lead, p);
  }
  for (uint32_t i = static_cast<uint32_t>(lead); i < info.n_fields; ++i) {
    const cldn_field_t& f = info.fields[i];
    // buildV5Plan (v5_codec.cpp:719-740) / BuildV5Decoders (:965-982): adaptive ints leave the interleaved stream.
    if (v5 && is_adaptive_int(f.type)) {
      SectionField& s = p->sections[p->n_sections++];
      s.offset = f.offset;
      s.type = f.type;
      s.bpv = static_cast<uint8_t>(size_of_type(f.type));
      s.field_index = static_cast<uint16_t>(i);
      continue;
    }
    if (int rc = push_compatible(info, f, decoder, p)) return rc;
  }
  finish_plan(p);
  return CLDN_OK;
}

int build_encode_plan(const cldn_info_t& info, Plan* plan) { return build_plan(info, false, plan); }
int build_decode_plan(const cldn_info_t& info, Plan* plan) { return build_plan(info, true, plan); }

// EncodeHeader(YAML), cloudini.cpp:294-318
std::vector<uint8_t> make_header(const cldn_info_t& info) {
  const std::string yaml = info_to_yaml(info);
  std::vector<uint8_t> h;
  h.reserve(yaml.size() + 14);
  const char* magic = "CLOUDINI_V";
  h.insert(h.end(), magic, magic + 10);
  h.push_back(static_cast<uint8_t>('0' + (info.version / 10)));
  h.push_back(static_cast<uint8_t>('0' + (info.version % 10)));
  h.push_back('\n');
  h.insert(h.end(), yaml.begin(), yaml.end());
  h.push_back('\0');
  return h;
}

}  // namespace cldn

using namespace cldn;

extern "C" {

const char* cldn_b200_version(void) { return "cloudini_b200 0.1.0 (wire v5, sm_100a)"; }
const char* cldn_b200_last_error(void) { return last_error_cstr(); }

void cldn_b200_info_init(cldn_info_t* info) {
  memset(info, 0, sizeof(*info));
  info->height = 1;
  info->encoding_opt = CLDN_ENC_LOSSY;
  info->compression_opt = CLDN_COMP_ZSTD;  // cloudini.hpp:84
  info->use_threads = 1;
  info->version = CLDN_ENCODING_VERSION;
}

int cldn_b200_info_to_yaml(const cldn_info_t* info, char* out, size_t capacity, size_t* needed) {
  if (!info) { set_error("null info"); return CLDN_ERR_INVALID_ARGUMENT; }
  const std::string y = info_to_yaml(*info);
  if (needed) *needed = y.size() + 1;
  if (!out || capacity < y.size() + 1) { set_error("yaml buffer too small"); return CLDN_ERR_BUFFER_TOO_SMALL; }
  memcpy(out, y.c_str(), y.size() + 1);
  return CLDN_OK;
}

int cldn_b200_info_from_yaml(const char* yaml, size_t yaml_len, cldn_info_t* info) {
  if (!yaml || !info) { set_error("null argument"); return CLDN_ERR_INVALID_ARGUMENT; }
  if (yaml_len == 0) yaml_len = strlen(yaml);
  return info_from_yaml(yaml, yaml_len, info);
}

int cldn_b200_encode_header(const cldn_info_t* info, uint8_t* out, size_t capacity, size_t* written) {
  if (!info) { set_error("null info"); return CLDN_ERR_INVALID_ARGUMENT; }
  const std::vector<uint8_t> h = make_header(*info);
  if (written) *written = h.size();
  if (!out || capacity < h.size()) { set_error("header buffer too small"); return CLDN_ERR_BUFFER_TOO_SMALL; }
  memcpy(out, h.data(), h.size());
  return CLDN_OK;
}

// EncodeHeader, HeaderEncoding::BINARY (cloudini.cpp:319-344; size: ComputeHeaderSize :232-247)
int cldn_b200_encode_header_binary(const cldn_info_t* info, uint8_t* out, size_t capacity, size_t* written) {
  if (!info) { set_error("null info"); return CLDN_ERR_INVALID_ARGUMENT; }
  if (info->n_fields > CLDN_MAX_FIELDS) { set_error("too many fields"); return CLDN_ERR_INVALID_ARGUMENT; }
  std::vector<uint8_t> h;
  auto put = [&h](const void* p, size_t k) { const uint8_t* b = static_cast<const uint8_t*>(p); h.insert(h.end(), b, b + k); };
  put("CLOUDINI_V", 10);
  const char digits[2] = {static_cast<char>('0' + info->version / 10), static_cast<char>('0' + info->version % 10)};
  put(digits, 2);
  put(&info->width, 4);
  put(&info->height, 4);
  put(&info->point_step, 4);
  put(&info->encoding_opt, 1);
  put(&info->compression_opt, 1);
  const uint16_t nf = static_cast<uint16_t>(info->n_fields);
  put(&nf, 2);
  for (uint32_t i = 0; i < info->n_fields; ++i) {
    const cldn_field_t& f = info->fields[i];
    const uint16_t len = static_cast<uint16_t>(strnlen(f.name, CLDN_MAX_NAME));
    put(&len, 2);
    put(f.name, len);
    put(&f.offset, 4);
    put(&f.type, 1);
    const float res = f.has_resolution ? f.resolution : -1.0f;  // "no resolution" travels as -1
    put(&res, 4);
  }
  if (written) *written = h.size();
  if (!out || capacity < h.size()) { set_error("header buffer too small"); return CLDN_ERR_BUFFER_TOO_SMALL; }
  memcpy(out, h.data(), h.size());
  return CLDN_OK;
}

// DecodeHeader, cloudini.cpp:353-428
int cldn_b200_decode_header(const uint8_t* blob, size_t n, cldn_info_t* info, size_t* header_bytes) {
  if (!blob || !info) { set_error("null argument"); return CLDN_ERR_INVALID_ARGUMENT; }
  if (n < 12) { set_error("Input too small to contain Cloudini header"); return CLDN_ERR_BAD_HEADER; }
  if (memcmp(blob, "CLOUDINI_V", 10) != 0) {
    set_error("Invalid magic header. Expected 'CLOUDINI_V', got: %.10s", reinterpret_cast<const char*>(blob));
    return CLDN_ERR_BAD_HEADER;
  }
  auto digit = [](uint8_t c) -> int { return (c >= '0' && c <= '9') ? c - '0' : 0; };
  const int version = digit(blob[10]) * 10 + digit(blob[11]);
  if (version < 2 || version > CLDN_ENCODING_VERSION) {
    set_error("Unsupported encoding version. Current is:%d, got: %d", CLDN_ENCODING_VERSION, version);
    return CLDN_ERR_BAD_HEADER;
  }
  size_t pos = 12;
  if (n - pos >= 2 && blob[pos] == '\n' && blob[pos + 1] != '{') {
    ++pos;
    const void* nul = memchr(blob + pos, 0, n - pos);
    if (!nul) { set_error("Malformed YAML header: missing null terminator"); return CLDN_ERR_BAD_HEADER; }
    const size_t ylen = static_cast<const uint8_t*>(nul) - (blob + pos);
    int rc = info_from_yaml(reinterpret_cast<const char*>(blob + pos), ylen, info);
    if (rc != CLDN_OK) return rc;
    info->version = static_cast<uint8_t>(version);  // the magic's version is authoritative (cloudini.cpp:389-392)
    if (header_bytes) *header_bytes = pos + ylen + 1;
    return CLDN_OK;
  }
  // legacy binary header (cloudini.cpp:395-427)
  cldn_b200_info_init(info);
  info->version = static_cast<uint8_t>(version);
  auto need = [&](size_t k) { return n - pos >= k; };
  auto fail = [&]() { set_error("decode: not enough input data"); return CLDN_ERR_BAD_HEADER; };
  if (!need(14)) return fail();
  memcpy(&info->width, blob + pos, 4); pos += 4;
  memcpy(&info->height, blob + pos, 4); pos += 4;
  memcpy(&info->point_step, blob + pos, 4); pos += 4;
  info->encoding_opt = blob[pos++];
  info->compression_opt = blob[pos++];
  uint16_t nf; memcpy(&nf, blob + pos, 2); pos += 2;
  if (nf > CLDN_MAX_FIELDS) { set_error("too many fields in binary header"); return CLDN_ERR_UNSUPPORTED; }
  info->n_fields = nf;
  for (uint16_t i = 0; i < nf; ++i) {
    cldn_field_t& f = info->fields[i];
    memset(&f, 0, sizeof(f));
    if (!need(2)) return fail();
    uint16_t len; memcpy(&len, blob + pos, 2); pos += 2;
    if (!need(len)) { set_error("decode(string): not enough input data"); return CLDN_ERR_BAD_HEADER; }
    const size_t cp = std::min<size_t>(len, CLDN_MAX_NAME - 1);
    memcpy(f.name, blob + pos, cp); pos += len;
    if (!need(9)) return fail();
    memcpy(&f.offset, blob + pos, 4); pos += 4;
    f.type = blob[pos++];
    float res; memcpy(&res, blob + pos, 4); pos += 4;
    if (res > 0) { f.has_resolution = 1; f.resolution = res; }
  }
  if (header_bytes) *header_bytes = pos;
  return CLDN_OK;
}

size_t cldn_b200_max_compressed_size(const cldn_info_t* info, size_t points_count, int include_header) {
  if (!info) { set_error("null info"); return 0; }
  bool ok;
  return max_compressed_size(*info, points_count, include_header != 0, &ok);
}

}  // extern "C"
