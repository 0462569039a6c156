/*
 * TEST INFRASTRUCTURE ONLY — "parity pinned".
 *
 * Plain-C, single-threaded CPU restatement of the reference's stage-1 codec path, written from the algorithm
 * (not copied): used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg as the CHECKER for the CUDA
 * kernels. Nothing in the product (cloudini_b200/, include/) links, imports or calls this file.
 *
 * Pinned against the reference itself: tests/test_oracle.py compares every function here byte-for-byte with
 * oracle/_ref/libcloudini_ref.so (the unmodified reference sources compiled in place) on the reference tests'
 * known-answer cases (test_field_encoders.cpp:590-769, test_header.cpp:107-163, test_intrinsics.cpp:37-41) and with
 * the committed golden vectors in tests/golden/ (generated from the reference by tests/golden/make_golden.py).
 *
 * Reference map (paths relative to /root/reference/cloudini_lib):
 *   varint / zigzag            include/cloudini_lib/encoding_utils.hpp:55-67, 98-148
 *   float -> int32 rounding    include/cloudini_lib/intrinsics.hpp:288-300 (SSE4.1 branch)
 *   FloatN encode / decode     src/field_encoder.cpp:24-91, src/field_decoder.cpp:24-86
 *   scalar lossy / int / copy  include/cloudini_lib/field_encoder.hpp:51-118,343-357; field_decoder.hpp:56-130,331-353
 *   planner                    src/codec_common.cpp:29-198, src/v4_codec.cpp:26-64, src/v5_codec.cpp:719-763,883-892
 *   V4 chunk loop              src/v4_codec.cpp:66-117
 *   V5 adaptive ints           src/v5_codec.cpp:160-491 (sizes, mode choice, writers), 764-879 (reader), 900-1012
 *   framing + header + sizing  src/chunk_writer.cpp:27-48, src/cloudini.cpp:165-190, 249-344, 353-428, 501-684
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (see oracle/build_oracle.sh). x86-64 SSE2 scalar float math is
 * IEEE single/double, so `v * m` below is the same correctly-rounded product as _mm_mul_ps.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/cloudini_b200.h" /* only for the POD cldn_info_t / cldn_field_t layout and enums */

#define ORC_CHUNK 32768u /* codec_common.hpp:28 */
#define ORC_PROBE 4096u  /* v5_codec.cpp:76 */

static char g_err[256];
const char* orc_last_error(void) { return g_err; }
static int fail(const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return -1;
}

/* ---------------------------------------------------------------- basic helpers */
static int type_size(uint8_t t) { /* basic_types.hpp:73-96 */
  switch (t) {
    case CLDN_INT8: case CLDN_UINT8: return 1;
    case CLDN_INT16: case CLDN_UINT16: return 2;
    case CLDN_INT32: case CLDN_UINT32: case CLDN_FLOAT32: return 4;
    case CLDN_FLOAT64: case CLDN_INT64: case CLDN_UINT64: return 8;
    default: return 0;
  }
}

/* encodeVarint64: zigzag, +1 (0 is the NaN marker), LEB128 */
static size_t put_varint(int64_t v, uint8_t* p) {
  uint64_t u = (((uint64_t)v) << 1) ^ (uint64_t)(v >> 63);
  u += 1;
  size_t n = 0;
  while (u > 0x7F) {
    p[n++] = (uint8_t)((u & 0x7F) | 0x80);
    u >>= 7;
  }
  p[n++] = (uint8_t)u;
  return n;
}
static size_t varint_size(int64_t v) {
  uint8_t tmp[10];
  return put_varint(v, tmp);
}
static size_t put_uvarint(uint64_t u, uint8_t* p) { /* v5_codec.cpp:160-174 */
  size_t n = 0;
  while (u > 0x7F) {
    p[n++] = (uint8_t)((u & 0x7F) | 0x80);
    u >>= 7;
  }
  p[n++] = (uint8_t)u;
  return n;
}
static size_t uvarint_size(uint64_t u) {
  size_t n = 1;
  while (u > 0x7F) { u >>= 7; ++n; }
  return n;
}
/* decodeVarint: returns bytes consumed, 0 on error (message set) */
static size_t get_varint(const uint8_t* p, size_t avail, int64_t* out) {
  if (avail == 0) { fail("decodeVarint: empty input"); return 0; }
  uint64_t u = 0;
  unsigned shift = 0;
  size_t n = 0;
  for (;;) {
    if (n >= avail) { fail("decodeVarint: truncated input"); return 0; }
    const uint8_t b = p[n++];
    const uint64_t payload = b & 0x7F;
    if (shift >= 64 || (shift == 63 && payload > 1)) { fail("decodeVarint: value overflow"); return 0; }
    u |= payload << shift;
    if (!(b & 0x80)) break;
    if (shift >= 63) { fail("decodeVarint: value overflow"); return 0; }
    shift += 7;
  }
  if (u == 0) { fail("decodeVarint: unexpected NaN marker"); return 0; }
  u -= 1;
  *out = (int64_t)((u >> 1) ^ (uint64_t)(-(int64_t)(u & 1)));
  return n;
}
static size_t get_uvarint(const uint8_t* p, size_t avail, uint64_t* out) { /* v5_codec.cpp:176-194 */
  uint64_t v = 0;
  unsigned shift = 0;
  size_t n = 0;
  for (;;) {
    if (n >= avail) { fail("V5 adaptive int: truncated unsigned varint"); return 0; }
    const uint8_t b = p[n++];
    v |= ((uint64_t)(b & 0x7F)) << shift;
    if (!(b & 0x80)) break;
    shift += 7;
    if (shift >= 64) { fail("V5 adaptive int: unsigned varint overflow"); return 0; }
  }
  *out = v;
  return n;
}

/* _mm_round_ps(nearest-even) + _mm_cvtps_epi32: out-of-range / NaN give the "integer indefinite" 0x80000000 */
static int32_t round_even_i32(float s) {
  const float r = nearbyintf(s); /* default rounding mode: to nearest, ties to even */
  if (!(r >= -2147483648.0f && r < 2147483648.0f)) return INT32_MIN;
  return (int32_t)r;
}
/* static_cast<int64_t>(std::round(x)) as compiled for x86-64 (cvttss2si / cvttsd2si): indefinite = INT64_MIN */
static int64_t round_away_i64_f(float s) {
  const float r = roundf(s);
  if (!(r >= -9223372036854775808.0f && r < 9223372036854775808.0f)) return INT64_MIN;
  return (int64_t)r;
}
static int64_t round_away_i64_d(double s) {
  const double r = round(s);
  if (!(r >= -9223372036854775808.0 && r < 9223372036854775808.0)) return INT64_MIN;
  return (int64_t)r;
}

static int64_t read_int(const uint8_t* p, uint8_t t) { /* ToInt64<T> / readIntAsI64 */
  switch (t) {
    case CLDN_INT8: { int8_t v; memcpy(&v, p, 1); return v; }
    case CLDN_UINT8: { uint8_t v; memcpy(&v, p, 1); return v; }
    case CLDN_INT16: { int16_t v; memcpy(&v, p, 2); return v; }
    case CLDN_UINT16: { uint16_t v; memcpy(&v, p, 2); return v; }
    case CLDN_INT32: { int32_t v; memcpy(&v, p, 4); return v; }
    case CLDN_UINT32: { uint32_t v; memcpy(&v, p, 4); return v; }
    default: { int64_t v; memcpy(&v, p, 8); return v; }
  }
}
static uint64_t read_raw(const uint8_t* p, int bytes) {
  uint64_t v = 0;
  memcpy(&v, p, (size_t)bytes);
  return v;
}

/* ---------------------------------------------------------------- planning */
enum { K_FLOATN, K_F32, K_F64, K_INT, K_COPY, K_XOR32, K_XOR64, K_GORILLA64, K_UNSUPPORTED };
typedef struct {
  int kind, lanes, size;
  uint8_t type;
  uint32_t off[4];
  float enc_mul[4], dec_mul[4];
  double enc_mul_d, dec_mul_d;
  /* state */
  int32_t prev32[4];
  int64_t prev64;
  uint64_t prev_bits; /* XOR / Gorilla */
  int g_first, g_lead, g_trail; /* Gorilla window (field_encoder.hpp:157-183) */
} op_t;
typedef struct {
  uint32_t off;
  uint8_t type;
  int bpv;
  int committed, mode;
} sec_t;
typedef struct {
  op_t ops[CLDN_MAX_FIELDS];
  int n_ops;
  sec_t secs[CLDN_MAX_FIELDS];
  int n_secs;
  int v5;
  size_t min_point_bytes;
} plan_t;

static size_t leading_floats(const cldn_info_t* in) { /* codec_common.cpp:69-82 */
  if (in->encoding_opt != CLDN_ENC_LOSSY) return 0;
  size_t n = 0;
  for (uint32_t i = 0; i < in->n_fields; ++i) {
    if (in->fields[i].type != CLDN_FLOAT32 || !in->fields[i].has_resolution) break;
    ++n;
  }
  return (n == 3 || n == 4) ? n : 0;
}
static int is_adaptive(uint8_t t) {
  return t == CLDN_INT16 || t == CLDN_UINT16 || t == CLDN_INT32 || t == CLDN_UINT32 || t == CLDN_INT64 || t == CLDN_UINT64;
}
static int uses_v5(const cldn_info_t* in) { /* v5_codec.cpp:883-892 */
  if (in->version < 5 || in->encoding_opt != CLDN_ENC_LOSSY) return 0;
  for (uint32_t i = (uint32_t)leading_floats(in); i < in->n_fields; ++i)
    if (is_adaptive(in->fields[i].type)) return 1;
  return 0;
}

static int make_plan(const cldn_info_t* in, int decoder, plan_t* pl) {
  memset(pl, 0, sizeof(*pl));
  pl->v5 = uses_v5(in);
  if (!pl->v5 && in->encoding_opt == CLDN_ENC_NONE) { /* v4_codec.cpp:29-34 */
    for (uint32_t i = 0; i < in->n_fields; ++i) {
      op_t* o = &pl->ops[pl->n_ops++];
      o->kind = K_COPY; o->lanes = 1; o->type = in->fields[i].type; o->size = type_size(o->type); o->off[0] = in->fields[i].offset;
      pl->min_point_bytes += (size_t)o->size;
    }
    return 0;
  }
  const size_t lead = leading_floats(in);
  if (lead) {
    op_t* o = &pl->ops[pl->n_ops++];
    o->kind = K_FLOATN; o->lanes = (int)lead; o->size = 4;
    for (size_t i = 0; i < lead; ++i) {
      o->off[i] = in->fields[i].offset;
      o->enc_mul[i] = 1.0F / in->fields[i].resolution; /* field_encoder.cpp:34 */
      o->dec_mul[i] = in->fields[i].resolution;        /* field_decoder.cpp:33 */
      if (!((decoder ? o->dec_mul[i] : o->enc_mul[i]) > 0.0f)) return fail("FieldEncoderFloatN_Lossy requires a resolution with value > 0.0");
    }
    pl->min_point_bytes += lead;
  }
  const int lossy = in->encoding_opt == CLDN_ENC_LOSSY;
  for (uint32_t i = (uint32_t)lead; i < in->n_fields; ++i) {
    const cldn_field_t* f = &in->fields[i];
    if (pl->v5 && is_adaptive(f->type)) {
      sec_t* s = &pl->secs[pl->n_secs++];
      s->off = f->offset; s->type = f->type; s->bpv = type_size(f->type);
      continue;
    }
    op_t* o = &pl->ops[pl->n_ops++];
    o->lanes = 1; o->type = f->type; o->size = type_size(f->type); o->off[0] = f->offset;
    switch (f->type) { /* codec_common.cpp:116-198 */
      case CLDN_FLOAT32:
        if ((lossy && f->has_resolution) || (decoder && f->has_resolution && in->encoding_opt != CLDN_ENC_LOSSLESS)) {
          o->kind = K_F32;
          o->enc_mul[0] = (float)(1.0 / (double)f->resolution); /* field_encoder.hpp:101-102 */
          o->dec_mul[0] = f->resolution;
          if (!(f->resolution > 0.0f)) return fail("FieldEncoder(Float/Lossy) requires a resolution with value > 0.0");
          pl->min_point_bytes += 1;
        } else if (in->encoding_opt == CLDN_ENC_LOSSLESS) {
          o->kind = K_XOR32; /* field_encoder.hpp:360-370 */
          pl->min_point_bytes += 4;
        } else {
          o->kind = K_COPY;
          pl->min_point_bytes += 4;
        }
        break;
      case CLDN_FLOAT64:
        if ((lossy && f->has_resolution) || (decoder && f->has_resolution && in->encoding_opt != CLDN_ENC_LOSSLESS)) {
          o->kind = K_F64;
          o->enc_mul_d = 1.0 / (double)f->resolution;
          o->dec_mul_d = (double)f->resolution;
          if (!(f->resolution > 0.0f)) return fail("FieldEncoder(Float/Lossy) requires a resolution with value > 0.0");
          pl->min_point_bytes += 1;
        } else if (!f->has_resolution && in->version >= 4) {
          o->kind = K_GORILLA64; /* codec_common.cpp:129-131; minInputBytes() = 0 (field_decoder.hpp:163-166) */
        } else {
          o->kind = K_XOR64;
          pl->min_point_bytes += 8;
        }
        break;
      case CLDN_INT16: case CLDN_UINT16: case CLDN_INT32: case CLDN_UINT32: case CLDN_INT64: case CLDN_UINT64:
        o->kind = K_INT;
        pl->min_point_bytes += 1;
        break;
      case CLDN_INT8: case CLDN_UINT8:
        o->kind = K_COPY;
        pl->min_point_bytes += 1;
        break;
      default:
        return fail("Unsupported field type");
    }
  }
  return 0;
}

static void reset_ops(plan_t* pl) {
  for (int i = 0; i < pl->n_ops; ++i) {
    memset(pl->ops[i].prev32, 0, sizeof(pl->ops[i].prev32));
    pl->ops[i].prev64 = 0;
    pl->ops[i].prev_bits = 0;
    pl->ops[i].g_first = 1;
    pl->ops[i].g_lead = 255; /* kLeadingSentinel */
    pl->ops[i].g_trail = 0;
  }
}

/* ---- Gorilla / Chimp-style bit packing of FLOAT64 (field_encoder.hpp:157-312): every value is flushed to a byte
 * boundary, bits are appended LSB-first. Returns the number of bytes written (<= 10). */
typedef struct { uint64_t lo, hi; unsigned n; } bitacc_t;
static void acc_put(bitacc_t* a, uint64_t bits, unsigned nbits) {
  if (nbits < 64) bits &= ((uint64_t)1 << nbits) - 1;
  if (a->n < 64) {
    a->lo |= bits << a->n;
    if (a->n + nbits > 64) a->hi |= bits >> (64 - a->n);
  } else {
    a->hi |= bits << (a->n - 64);
  }
  a->n += nbits;
}
static size_t gorilla_encode(op_t* o, uint64_t cur, uint8_t* out) {
  bitacc_t a = {0, 0, 0};
  if (o->g_first) {
    o->g_first = 0;
    o->prev_bits = cur;
    acc_put(&a, cur, 64);
  } else {
    const uint64_t x = cur ^ o->prev_bits;
    o->prev_bits = cur;
    if (x == 0) {
      acc_put(&a, 0, 1);
    } else {
      acc_put(&a, 1, 1);
      const int leading = __builtin_clzll(x), trailing = __builtin_ctzll(x);
      if (o->g_lead != 255 && leading >= o->g_lead && trailing >= o->g_trail) {
        acc_put(&a, 0, 1);
        acc_put(&a, x >> o->g_trail, (unsigned)(64 - o->g_lead - o->g_trail));
      } else {
        acc_put(&a, 1, 1);
        const int stored = leading > 31 ? 31 : leading;
        const unsigned meaningful = (unsigned)(64 - stored - trailing);
        acc_put(&a, (uint64_t)stored, 5);
        acc_put(&a, (uint64_t)(meaningful - 1), 6);
        acc_put(&a, x >> trailing, meaningful);
        o->g_lead = stored;
        o->g_trail = trailing;
      }
    }
  }
  const size_t bytes = (a.n + 7) / 8;
  for (size_t i = 0; i < bytes; ++i) out[i] = (uint8_t)(i < 8 ? (a.lo >> (8 * i)) : (a.hi >> (8 * (i - 8))));
  return bytes;
}
/* bit reader over one value's bytes (field_decoder.hpp:200-300): returns bytes consumed, 0 on truncation */
typedef struct { const uint8_t* p; size_t avail; unsigned bitpos; } bitrd_t;
static int rd_bits(bitrd_t* r, unsigned nbits, uint64_t* out) {
  if (((size_t)r->bitpos + nbits + 7) / 8 > r->avail) return fail("FieldDecoderFloat_Gorilla: truncated input");
  uint64_t v = 0;
  for (unsigned i = 0; i < nbits; ++i) {
    const unsigned b = r->bitpos + i;
    v |= (uint64_t)((r->p[b >> 3] >> (b & 7)) & 1u) << i;
  }
  r->bitpos += nbits;
  *out = v;
  return 0;
}
static size_t gorilla_decode(op_t* o, const uint8_t* p, size_t avail, uint64_t* value) {
  bitrd_t r = {p, avail, 0};
  uint64_t v;
  if (o->g_first) {
    o->g_first = 0;
    if (rd_bits(&r, 64, &v)) return 0;
    o->prev_bits = v;
  } else {
    uint64_t flag;
    if (rd_bits(&r, 1, &flag)) return 0;
    if (flag == 0) {
      v = o->prev_bits;
    } else {
      uint64_t control, bits, x;
      if (rd_bits(&r, 1, &control)) return 0;
      if (control == 0) {
        /* uint8_t(64 - prev_leading - prev_trailing): 65 before any window (sentinel 255): the reference's getBits then
         * consumes 65 bits and keeps the low 64 (field_decoder.hpp:213-241, 274-276) */
        const unsigned meaningful = (unsigned)(uint8_t)(64 - o->g_lead - o->g_trail);
        if (meaningful > 65) { fail("oracle: impossible Gorilla window"); return 0; }
        if (rd_bits(&r, meaningful > 64 ? 64 : meaningful, &bits)) return 0;
        if (meaningful > 64) { uint64_t dropped; if (rd_bits(&r, 1, &dropped)) return 0; }
        x = bits << ((unsigned)o->g_trail & 63u); /* uint8_t shift count, x86 semantics (count mod 64) for forged windows */
      } else {
        uint64_t lead, m1;
        if (rd_bits(&r, 5, &lead) || rd_bits(&r, 6, &m1)) return 0;
        const unsigned meaningful = (unsigned)m1 + 1;
        if (rd_bits(&r, meaningful, &bits)) return 0;
        const unsigned trailing = (unsigned)(uint8_t)(64 - lead - meaningful);
        x = bits << (trailing & 63u);
        o->g_lead = (int)lead;
        o->g_trail = (int)trailing;
      }
      v = x ^ o->prev_bits;
      o->prev_bits = v;
    }
  }
  *value = v;
  return (r.bitpos + 7) / 8; /* leftover bits of the last byte are padding */
}

/* ---------------------------------------------------------------- per-point regular encoders */
static size_t encode_point(plan_t* pl, const uint8_t* pt, uint8_t* out) {
  size_t n = 0;
  for (int k = 0; k < pl->n_ops; ++k) {
    op_t* o = &pl->ops[k];
    switch (o->kind) {
      case K_FLOATN: /* field_encoder.cpp:42-91 */
        for (int l = 0; l < o->lanes; ++l) {
          float v;
          memcpy(&v, pt + o->off[l], 4);
          const int32_t q = round_even_i32(v * o->enc_mul[l]);
          const int32_t d = (int32_t)((uint32_t)q - (uint32_t)o->prev32[l]); /* _mm_sub_epi32 wraps */
          o->prev32[l] = q;
          if (isnan(v)) {
            out[n++] = 0;
            o->prev32[l] = 0;
          } else {
            n += put_varint((int64_t)d, out + n);
          }
        }
        break;
      case K_F32: { /* field_encoder.hpp:343-357 */
        float v;
        memcpy(&v, pt + o->off[0], 4);
        if (isnan(v)) { out[n++] = 0; o->prev64 = 0; break; }
        const int64_t q = round_away_i64_f(v * o->enc_mul[0]);
        const int64_t d = (int64_t)((uint64_t)q - (uint64_t)o->prev64);
        o->prev64 = q;
        n += put_varint(d, out + n);
      } break;
      case K_F64: {
        double v;
        memcpy(&v, pt + o->off[0], 8);
        if (isnan(v)) { out[n++] = 0; o->prev64 = 0; break; }
        const int64_t q = round_away_i64_d(v * o->enc_mul_d);
        const int64_t d = (int64_t)((uint64_t)q - (uint64_t)o->prev64);
        o->prev64 = q;
        n += put_varint(d, out + n);
      } break;
      case K_INT: { /* field_encoder.hpp:78-85 */
        const int64_t v = read_int(pt + o->off[0], o->type);
        const int64_t d = (int64_t)((uint64_t)v - (uint64_t)o->prev64);
        o->prev64 = v;
        n += put_varint(d, out + n);
      } break;
      case K_XOR32: case K_XOR64: { /* field_encoder.hpp:360-370: residual = bits ^ previous bits, raw */
        uint64_t cur = 0;
        memcpy(&cur, pt + o->off[0], (size_t)o->size);
        const uint64_t res = cur ^ o->prev_bits;
        o->prev_bits = cur;
        memcpy(out + n, &res, (size_t)o->size);
        n += (size_t)o->size;
      } break;
      case K_GORILLA64: {
        uint64_t cur;
        memcpy(&cur, pt + o->off[0], 8);
        n += gorilla_encode(o, cur, out + n);
      } break;
      default: /* copy, field_encoder.hpp:56-60 */
        memcpy(out + n, pt + o->off[0], (size_t)o->size);
        n += (size_t)o->size;
        break;
    }
  }
  return n;
}

/* ---------------------------------------------------------------- V5 adaptive integer sections (encode) */
typedef struct {
  int64_t* v;     /* values as int64 */
  uint64_t* raw;  /* raw bits */
  size_t n;
} vals_t;

static size_t delta_section_size(const vals_t* a, size_t n) { /* v5_codec.cpp:258-267 */
  size_t bytes = 1;
  int64_t prev = 0;
  for (size_t i = 0; i < n; ++i) {
    bytes += varint_size((int64_t)((uint64_t)a->v[i] - (uint64_t)prev));
    prev = a->v[i];
  }
  return bytes;
}
static size_t delta_rle_section(const vals_t* a, size_t n, uint8_t* out) { /* sizes :290-299, writer :447-460; out may be NULL */
  size_t bytes = 5;
  uint32_t runs = 0;
  int64_t prev = 0;
  size_t i = 0;
  while (i < n) {
    const int64_t diff = (int64_t)((uint64_t)a->v[i] - (uint64_t)prev);
    prev = a->v[i];
    size_t j = i + 1;
    while (j < n && (int64_t)((uint64_t)a->v[j] - (uint64_t)prev) == diff) { prev = a->v[j]; ++j; }
    if (out) {
      bytes += put_varint(diff, out + bytes);
      bytes += put_uvarint(j - i, out + bytes);
    } else {
      bytes += varint_size(diff) + uvarint_size(j - i);
    }
    ++runs;
    i = j;
  }
  if (out) { out[0] = 3; memcpy(out + 1, &runs, 4); }
  return bytes;
}
static size_t rle_section(const vals_t* a, size_t n, int bpv, uint8_t* out) { /* sizes :301-318, writer :471-491 */
  size_t bytes = 5;
  uint32_t runs = 0;
  size_t i = 0;
  while (i < n) {
    size_t j = i + 1;
    while (j < n && a->raw[j] == a->raw[i]) ++j;
    if (out) {
      memcpy(out + bytes, &a->raw[i], (size_t)bpv);
      bytes += (size_t)bpv;
      bytes += put_uvarint(j - i, out + bytes);
    } else {
      bytes += (size_t)bpv + uvarint_size(j - i);
    }
    ++runs;
    i = j;
  }
  if (out) { out[0] = 2; memcpy(out + 1, &runs, 4); }
  return bytes;
}
static unsigned palette_bits(size_t unique) { /* v5_codec.cpp:196-207 */
  if (unique <= 1) return 0;
  unsigned bits = 0;
  size_t m = unique - 1;
  while (m) { ++bits; m >>= 1; }
  return bits;
}
/* unique values in first-appearance order; idx[i] = palette index of value i. Returns the palette size. */
static size_t build_palette(const vals_t* a, size_t n, uint64_t* pal, uint32_t* idx) {
  size_t cap = 16;
  while (cap < 2 * n) cap <<= 1;
  uint32_t* slots = (uint32_t*)calloc(cap, sizeof(uint32_t)); /* palette index + 1, 0 = empty */
  size_t count = 0;
  for (size_t i = 0; i < n; ++i) {
    uint64_t h = a->raw[i] * 0x9E3779B97F4A7C15ull; /* any hash works: only first-appearance order is observable */
    size_t s = (size_t)(h >> 20) & (cap - 1);
    for (;;) {
      if (slots[s] == 0) { pal[count] = a->raw[i]; slots[s] = (uint32_t)(count + 1); idx[i] = (uint32_t)count; ++count; break; }
      if (pal[slots[s] - 1] == a->raw[i]) { idx[i] = slots[s] - 1; break; }
      s = (s + 1) & (cap - 1);
    }
  }
  free(slots);
  return count;
}
static size_t palette_section(const vals_t* a, size_t n, int bpv, uint8_t* out) { /* sizes :381-385, writer :462-469,209-227 */
  uint64_t* pal = (uint64_t*)malloc((n ? n : 1) * sizeof(uint64_t));
  uint32_t* idx = (uint32_t*)malloc((n ? n : 1) * sizeof(uint32_t));
  const size_t u = build_palette(a, n, pal, idx);
  const unsigned bits = palette_bits(u);
  size_t bytes = 3 + u * (size_t)bpv + ((size_t)bits * n + 7) / 8;
  if (out) {
    out[0] = 1;
    const uint16_t cnt = (uint16_t)u;
    memcpy(out + 1, &cnt, 2);
    size_t p = 3;
    for (size_t k = 0; k < u; ++k) { memcpy(out + p, &pal[k], (size_t)bpv); p += (size_t)bpv; }
    if (bits) {
      uint64_t scratch = 0;
      unsigned held = 0;
      for (size_t i = 0; i < n; ++i) {
        scratch |= ((uint64_t)idx[i]) << held;
        held += bits;
        while (held >= 8) { out[p++] = (uint8_t)(scratch & 0xFF); scratch >>= 8; held -= 8; }
      }
      if (held) out[p++] = (uint8_t)(scratch & 0xFF);
    }
  }
  free(pal);
  free(idx);
  return bytes;
}
static size_t delta_section(const vals_t* a, size_t n, uint8_t* out) { /* v5_codec.cpp:423-432 */
  size_t bytes = 1;
  int64_t prev = 0;
  out[0] = 0;
  for (size_t i = 0; i < n; ++i) {
    bytes += put_varint((int64_t)((uint64_t)a->v[i] - (uint64_t)prev), out + bytes);
    prev = a->v[i];
  }
  return bytes;
}
/* selectBestAdaptiveIntMode over the first `n` values (v5_codec.cpp:387-421): strict '<', order Delta, Palette, Rle, DeltaRle */
static int choose_mode(const vals_t* a, size_t n, int bpv) {
  size_t best = delta_section_size(a, n);
  int mode = 0;
  const size_t pal = palette_section(a, n, bpv, NULL);
  if (pal < best) { best = pal; mode = 1; }
  const size_t rle = rle_section(a, n, bpv, NULL);
  if (rle < best) { best = rle; mode = 2; }
  const size_t drl = delta_rle_section(a, n, NULL);
  if (drl < best) { mode = 3; }
  return mode;
}

/* ---------------------------------------------------------------- public: header, sizing */
static const char* tname(uint8_t t) {
  static const char* n[] = {"UNKNOWN", "INT8", "UINT8", "INT16", "UINT16", "INT32", "UINT32", "FLOAT32", "FLOAT64", "INT64", "UINT64"};
  return t <= 10 ? n[t] : "UNKNOWN";
}
/* EncodingInfoToYAML, cloudini.cpp:165-190 */
size_t orc_info_to_yaml(const cldn_info_t* in, char* out, size_t cap) {
  size_t n = 0;
#define APP(...) do { int w__ = snprintf(out ? out + n : NULL, out && cap > n ? cap - n : 0, __VA_ARGS__); n += (size_t)w__; } while (0)
  APP("version: %d\n", (int)in->version);
  APP("width: %u\n", in->width);
  APP("height: %u\n", in->height);
  APP("point_step: %u\n", in->point_step);
  APP("encoding_opt: %s\n", in->encoding_opt == 0 ? "NONE" : in->encoding_opt == 1 ? "LOSSY" : in->encoding_opt == 2 ? "LOSSLESS" : "UNKNOWN");
  APP("compression_opt: %s\n", in->compression_opt == 0 ? "NONE" : in->compression_opt == 1 ? "LZ4" : in->compression_opt == 2 ? "ZSTD" : "UNKNOWN");
  if (in->encoding_config[0]) APP("encoding_config: %s\n", in->encoding_config);
  APP("fields:\n");
  for (uint32_t i = 0; i < in->n_fields; ++i) {
    const cldn_field_t* f = &in->fields[i];
    APP("  - name: %s\n", f->name);
    APP("    offset: %u\n", f->offset);
    APP("    type: %s\n", tname(f->type));
    if (f->has_resolution) APP("    resolution: %g\n", (double)f->resolution); /* ostream << float */
    else APP("    resolution: null\n");
  }
#undef APP
  return n;
}
/* EncodeHeader (YAML), cloudini.cpp:294-318 */
size_t orc_header(const cldn_info_t* in, uint8_t* out, size_t cap) {
  const size_t y = orc_info_to_yaml(in, NULL, 0);
  const size_t total = 10 + 2 + 1 + y + 1;
  if (!out || cap < total) return total;
  memcpy(out, "CLOUDINI_V", 10);
  out[10] = (uint8_t)('0' + in->version / 10);
  out[11] = (uint8_t)('0' + in->version % 10);
  out[12] = '\n';
  orc_info_to_yaml(in, (char*)out + 13, y + 1);
  out[13 + y] = 0;
  return total;
}
static size_t max_field(const cldn_field_t* f, uint8_t enc) { /* codec_common.cpp:29-59 */
  switch (f->type) {
    case CLDN_INT16: case CLDN_UINT16: case CLDN_INT32: case CLDN_UINT32: case CLDN_INT64: case CLDN_UINT64: return 10;
    case CLDN_FLOAT32: return (enc == CLDN_ENC_LOSSY && f->has_resolution) ? 10 : 7;
    case CLDN_FLOAT64: return (enc == CLDN_ENC_LOSSY && f->has_resolution) ? 10 : 11;
    case CLDN_INT8: case CLDN_UINT8: return 1;
    default: return 0;
  }
}
/* MaxCompressedSize (compression NONE / LZ4 / ZSTD bounds), cloudini.cpp:249-292 */
size_t orc_max_compressed_size(const cldn_info_t* in, size_t points, int include_header) {
  if (in->point_step == 0) { fail("point_step cannot be 0"); return 0; }
  size_t per_point = 0;
  for (uint32_t i = 0; i < in->n_fields; ++i) per_point += max_field(&in->fields[i], in->encoding_opt);
  size_t total = include_header ? orc_header(in, NULL, 0) : 0;
  const int v5 = uses_v5(in);
  size_t left = points;
  while (left > 0) {
    const size_t n = left < ORC_CHUNK ? left : ORC_CHUNK;
    left -= n;
    size_t c = n * per_point;
    if (v5) c += in->n_fields * 32u + 1024u;
    total += 4;
    if (in->compression_opt == CLDN_COMP_NONE) total += c;
    else if (in->compression_opt == CLDN_COMP_LZ4) total += c + c / 255 + 16;
    else total += c + (c >> 8) + ((c < (128u << 10)) ? (((128u << 10) - c) >> 11) : 0);
  }
  return total;
}

/* ---------------------------------------------------------------- public: encode */
/* PointcloudEncoder::encode (compression NONE), cloudini.cpp:501-623. Returns bytes written or -1. */
long long orc_encode(const cldn_info_t* in, const uint8_t* cloud, size_t cloud_bytes, uint8_t* out, size_t cap,
                     int write_header) {
  if (in->point_step == 0) return fail("point_step cannot be 0");
  if (cloud_bytes % in->point_step) return fail("Input cloud_data size is not a multiple of point_step");
  if (in->compression_opt != CLDN_COMP_NONE) return fail("oracle: stage 2 is outside the restated path");
  const size_t points = cloud_bytes / in->point_step;
  const size_t hdr = write_header ? orc_header(in, NULL, 0) : 0;
  if (cap < orc_max_compressed_size(in, points, 0) + hdr) return fail("Output buffer too small for worst-case compressed size");
  plan_t pl;
  if (make_plan(in, 0, &pl)) return -1;
  size_t pos = 0;
  if (write_header) pos += orc_header(in, out, cap);
  vals_t vals[CLDN_MAX_FIELDS];
  for (int s = 0; s < pl.n_secs; ++s) {
    vals[s].v = (int64_t*)malloc(ORC_CHUNK * sizeof(int64_t));
    vals[s].raw = (uint64_t*)malloc(ORC_CHUNK * sizeof(uint64_t));
  }
  /* the reference serialises a chunk into its own stage-1 buffer and only then copies it behind the u32 prefix,
   * failing if it does not fit (chunk_writer.cpp:27-40): a committed section mode may exceed MaxCompressedSize's budget
   * (DeltaRle on a 64-bit field: 11 bytes per value against the 10 budgeted) */
  uint8_t* chunk_buf = (uint8_t*)malloc((size_t)ORC_CHUNK * ((size_t)in->point_step * 2 + 16 + (size_t)pl.n_secs * 12) + 4096);
  size_t done = 0;
  while (done < points) {
    const size_t n = (points - done) < ORC_CHUNK ? (points - done) : ORC_CHUNK;
    uint8_t* body = chunk_buf;
    size_t b = 0;
    reset_ops(&pl); /* v4_codec.cpp:69 / v5_codec.cpp:910-912 */
    for (size_t i = 0; i < n; ++i) {
      const uint8_t* pt = cloud + (done + i) * in->point_step;
      b += encode_point(&pl, pt, body + b);
      for (int s = 0; s < pl.n_secs; ++s) {
        vals[s].v[i] = read_int(pt + pl.secs[s].off, pl.secs[s].type);
        vals[s].raw[i] = read_raw(pt + pl.secs[s].off, pl.secs[s].bpv);
      }
    }
    for (int s = 0; s < pl.n_secs; ++s) {
      sec_t* sc = &pl.secs[s];
      if (!sc->committed) { /* first chunk: probe the first 4096 values, or the whole chunk if it is not larger (v5_codec.cpp:934-949) */
        const size_t probe = n > ORC_PROBE ? ORC_PROBE : n;
        sc->mode = choose_mode(&vals[s], probe, sc->bpv);
        sc->committed = 1;
      }
      switch (sc->mode) {
        case 0: b += delta_section(&vals[s], n, body + b); break;
        case 1: b += palette_section(&vals[s], n, sc->bpv, body + b); break;
        case 2: b += rle_section(&vals[s], n, sc->bpv, body + b); break;
        default: b += delta_rle_section(&vals[s], n, body + b); break;
      }
    }
    const uint32_t sz = (uint32_t)b; /* chunk_writer.cpp:33-40 */
    if (cap - pos < 4 || cap - pos - 4 < b) {
      for (int s = 0; s < pl.n_secs; ++s) { free(vals[s].v); free(vals[s].raw); }
      free(chunk_buf);
      return fail("Output buffer too small for uncompressed chunk");
    }
    memcpy(out + pos, &sz, 4);
    memcpy(out + pos + 4, chunk_buf, b);
    pos += 4 + b;
    done += n;
  }
  for (int s = 0; s < pl.n_secs; ++s) { free(vals[s].v); free(vals[s].raw); }
  free(chunk_buf);
  return (long long)pos;
}

/* ---------------------------------------------------------------- public: decode */
static void store_low(uint8_t* dst, uint64_t v, int bytes) { memcpy(dst, &v, (size_t)bytes); }

static int decode_section(const sec_t* sc, const uint8_t** pp, size_t* avail, uint8_t* out, size_t step, size_t n) {
  const uint8_t* p = *pp;
  size_t a = *avail;
  if (a == 0) return fail("V5 adaptive int: missing mode byte");
  const uint8_t mode = *p++; --a;
  if (mode > 3) return fail("V5 adaptive int: unknown mode byte");
  if (mode == 0) {
    int64_t prev = 0;
    for (size_t i = 0; i < n; ++i) {
      int64_t d;
      const size_t c = get_varint(p, a, &d);
      if (!c) return -1;
      p += c; a -= c;
      prev = (int64_t)((uint64_t)prev + (uint64_t)d);
      store_low(out + i * step + sc->off, (uint64_t)prev, sc->bpv);
    }
  } else if (mode == 1) {
    if (a < 2) return fail("decode: not enough input data");
    uint16_t cnt;
    memcpy(&cnt, p, 2); p += 2; a -= 2;
    if (cnt == 0) return fail("V5 adaptive int: empty palette");
    if (a < (size_t)cnt * (size_t)sc->bpv) return fail("V5 adaptive int: truncated palette");
    const uint8_t* pal = p;
    p += (size_t)cnt * (size_t)sc->bpv; a -= (size_t)cnt * (size_t)sc->bpv;
    const unsigned bits = palette_bits(cnt);
    const size_t ib = ((size_t)bits * n + 7) / 8;
    if (a < ib) return fail("V5 adaptive int: truncated palette indexes");
    uint64_t scratch = 0;
    unsigned held = 0;
    const uint8_t* ip = p;
    for (size_t i = 0; i < n; ++i) {
      uint32_t k = 0;
      if (bits) {
        while (held < bits) { scratch |= ((uint64_t)(*ip++)) << held; held += 8; }
        k = (uint32_t)(scratch & ((1ull << bits) - 1));
        scratch >>= bits; held -= bits;
      }
      if (k >= cnt) return fail("V5 adaptive int: palette index out of range");
      store_low(out + i * step + sc->off, read_raw(pal + (size_t)k * (size_t)sc->bpv, sc->bpv), sc->bpv);
    }
    p += ib; a -= ib;
  } else {
    if (a < 4) return fail("decode: not enough input data");
    uint32_t runs;
    memcpy(&runs, p, 4); p += 4; a -= 4;
    size_t oi = 0;
    int64_t prev = 0;
    for (uint32_t r = 0; r < runs; ++r) {
      uint64_t raw = 0;
      int64_t diff = 0;
      if (mode == 2) {
        if (a < (size_t)sc->bpv) return fail("V5 adaptive int: truncated RLE value");
        raw = read_raw(p, sc->bpv);
        p += sc->bpv; a -= (size_t)sc->bpv;
      } else {
        const size_t c = get_varint(p, a, &diff);
        if (!c) return -1;
        p += c; a -= c;
      }
      uint64_t len;
      const size_t c2 = get_uvarint(p, a, &len);
      if (!c2) return -1;
      p += c2; a -= c2;
      if (len > n - oi) return fail("V5 adaptive int: run exceeds point count");
      for (uint64_t k = 0; k < len; ++k) {
        if (mode == 3) { prev = (int64_t)((uint64_t)prev + (uint64_t)diff); raw = (uint64_t)prev; }
        store_low(out + oi * step + sc->off, raw, sc->bpv);
        ++oi;
      }
    }
    if (oi != n) return fail("V5 adaptive int: run count does not fill chunk");
  }
  *pp = p;
  *avail = a;
  return 0;
}

/* PointcloudDecoder::decode for version >= 3, compression NONE (cloudini.cpp:635-684). payload excludes the header. */
int orc_decode(const cldn_info_t* in, const uint8_t* payload, size_t bytes, uint8_t* out, size_t out_bytes) {
  if (in->compression_opt != CLDN_COMP_NONE) return fail("oracle: stage 2 is outside the restated path");
  if (in->version < 3) return fail("oracle: version < 3 is outside the restated path");
  plan_t pl;
  if (make_plan(in, 1, &pl)) return -1;
  if (bytes >= 10 && memcmp(payload, "CLOUDINI_V", 10) == 0) return fail("compressed_data contains the header. You should use DecodeHeader first");
  size_t remaining = (size_t)in->width * in->height;
  size_t done = 0;
  const size_t step = in->point_step;
  while (bytes > 0) {
    if (remaining == 0) return fail("Encoded data contains more chunks than declared points");
    if (bytes < 4) return fail("decode: not enough input data");
    uint32_t csz;
    memcpy(&csz, payload, 4);
    payload += 4; bytes -= 4;
    if (csz > bytes) return fail("Invalid chunk size found while decoding");
    const size_t n = remaining < ORC_CHUNK ? remaining : ORC_CHUNK;
    const uint8_t* p = payload;
    size_t a = csz;
    if ((done + n) * step > out_bytes) return fail("Output buffer is too small to hold the decoded data");
    uint8_t* chunk_out = out + done * step;
    reset_ops(&pl);
    for (size_t i = 0; i < n; ++i) {
      if (!pl.v5 && a < pl.min_point_bytes) return fail("Truncated encoded data: not enough bytes for a complete point");
      uint8_t* pt = chunk_out + i * step;
      for (int k = 0; k < pl.n_ops; ++k) {
        op_t* o = &pl.ops[k];
        if (o->kind == K_COPY) { /* field_decoder.hpp:66-72 */
          if (a < (size_t)o->size) return fail("Span: trim_front out of range");
          if (o->off[0] != CLDN_SKIP_STORE_OFFSET) memcpy(pt + o->off[0], p, (size_t)o->size);
          p += o->size; a -= (size_t)o->size;
          continue;
        }
        if (o->kind == K_XOR32 || o->kind == K_XOR64) { /* field_decoder.hpp:356-370 */
          if (a < (size_t)o->size) return fail("Span: trim_front out of range");
          uint64_t res = 0;
          memcpy(&res, p, (size_t)o->size);
          p += o->size; a -= (size_t)o->size;
          o->prev_bits ^= res;
          if (o->off[0] != CLDN_SKIP_STORE_OFFSET) memcpy(pt + o->off[0], &o->prev_bits, (size_t)o->size);
          continue;
        }
        if (o->kind == K_GORILLA64) {
          uint64_t v;
          const size_t c = gorilla_decode(o, p, a, &v);
          if (!c) return -1;
          p += c; a -= c;
          if (o->off[0] != CLDN_SKIP_STORE_OFFSET) memcpy(pt + o->off[0], &v, 8);
          continue;
        }
        for (int l = 0; l < o->lanes; ++l) {
          if (a == 0) return fail("decode: truncated input");
          const int is_float = o->kind != K_INT;
          if (is_float && p[0] == 0) { /* NaN marker */
            ++p; --a;
            if (o->kind == K_FLOATN) o->prev32[l] = 0; else o->prev64 = 0;
            if (o->off[l] != CLDN_SKIP_STORE_OFFSET) {
              if (o->kind == K_F64) { const uint64_t qn = 0x7FF8000000000000ull; memcpy(pt + o->off[l], &qn, 8); }
              else { const uint32_t qn = 0x7FC00000u; memcpy(pt + o->off[l], &qn, 4); }
            }
            continue;
          }
          int64_t d;
          const size_t c = get_varint(p, a, &d);
          if (!c) return -1;
          p += c; a -= c;
          if (o->kind == K_FLOATN) { /* field_decoder.cpp:62-70 */
            o->prev32[l] = (int32_t)((uint32_t)(int32_t)d + (uint32_t)o->prev32[l]);
            const float f = (float)o->prev32[l] * o->dec_mul[l];
            if (o->off[l] != CLDN_SKIP_STORE_OFFSET) memcpy(pt + o->off[l], &f, 4);
          } else {
            o->prev64 = (int64_t)((uint64_t)o->prev64 + (uint64_t)d);
            if (o->off[0] == CLDN_SKIP_STORE_OFFSET) continue;
            if (o->kind == K_F32) { const float f = (float)o->prev64 * o->dec_mul[0]; memcpy(pt + o->off[0], &f, 4); }
            else if (o->kind == K_F64) { const double f = (double)o->prev64 * o->dec_mul_d; memcpy(pt + o->off[0], &f, 8); }
            else store_low(pt + o->off[0], (uint64_t)o->prev64, o->size);
          }
        }
      }
    }
    for (int s = 0; s < pl.n_secs; ++s) {
      if (decode_section(&pl.secs[s], &p, &a, chunk_out, step, n)) return -1;
    }
    if (pl.v5 && a != 0) return fail("V5 chunk has trailing bytes after decode");
    payload += csz; bytes -= csz;
    remaining -= n;
    done += n;
  }
  if (remaining != 0) return fail("Encoded data ended before all declared points were decoded");
  return 0;
}

/* ---------------------------------------------------------------- timing helper for bench.py ("port" baseline) */
#include <time.h>
double orc_time_encode(const cldn_info_t* in, const uint8_t* cloud, size_t cloud_bytes, uint8_t* out, size_t cap, int reps,
                       long long* encoded) {
  struct timespec t0, t1;
  long long n = 0;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int r = 0; r < reps; ++r) n = orc_encode(in, cloud, cloud_bytes, out, cap, 1);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  if (encoded) *encoded = n;
  return n < 0 ? -1.0 : (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
double orc_time_decode(const cldn_info_t* in, const uint8_t* payload, size_t bytes, uint8_t* out, size_t out_bytes, int reps) {
  struct timespec t0, t1;
  int rc = 0;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int r = 0; r < reps; ++r) rc |= orc_decode(in, payload, bytes, out, out_bytes);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  return rc ? -1.0 : (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

/* ---- applyVizLossyPreprocessing (src/ros_msg_utils.cpp:249-341), restated --------------------------
 * Sequential: for every point in order, drop it when x, y or z is not finite (:311-313); quantise with
 * lround(v * (1/res)) truncated to int32 (:314-317), pack 21 bits per axis with a 2^20 bias (packVoxelKey21, :42-49), and
 * keep the point only if the key was not seen before (:318-320). The reference uses ankerl::unordered_dense::set; any
 * exact set gives the same survivors, so a plain open-addressing table is used here.
 * Returns the number of survivors (their bytes are written to `out`), or -1 when the call is one of the reference's
 * no-ops (no geometry triple, bad resolution, empty cloud: :250-278) — then nothing is written and `info` is untouched.
 * On success info->width / height and the FLOAT64 resolutions are updated like pc_info (:327-340). */
static uint64_t orc_voxel_key(int32_t qx, int32_t qy, int32_t qz) {
  const uint64_t mask = (1ull << 21) - 1ull;
  const int64_t bias = 1ll << 20;
  const uint64_t ux = (uint64_t)((int64_t)qx + bias) & mask;
  const uint64_t uy = (uint64_t)((int64_t)qy + bias) & mask;
  const uint64_t uz = (uint64_t)((int64_t)qz + bias) & mask;
  return ux | (uy << 21) | (uz << 42);
}
long long orc_viz_preprocess(cldn_info_t* info, const uint8_t* cloud, size_t cloud_bytes, uint8_t* out, size_t out_capacity) {
  if (info->n_fields < 3 || info->point_step == 0) return -1;
  const cldn_field_t* f0 = &info->fields[0];
  const cldn_field_t* f1 = &info->fields[1];
  const cldn_field_t* f2 = &info->fields[2];
  if (!(f0->type == CLDN_FLOAT32 && f1->type == CLDN_FLOAT32 && f2->type == CLDN_FLOAT32 && f0->has_resolution &&
        f1->has_resolution && f2->has_resolution && f0->resolution == f1->resolution && f0->resolution == f2->resolution &&
        f1->offset == f0->offset + 4u && f2->offset == f0->offset + 8u)) {
    return -1;
  }
  const float res = f0->resolution;
  if (!(res > 0.0f) || !isfinite(res)) return -1;
  const float inv_res = 1.0f / res;
  const size_t step = info->point_step, n_in = cloud_bytes / step;
  if (n_in == 0) return -1;
  size_t cap = 16;
  while (cap < 2 * n_in) cap <<= 1;
  uint64_t* table = (uint64_t*)malloc(cap * sizeof(uint64_t));
  if (!table) return -2;
  memset(table, 0xFF, cap * sizeof(uint64_t)); /* keys use 63 bits: all-ones is free */
  size_t kept = 0;
  for (size_t i = 0; i < n_in; ++i) {
    const uint8_t* p = cloud + i * step;
    float v[3];
    memcpy(&v[0], p + f0->offset, 4);
    memcpy(&v[1], p + f1->offset, 4);
    memcpy(&v[2], p + f2->offset, 4);
    if (!isfinite(v[0]) || !isfinite(v[1]) || !isfinite(v[2])) continue;
    const uint64_t key = orc_voxel_key((int32_t)lroundf(v[0] * inv_res), (int32_t)lroundf(v[1] * inv_res), (int32_t)lroundf(v[2] * inv_res));
    uint64_t h = key * 0x9E3779B97F4A7C15ull;
    h ^= h >> 29;
    size_t s = (size_t)h & (cap - 1);
    int seen = 0;
    while (table[s] != ~0ull) {
      if (table[s] == key) { seen = 1; break; }
      s = (s + 1) & (cap - 1);
    }
    if (seen) continue;
    table[s] = key;
    if ((kept + 1) * step > out_capacity) { free(table); return -2; }
    memcpy(out + kept * step, p, step);
    ++kept;
  }
  free(table);
  info->width = (uint32_t)kept;
  info->height = 1;
  for (uint32_t i = 0; i < info->n_fields; ++i) {
    if (info->fields[i].type == CLDN_FLOAT64 && !info->fields[i].has_resolution) {
      info->fields[i].has_resolution = 1;
      info->fields[i].resolution = 1e-6f;
    }
  }
  return (long long)kept;
}
