// C-ABI layer: handles, device memory, stream plumbing and kernel launches. See include/cloudini_b200.h.
// There is deliberately no CPU implementation of the codec here: if CUDA is unavailable every compute entry point
// fails with CLDN_ERR_CUDA.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <stdexcept>
#include <utility>
#include <string>
#include <vector>

#include "cldn_device.cuh"
#include "cldn_kernels.h"
#include "cldn_plan.h"

namespace cldn {
std::string info_to_yaml(const cldn_info_t& info);
int info_from_yaml(const char* yaml, size_t len, cldn_info_t* info);
size_t max_compressed_size(const cldn_info_t& info, size_t points, bool include_header, bool* ok);
std::vector<uint8_t> make_header(const cldn_info_t& info);
}  // namespace cldn

using namespace cldn;

#define CUDA_TRY(expr)                                                                          \
  do {                                                                                          \
    cudaError_t e__ = (expr);                                                                   \
    if (e__ != cudaSuccess) {                                                                   \
      set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(e__), __FILE__, __LINE__, cudaGetErrorString(e__)); \
      return CLDN_ERR_CUDA;                                                                     \
    }                                                                                           \
  } while (0)

namespace {

// Grow-only device / pinned-host buffers.
template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;
  int reserve(size_t n, bool zero = false) {
    if (n <= cap) return CLDN_OK;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = std::max<size_t>(n, 16);
    CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&p), want * sizeof(T)));
    cap = want;
    if (zero) {
      // cudaMemset runs on the legacy stream and is asynchronous to the host; the handles' streams are non-blocking and
      // would not order their kernels behind it, so the zeroes are made final here (growth is rare)
      CUDA_TRY(cudaMemset(p, 0, want * sizeof(T)));
      CUDA_TRY(cudaDeviceSynchronize());
    }
    return CLDN_OK;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
template <typename T>
struct PinBuf {
  T* p = nullptr;
  size_t cap = 0;
  int reserve(size_t n) {
    if (n <= cap) return CLDN_OK;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    size_t want = std::max<size_t>(n, 16);
    CUDA_TRY(cudaMallocHost(reinterpret_cast<void**>(&p), want * sizeof(T)));
    cap = want;
    return CLDN_OK;
  }
  void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

// Ring of pinned upload slots: a slot is only rewritten by the host after the async copy that read it has completed
// (back-to-back asynchronous batch calls would otherwise race on a single pinned table).
template <typename T, int SLOTS = 8>
struct PinRing {
  PinBuf<T> slot[SLOTS];
  cudaEvent_t ev[SLOTS] = {};
  bool used[SLOTS] = {};
  int next = 0;
  // returns a host pointer with room for n elements
  int acquire(size_t n, T** out, int* idx) {
    const int i = next;
    next = (next + 1) % SLOTS;
    if (used[i]) CUDA_TRY(cudaEventSynchronize(ev[i]));
    if (!ev[i]) CUDA_TRY(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
    if (int rc = slot[i].reserve(n)) return rc;
    *out = slot[i].p;
    *idx = i;
    return CLDN_OK;
  }
  int commit(int idx, cudaStream_t stream) {
    CUDA_TRY(cudaEventRecord(ev[idx], stream));
    used[idx] = true;
    return CLDN_OK;
  }
  void release() {
    for (int i = 0; i < SLOTS; ++i) {
      slot[i].release();
      if (ev[i]) cudaEventDestroy(ev[i]);
      ev[i] = nullptr;
      used[i] = false;
    }
  }
};

// Extra streams + an event pool for the host-pointer path: frame f+1 is uploaded while frame f is encoded and
// frame f-1 is downloaded (PCIe is full duplex), instead of one serial copy-compute-copy sequence.
struct CopyPipeline {
  cudaStream_t h2d = nullptr, d2h = nullptr;
  std::vector<cudaEvent_t> ev;
  int ensure(size_t n_events) {
    if (!h2d) CUDA_TRY(cudaStreamCreateWithFlags(&h2d, cudaStreamNonBlocking));
    if (!d2h) CUDA_TRY(cudaStreamCreateWithFlags(&d2h, cudaStreamNonBlocking));
    while (ev.size() < n_events) {
      cudaEvent_t e;
      CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      ev.push_back(e);
    }
    return CLDN_OK;
  }
  void release() {
    for (cudaEvent_t e : ev) cudaEventDestroy(e);
    ev.clear();
    if (h2d) cudaStreamDestroy(h2d);
    if (d2h) cudaStreamDestroy(d2h);
    h2d = d2h = nullptr;
  }
};

// ---- stage 2 (general-purpose compression of every chunk) ---------------------------------------------------------
// Not the product (north_star: delegated / bypassed) and not part of the GPU path: when a caller asks for LZ4 / ZSTD
// through the HOST-pointer API, the stage-1 bytes produced by the kernels are handed chunk by chunk to the system's
// liblz4 / libzstd (dlopen, like the reference links them: codec_common.cpp:220-299). Compressed bytes are not
// expected to equal the reference's (different library versions); they decompress to the same stage-1 bytes.
struct Stage2 {
  void* lz4 = nullptr;
  void* zstd = nullptr;
  int (*lz4_compress)(const char*, char*, int, int) = nullptr;
  int (*lz4_decompress)(const char*, char*, int, int) = nullptr;
  size_t (*zstd_compress)(void*, size_t, const void*, size_t, int) = nullptr;
  size_t (*zstd_decompress)(void*, size_t, const void*, size_t) = nullptr;
  unsigned (*zstd_is_error)(size_t) = nullptr;
  bool load(int option) {
    if (option == CLDN_COMP_LZ4) {
      if (!lz4) {
        lz4 = dlopen("liblz4.so.1", RTLD_NOW);
        if (!lz4) lz4 = dlopen("liblz4.so", RTLD_NOW);
        if (lz4) {
          lz4_compress = reinterpret_cast<int (*)(const char*, char*, int, int)>(dlsym(lz4, "LZ4_compress_default"));
          lz4_decompress = reinterpret_cast<int (*)(const char*, char*, int, int)>(dlsym(lz4, "LZ4_decompress_safe"));
        }
      }
      if (!lz4_compress || !lz4_decompress) { set_error("stage 2: liblz4 is not available on this host"); return false; }
      return true;
    }
    if (!zstd) {
      zstd = dlopen("libzstd.so.1", RTLD_NOW);
      if (!zstd) zstd = dlopen("libzstd.so", RTLD_NOW);
      if (zstd) {
        zstd_compress = reinterpret_cast<size_t (*)(void*, size_t, const void*, size_t, int)>(dlsym(zstd, "ZSTD_compress"));
        zstd_decompress = reinterpret_cast<size_t (*)(void*, size_t, const void*, size_t)>(dlsym(zstd, "ZSTD_decompress"));
        zstd_is_error = reinterpret_cast<unsigned (*)(size_t)>(dlsym(zstd, "ZSTD_isError"));
      }
    }
    if (!zstd_compress || !zstd_decompress || !zstd_is_error) { set_error("stage 2: libzstd is not available on this host"); return false; }
    return true;
  }
};
static Stage2 g_stage2;

// Re-frames a stage-1 payload ([u32 size][body])* into ([u32 csize][compressed body])* — WriteStage1Chunk, chunk_writer.cpp:42-47.
static int stage2_compress_payload(int option, const uint8_t* src, size_t src_bytes, uint8_t* dst, size_t dst_cap, size_t* written) {
  size_t ip = 0, op = 0;
  while (ip < src_bytes) {
    uint32_t sz;
    memcpy(&sz, src + ip, 4);
    ip += 4;
    if (dst_cap - op < 4) { set_error("Output buffer too small for compressed chunk"); return CLDN_ERR_BUFFER_TOO_SMALL; }
    uint32_t csz = 0;
    if (option == CLDN_COMP_LZ4) {  // codec_common.cpp:232-237
      const int c = g_stage2.lz4_compress(reinterpret_cast<const char*>(src + ip), reinterpret_cast<char*>(dst + op + 4), static_cast<int>(sz),
                                          static_cast<int>(std::min<size_t>(dst_cap - op - 4, 0x7FFFFFFF)));
      if (c <= 0) { set_error("LZ4 compression failed"); return CLDN_ERR_BUFFER_TOO_SMALL; }
      csz = static_cast<uint32_t>(c);
    } else {                        // ZSTD level 1, codec_common.cpp:242
      const size_t c = g_stage2.zstd_compress(dst + op + 4, dst_cap - op - 4, src + ip, sz, 1);
      if (g_stage2.zstd_is_error(c)) { set_error("ZSTD compression failed"); return CLDN_ERR_BUFFER_TOO_SMALL; }
      csz = static_cast<uint32_t>(c);
    }
    memcpy(dst + op, &csz, 4);
    op += 4 + csz;
    ip += sz;
  }
  *written = op;
  return CLDN_OK;
}

// Inverse: ([u32 csize][compressed])* -> ([u32 size][stage-1 body])*; `max_body` bounds a decompressed chunk.
static int stage2_decompress_payload(int option, const uint8_t* src, size_t src_bytes, std::vector<uint8_t>& dst, size_t max_body) {
  size_t ip = 0;
  dst.clear();
  while (ip < src_bytes) {
    if (src_bytes - ip < 4) { set_error("decode: not enough input data"); return CLDN_ERR_CORRUPT_DATA; }
    uint32_t csz;
    memcpy(&csz, src + ip, 4);
    ip += 4;
    if (csz > src_bytes - ip) { set_error("Invalid chunk size found while decoding"); return CLDN_ERR_CORRUPT_DATA; }
    const size_t at = dst.size();
    dst.resize(at + 4 + max_body);
    uint32_t sz = 0;
    if (option == CLDN_COMP_LZ4) {
      const int r = g_stage2.lz4_decompress(reinterpret_cast<const char*>(src + ip), reinterpret_cast<char*>(dst.data() + at + 4), static_cast<int>(csz),
                                            static_cast<int>(std::min<size_t>(max_body, 0x7FFFFFFF)));
      if (r < 0) { set_error("LZ4 decompression failed"); return CLDN_ERR_CORRUPT_DATA; }
      sz = static_cast<uint32_t>(r);
    } else {
      const size_t r = g_stage2.zstd_decompress(dst.data() + at + 4, max_body, src + ip, csz);
      if (g_stage2.zstd_is_error(r)) { set_error("ZSTD decompression failed"); return CLDN_ERR_CORRUPT_DATA; }
      sz = static_cast<uint32_t>(r);
    }
    memcpy(dst.data() + at, &sz, 4);
    dst.resize(at + 4 + sz);
    ip += csz;
  }
  return CLDN_OK;
}

int select_device(int device) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    set_error("no CUDA device available (%s): cloudini_b200 has no CPU fallback", e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    return CLDN_ERR_CUDA;
  }
  if (device >= 0) CUDA_TRY(cudaSetDevice(device));
  return CLDN_OK;
}

const char* dev_error_text(uint32_t code) {
  switch (code) {
    case DEV_ERR_TRUNCATED: return "Truncated encoded data: not enough bytes for a complete point";
    case DEV_ERR_VARINT_OVERFLOW: return "decodeVarint: value overflow";
    case DEV_ERR_NAN_MARKER: return "decodeVarint: unexpected NaN marker";
    case DEV_ERR_TRAILING: return "V5 chunk has trailing bytes after decode";
    case DEV_ERR_BAD_MODE: return "V5 adaptive int: missing or unknown mode byte";
    case DEV_ERR_PALETTE: return "V5 adaptive int: bad palette section";
    case DEV_ERR_RLE: return "V5 adaptive int: bad run-length section";
    case DEV_ERR_CHUNK_SIZE: return "Invalid chunk size found while decoding";
    case DEV_ERR_CHUNK_COUNT: return "Encoded data does not match the declared number of points (chunk count)";
    case DEV_ERR_OUTPUT_SMALL: return "Output buffer is too small to hold the decoded data";
    case DEV_ERR_ENCODE_OUTPUT_SMALL: return "Output buffer too small for uncompressed chunk";  // chunk_writer.cpp:33-35
    case DEV_ERR_LZ4: return "LZ4 decompression failed";  // codec_common.cpp:279-281
    default: return "unknown device error";
  }
}

}  // namespace

// =====================================================================================================================
struct cldn_encoder {
  cldn_info_t info;       // as given, except compression_opt = NONE (what the GPU path produces)
  int stage2 = 0;         // the caller's compression_opt (host-pointer API only)
  std::vector<uint8_t> s2_tmp;
  Plan plan;
  std::vector<uint8_t> header;
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  uint32_t tile_points = 0;
  uint32_t epoch = 0;

  DevBuf<Plan> d_plan;
  DevBuf<uint8_t> d_header;
  DevBuf<uint64_t> d_status;
  DevBuf<EncFrame> d_frames;
  PinRing<EncFrame> h_frames;
  DevBuf<uint64_t> d_sizes;
  PinBuf<uint64_t> h_sizes;
  DevBuf<uint32_t> d_err;
  PinBuf<uint32_t> h_err;
  // host-memory path staging
  DevBuf<uint8_t> d_in, d_out;
  CopyPipeline pipe;
  // V5 section scratch
  DevBuf<uint8_t> d_modes;
  DevBuf<uint8_t> d_sec_scratch;
  DevBuf<uint32_t> d_sec_sizes;
  DevBuf<uint32_t> d_sec_excl;
  DevBuf<uint32_t> d_chunk_frame;
  PinRing<uint32_t> h_chunk_frame;
  DevBuf<uint64_t> d_hash;
  DevBuf<uint8_t> d_gorilla;  // Gorilla pre-pass records: [frame][op][point][12]
  size_t last_frames = 0;
  // stage 2 on the device (LZ4): stage-1 arena, per-chunk slots, chunk tables
  DevBuf<uint8_t> d_s1, d_lz_scratch;
  DevBuf<uint32_t> d_lz_chunk_frame, d_lz_chunk_sizes;
  DevBuf<Lz4Frame> d_lz_frames;
  PinRing<Lz4Frame> h_lz_frames;
  PinRing<uint32_t> h_lz_chunk_frame;
};

// LZ4_COMPRESSBOUND (lz4.h): worst case of a block for n input bytes
static size_t lz4_bound(size_t n) { return n + n / 255 + 16; }

extern "C" {

uint64_t cldn_b200_kernel_launch_count(void) { return kernel_launch_count(); }

int cldn_b200_bind_host_thread_to_device(int device) {
  if (int rc = select_device(device)) return rc;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  char bus_id[64] = {0};
  if (cudaDeviceGetPCIBusId(bus_id, sizeof(bus_id), dev) != cudaSuccess) return 0;
  // NVML through dlopen: the library has no link-time dependency on it
  static void* nvml = dlopen("libnvidia-ml.so.1", RTLD_NOW);
  if (!nvml) return 0;
  using InitFn = int (*)();
  using HandleFn = int (*)(const char*, void**);
  using AffFn = int (*)(void*, unsigned int, unsigned long*);
  InitFn init = reinterpret_cast<InitFn>(dlsym(nvml, "nvmlInit_v2"));
  HandleFn by_bus = reinterpret_cast<HandleFn>(dlsym(nvml, "nvmlDeviceGetHandleByPciBusId_v2"));
  AffFn affinity = reinterpret_cast<AffFn>(dlsym(nvml, "nvmlDeviceGetCpuAffinity"));
  if (!init || !by_bus || !affinity || init() != 0) return 0;
  void* handle = nullptr;
  if (by_bus(bus_id, &handle) != 0) return 0;
  constexpr unsigned kWords = 16;  // 1024 CPUs
  unsigned long words[kWords] = {0};
  if (affinity(handle, kWords, words) != 0) return 0;
  cpu_set_t set;
  CPU_ZERO(&set);
  int count = 0;
  for (unsigned w = 0; w < kWords; ++w) {
    for (unsigned b = 0; b < 8 * sizeof(unsigned long); ++b) {
      if ((words[w] >> b) & 1ul) { CPU_SET(w * 8 * sizeof(unsigned long) + b, &set); ++count; }
    }
  }
  if (count == 0 || sched_setaffinity(0, sizeof(set), &set) != 0) return 0;
  return count;
}

int cldn_b200_encoder_create(const cldn_info_t* info, int device, void* stream, cldn_encoder_t** out) {
  if (!info || !out) { set_error("null argument"); return CLDN_ERR_INVALID_ARGUMENT; }
  *out = nullptr;
  Plan plan;
  if (int rc = build_encode_plan(*info, &plan)) return rc;
  if (!plan.supported) {
    set_error("EncodingInfo contains a field encoder this build cannot run");
    return CLDN_ERR_UNSUPPORTED;
  }
  if (info->compression_opt > CLDN_COMP_ZSTD) { set_error("Unsupported compression option"); return CLDN_ERR_INVALID_ARGUMENT; }
  // ZSTD exists only as the host library; LZ4 also runs on the device, so a missing liblz4 only matters to the host-pointer API
  if (info->compression_opt == CLDN_COMP_ZSTD && !g_stage2.load(info->compression_opt)) return CLDN_ERR_UNSUPPORTED;
  if (int rc = select_device(device)) return rc;
  cldn_encoder* e = new cldn_encoder();
  e->info = *info;
  e->stage2 = info->compression_opt;        // the header declares it; the kernels always produce stage 1
  e->info.compression_opt = CLDN_COMP_NONE;
  e->plan = plan;
  e->header = make_header(*info);
  cudaGetDevice(&e->device);
  e->tile_points = choose_tile_points(plan);
  if (stream) {
    e->stream = static_cast<cudaStream_t>(stream);
  } else {
    if (cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess) {
      set_error("cudaStreamCreate failed");
      delete e;
      return CLDN_ERR_CUDA;
    }
    e->own_stream = true;
  }
  int rc = e->d_plan.reserve(1);
  if (!rc) rc = e->d_header.reserve(e->header.size());
  if (!rc) rc = e->d_err.reserve(1, true);
  if (!rc) rc = e->h_err.reserve(1);
  if (!rc) {
    cudaError_t ce = cudaMemcpy(e->d_plan.p, &e->plan, sizeof(Plan), cudaMemcpyHostToDevice);
    if (ce == cudaSuccess) ce = cudaMemcpy(e->d_header.p, e->header.data(), e->header.size(), cudaMemcpyHostToDevice);
    if (ce != cudaSuccess) { set_error("cudaMemcpy failed: %s", cudaGetErrorString(ce)); rc = CLDN_ERR_CUDA; }
  }
  if (rc) { cldn_b200_encoder_destroy(e); return rc; }
  *out = e;
  return CLDN_OK;
}

void cldn_b200_encoder_destroy(cldn_encoder_t* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  if (e->stream) cudaStreamSynchronize(e->stream);
  e->d_plan.release(); e->d_header.release(); e->d_status.release(); e->d_frames.release(); e->h_frames.release();
  e->d_sizes.release(); e->h_sizes.release(); e->d_err.release(); e->h_err.release(); e->d_in.release(); e->d_out.release();
  e->d_modes.release(); e->d_sec_scratch.release(); e->d_sec_sizes.release(); e->d_sec_excl.release();
  e->d_chunk_frame.release(); e->h_chunk_frame.release(); e->d_hash.release(); e->d_gorilla.release(); e->pipe.release();
  e->d_s1.release(); e->d_lz_scratch.release(); e->d_lz_chunk_frame.release(); e->d_lz_chunk_sizes.release(); e->d_lz_frames.release();
  e->h_lz_frames.release(); e->h_lz_chunk_frame.release();
  if (e->own_stream && e->stream) cudaStreamDestroy(e->stream);
  delete e;
}

int cldn_b200_encoder_header(const cldn_encoder_t* enc, const uint8_t** header, size_t* header_bytes) {
  if (!enc) { set_error("null encoder"); return CLDN_ERR_INVALID_ARGUMENT; }
  if (header) *header = enc->header.data();
  if (header_bytes) *header_bytes = enc->header.size();
  return CLDN_OK;
}

int cldn_b200_encoder_info(const cldn_encoder_t* enc, cldn_info_t* info) {
  if (!enc || !info) { set_error("null argument"); return CLDN_ERR_INVALID_ARGUMENT; }
  *info = enc->info;
  info->compression_opt = static_cast<uint8_t>(enc->stage2);  // as the caller gave it
  return CLDN_OK;
}

const uint64_t* cldn_b200_encoder_sizes_device(const cldn_encoder_t* enc) { return enc ? enc->d_sizes.p : nullptr; }

int cldn_b200_encoder_set_dims(cldn_encoder_t* e, uint32_t width, uint32_t height) {
  if (!e) { set_error("null encoder"); return CLDN_ERR_INVALID_ARGUMENT; }
  if (e->info.width == width && e->info.height == height) return CLDN_OK;
  CUDA_TRY(cudaSetDevice(e->device));
  e->info.width = width;
  e->info.height = height;
  cldn_info_t full = e->info;
  full.compression_opt = static_cast<uint8_t>(e->stage2);
  const std::vector<uint8_t> h = make_header(full);
  // the device copy of the previous header may still be read by kernels in flight on the handle's stream
  CUDA_TRY(cudaStreamSynchronize(e->stream));
  if (int rc = e->d_header.reserve(h.size())) return rc;
  CUDA_TRY(cudaMemcpy(e->d_header.p, h.data(), h.size(), cudaMemcpyHostToDevice));
  e->header = h;
  return CLDN_OK;
}

static int check_device_error(cudaStream_t stream, uint32_t* d_err, uint32_t* h_err) {
  CUDA_TRY(cudaMemcpyAsync(h_err, d_err, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
  CUDA_TRY(cudaStreamSynchronize(stream));
  if (*h_err != DEV_OK) {
    const uint32_t code = *h_err;
    cudaMemsetAsync(d_err, 0, sizeof(uint32_t), stream);
    set_error("%s", dev_error_text(code));
    return code == DEV_ERR_ENCODE_OUTPUT_SMALL ? CLDN_ERR_BUFFER_TOO_SMALL : CLDN_ERR_CORRUPT_DATA;
  }
  return CLDN_OK;
}

int cldn_b200_encoder_sync(cldn_encoder_t* e) {
  if (!e) { set_error("null encoder"); return CLDN_ERR_INVALID_ARGUMENT; }
  CUDA_TRY(cudaSetDevice(e->device));
  return check_device_error(e->stream, e->d_err.p, e->h_err.p);
}

// Enqueues the kernels for a batch whose inputs/outputs are device pointers.
static int encode_batch_device(cldn_encoder* e, size_t n_frames, const void* const* clouds, const size_t* cloud_bytes,
                               void* const* outs, const size_t* out_capacities, int write_header) {
  const cldn_info_t& info = e->info;
  if (info.point_step == 0) { set_error("point_step cannot be 0"); return CLDN_ERR_INVALID_ARGUMENT; }  // cloudini.cpp:523-525
  EncFrame* hf = nullptr;
  int hf_slot = 0;
  if (int rc = e->h_frames.acquire(n_frames, &hf, &hf_slot)) return rc;
  if (int rc = e->d_frames.reserve(n_frames)) return rc;
  if (int rc = e->d_sizes.reserve(n_frames)) return rc;
  const uint32_t T = e->tile_points;
  const size_t hdr = write_header ? e->header.size() : 0;
  uint64_t tiles = 0, chunks = 0;
  bool aligned16 = true;
  for (size_t f = 0; f < n_frames; ++f) {
    if (cloud_bytes[f] % info.point_step != 0) {
      set_error("Input cloud_data size is not a multiple of point_step");  // cloudini.cpp:526-528
      return CLDN_ERR_INVALID_ARGUMENT;
    }
    const size_t n = cloud_bytes[f] / info.point_step;
    if (n > 0xFFFFFFFFull) { set_error("too many points in one frame"); return CLDN_ERR_UNSUPPORTED; }
    bool ok;
    const size_t need = max_compressed_size(info, n, false, &ok) + hdr;  // cloudini.cpp:530-534
    if (!ok) return CLDN_ERR_INVALID_ARGUMENT;
    if (out_capacities[f] < need) {
      set_error("Output buffer too small for worst-case compressed size");
      return CLDN_ERR_BUFFER_TOO_SMALL;
    }
    if (n > 0 && (!clouds[f] || !outs[f])) { set_error("null frame pointer"); return CLDN_ERR_INVALID_ARGUMENT; }
    EncFrame& F = hf[f];
    F.in = static_cast<const uint8_t*>(clouds[f]);
    F.out = static_cast<uint8_t*>(outs[f]);
    F.out_cap = out_capacities[f];
    F.n_points = static_cast<uint32_t>(n);
    F.tile_begin = static_cast<uint32_t>(tiles);
    F.n_tiles = static_cast<uint32_t>((n + T - 1) / T);
    F.n_chunks = static_cast<uint32_t>((n + kChunkPoints - 1) / kChunkPoints);
    F.sec_excl = nullptr;
    tiles += F.n_tiles;
    chunks += F.n_chunks;
    if (reinterpret_cast<uintptr_t>(clouds[f]) & 15u) aligned16 = false;
  }
  if (tiles > 0x7FFFFFFFull) { set_error("batch too large"); return CLDN_ERR_UNSUPPORTED; }
  for (size_t f = 0; f < n_frames; ++f) hf[f].side = nullptr;
  if (e->plan.n_gorilla > 0) {
    // FieldEncoderFloat_Gorilla keeps a leading/trailing-zero window across the points of a chunk: a pre-pass walks
    // every chunk in order and leaves one 12-byte record per (op, point) for the (parallel) regular kernel to copy
    size_t total_points = 0;
    for (size_t f = 0; f < n_frames; ++f) total_points += hf[f].n_points;
    if (int rc = e->d_gorilla.reserve(total_points * 12 * e->plan.n_gorilla + 16)) return rc;
    size_t at = 0;
    for (size_t f = 0; f < n_frames; ++f) {
      hf[f].side = e->d_gorilla.p + at;
      at += static_cast<size_t>(hf[f].n_points) * 12 * e->plan.n_gorilla;
    }
  }
  if (int rc = e->d_status.reserve(static_cast<size_t>(tiles) + 1, true)) return rc;
  e->epoch = (e->epoch + 1) & 0x3FFFFFu;
  if (e->epoch == 0) {  // epoch wrapped: stale words could alias
    CUDA_TRY(cudaMemsetAsync(e->d_status.p, 0, e->d_status.cap * sizeof(uint64_t), e->stream));
    e->epoch = 1;
  }
  EncLaunch L;
  L.frames = e->d_frames.p;
  L.n_frames = static_cast<uint32_t>(n_frames);
  L.n_tiles_total = static_cast<uint32_t>(tiles);
  L.plan = e->d_plan.p;
  L.header = e->d_header.p;
  L.header_bytes = static_cast<uint32_t>(hdr);
  L.status = e->d_status.p;
  L.epoch = e->epoch;
  L.sizes = e->d_sizes.p;
  L.err = e->d_err.p;
  L.tile_points = T;
  L.flags = aligned16 ? kEncInputsAligned16 : 0u;
  L.uniform_tiles = 0;
  if (n_frames > 0 && hf[0].n_tiles > 0) {
    bool uniform = true;
    for (size_t f = 1; f < n_frames; ++f) uniform = uniform && hf[f].n_tiles == hf[0].n_tiles;
    if (uniform) L.uniform_tiles = hf[0].n_tiles;
  }

  if (e->plan.n_sections > 0) {
    // V5: adaptive integer sections are produced first (they determine where every later chunk starts).
    const uint32_t ns = e->plan.n_sections;
    // worst case of any mode is DeltaRle with one run per value: 5 + n * (max varint of the field + 1) bytes
    uint32_t per_value = 0;
    for (uint32_t s = 0; s < ns; ++s) {
      const uint32_t bpv = e->plan.sections[s].bpv;
      per_value = std::max<uint32_t>(per_value, (bpv == 2 ? 3u : bpv == 4 ? 5u : 10u) + 1u);
    }
    const uint32_t stride = 32 + kChunkPoints * per_value;   // 5 header bytes + 16 bytes the vector copy of place_sections_kernel may read past the end
    if (int rc = e->d_modes.reserve(n_frames * ns)) return rc;
    if (int rc = e->d_sec_scratch.reserve(static_cast<size_t>(chunks) * ns * stride)) return rc;
    if (int rc = e->d_sec_sizes.reserve(static_cast<size_t>(chunks) * ns + 1)) return rc;
    if (int rc = e->d_sec_excl.reserve(static_cast<size_t>(chunks) + n_frames + 1)) return rc;
    uint32_t* hcf = nullptr;
    int hcf_slot = 0;
    if (int rc = e->h_chunk_frame.acquire(static_cast<size_t>(chunks) + n_frames + 1, &hcf, &hcf_slot)) return rc;
    if (int rc = e->d_chunk_frame.reserve(static_cast<size_t>(chunks) + n_frames + 1)) return rc;
    if (int rc = e->d_hash.reserve(palette_overflow_scratch_bytes() / 8 + 16)) return rc;
    uint32_t cb = 0;
    for (size_t f = 0; f < n_frames; ++f) {
      EncFrame& F = hf[f];
      F.sec_excl = e->d_sec_excl.p + cb + f;  // (n_chunks + 1) entries per frame
      for (uint32_t c = 0; c < F.n_chunks; ++c) hcf[cb + c] = static_cast<uint32_t>(f);
      hcf[chunks + f] = cb;  // chunk_first[f]
      cb += F.n_chunks;
    }
    CUDA_TRY(cudaMemcpyAsync(e->d_chunk_frame.p, hcf, (chunks + n_frames) * sizeof(uint32_t), cudaMemcpyHostToDevice, e->stream));
    if (int rc = e->h_chunk_frame.commit(hcf_slot, e->stream)) return rc;
    CUDA_TRY(cudaMemcpyAsync(e->d_frames.p, hf, n_frames * sizeof(EncFrame), cudaMemcpyHostToDevice, e->stream));
    if (int rc = e->h_frames.commit(hf_slot, e->stream)) return rc;
    SecLaunch S{};
    S.frames = e->d_frames.p;
    S.n_frames = static_cast<uint32_t>(n_frames);
    S.n_chunks_total = static_cast<uint32_t>(chunks);
    S.chunk_frame = e->d_chunk_frame.p;
    S.chunk_first = e->d_chunk_frame.p + chunks;
    S.plan = e->d_plan.p;
    S.modes = e->d_modes.p;
    S.scratch = e->d_sec_scratch.p;
    S.sec_stride = stride;
    S.sec_sizes = e->d_sec_sizes.p;
    S.sec_excl = e->d_sec_excl.p;
    S.hash_scratch = e->d_hash.p;
    S.err = e->d_err.p;
    S.header_bytes = static_cast<uint32_t>(hdr);
    if (launch_encode_sections(e->plan, S, e->stream) < 0) { set_error("section kernel launch failed"); return CLDN_ERR_CUDA; }
    if (launch_gorilla_prepass(e->plan, L, e->stream) < 0) { set_error("gorilla pre-pass launch failed"); return CLDN_ERR_CUDA; }
    if (launch_encode_regular(e->plan, L, e->stream) < 0) { set_error("encode kernel launch failed"); return CLDN_ERR_CUDA; }
    if (launch_place_sections(e->plan, S, e->d_status.p, e->epoch, T, e->stream) < 0) { set_error("section placement launch failed"); return CLDN_ERR_CUDA; }
  } else {
    CUDA_TRY(cudaMemcpyAsync(e->d_frames.p, hf, n_frames * sizeof(EncFrame), cudaMemcpyHostToDevice, e->stream));
    if (int rc = e->h_frames.commit(hf_slot, e->stream)) return rc;
    if (launch_gorilla_prepass(e->plan, L, e->stream) < 0) { set_error("gorilla pre-pass launch failed"); return CLDN_ERR_CUDA; }
    if (launch_encode_regular(e->plan, L, e->stream) < 0) { set_error("encode kernel launch failed"); return CLDN_ERR_CUDA; }
  }
  CUDA_TRY(cudaGetLastError());
  e->last_frames = n_frames;
  return CLDN_OK;
}

}  // extern "C"

// No C++ exception may cross the C ABI (std::terminate under ctypes / C callers): allocation failures become a status.
template <class Fn>
static int guarded(Fn&& fn) {
  try {
    return fn();
  } catch (const std::bad_alloc&) {
    set_error("out of memory");
    return CLDN_ERR_INTERNAL;
  } catch (const std::exception& ex) {
    set_error("internal error: %s", ex.what());
    return CLDN_ERR_INTERNAL;
  }
}

// Device pointers in, device pointers out, compression LZ4: stage 1 into an internal arena, one LZ4 block per chunk on the
// device (cldn_lz4.cu), blobs packed behind their headers. Nothing crosses PCIe except the 8-byte sizes, if asked for.
static int encode_batch_device_lz4(cldn_encoder* e, size_t n_frames, const void* const* clouds, const size_t* cloud_bytes,
                                   void* const* outs, const size_t* out_capacities, int write_header) {
  const cldn_info_t& info = e->info;  // compression NONE: what the stage-1 kernels produce
  if (info.point_step == 0) { set_error("point_step cannot be 0"); return CLDN_ERR_INVALID_ARGUMENT; }
  cldn_info_t full = info;
  full.compression_opt = CLDN_COMP_LZ4;
  const size_t hdr = write_header ? e->header.size() : 0;
  std::vector<size_t> s1_off(n_frames), s1_cap(n_frames);
  size_t arena = 0, chunks = 0, max_chunk = 0;
  for (size_t f = 0; f < n_frames; ++f) {
    if (cloud_bytes[f] % info.point_step != 0) { set_error("Input cloud_data size is not a multiple of point_step"); return CLDN_ERR_INVALID_ARGUMENT; }
    const size_t n = cloud_bytes[f] / info.point_step;
    bool ok;
    const size_t need = max_compressed_size(full, n, false, &ok) + hdr;  // cloudini.cpp:530-534
    if (!ok) return CLDN_ERR_INVALID_ARGUMENT;
    if (out_capacities[f] < need) { set_error("Output buffer too small for worst-case compressed size"); return CLDN_ERR_BUFFER_TOO_SMALL; }
    s1_cap[f] = max_compressed_size(info, n, false, &ok) + 64 * ((n + kChunkPoints - 1) / kChunkPoints + 1);
    s1_off[f] = arena;
    arena += (s1_cap[f] + 255) & ~size_t(255);
    chunks += (n + kChunkPoints - 1) / kChunkPoints;
    max_chunk = std::max(max_chunk, max_compressed_size(info, std::min<size_t>(n, kChunkPoints), false, &ok) + 64);
  }
  if (chunks > 0x7FFFFFFFull) { set_error("batch too large"); return CLDN_ERR_UNSUPPORTED; }
  if (int rc = e->d_s1.reserve(arena + 256)) return rc;
  std::vector<const void*> s1_ptr_c(n_frames);
  std::vector<void*> s1_ptr(n_frames);
  for (size_t f = 0; f < n_frames; ++f) s1_ptr[f] = e->d_s1.p + s1_off[f];
  if (int rc = encode_batch_device(e, n_frames, clouds, cloud_bytes, s1_ptr.data(), s1_cap.data(), 0)) return rc;
  // chunk tables + per-chunk slots
  const size_t slot = (lz4_bound(max_chunk) + 15) & ~size_t(15);
  if (slot > 0xFFFFFFF0ull) { set_error("chunk too large for LZ4"); return CLDN_ERR_UNSUPPORTED; }
  if (int rc = e->d_lz_scratch.reserve(std::max<size_t>(chunks, 1) * slot)) return rc;
  if (int rc = e->d_lz_chunk_sizes.reserve(chunks + 1)) return rc;
  if (int rc = e->d_lz_chunk_frame.reserve(chunks + 1)) return rc;
  if (int rc = e->d_lz_frames.reserve(n_frames)) return rc;
  Lz4Frame* hf = nullptr; int hf_slot = 0;
  uint32_t* hcf = nullptr; int hcf_slot = 0;
  if (int rc = e->h_lz_frames.acquire(n_frames, &hf, &hf_slot)) return rc;
  if (int rc = e->h_lz_chunk_frame.acquire(chunks + 1, &hcf, &hcf_slot)) return rc;
  uint32_t cb = 0;
  for (size_t f = 0; f < n_frames; ++f) {
    const size_t n = cloud_bytes[f] / info.point_step;
    Lz4Frame& F = hf[f];
    memset(&F, 0, sizeof(F));
    F.plain = e->d_s1.p + s1_off[f];
    F.plain_bytes = e->d_sizes.p + f;
    F.blob_out = static_cast<uint8_t*>(outs[f]);
    F.blob_cap = out_capacities[f];
    F.n_chunks = static_cast<uint32_t>((n + kChunkPoints - 1) / kChunkPoints);
    F.chunk_begin = cb;
    for (uint32_t c = 0; c < F.n_chunks; ++c) hcf[cb + c] = static_cast<uint32_t>(f);
    cb += F.n_chunks;
  }
  CUDA_TRY(cudaMemcpyAsync(e->d_lz_frames.p, hf, n_frames * sizeof(Lz4Frame), cudaMemcpyHostToDevice, e->stream));
  if (int rc = e->h_lz_frames.commit(hf_slot, e->stream)) return rc;
  if (chunks) CUDA_TRY(cudaMemcpyAsync(e->d_lz_chunk_frame.p, hcf, chunks * sizeof(uint32_t), cudaMemcpyHostToDevice, e->stream));
  if (int rc = e->h_lz_chunk_frame.commit(hcf_slot, e->stream)) return rc;
  Lz4Launch Z;
  Z.frames = e->d_lz_frames.p;
  Z.n_frames = static_cast<uint32_t>(n_frames);
  Z.n_chunks_total = static_cast<uint32_t>(chunks);
  Z.chunk_frame = e->d_lz_chunk_frame.p;
  Z.scratch = e->d_lz_scratch.p;
  Z.slot_stride = static_cast<uint32_t>(slot);
  Z.chunk_sizes = e->d_lz_chunk_sizes.p;
  Z.sizes = e->d_sizes.p;   // the pack kernel overwrites the stage-1 sizes with the blob sizes (it runs after every reader)
  Z.err = e->d_err.p;
  if (launch_lz4_compress(Z, e->d_header.p, static_cast<uint32_t>(hdr), e->stream) < 0) { set_error("LZ4 kernel launch failed"); return CLDN_ERR_CUDA; }
  CUDA_TRY(cudaGetLastError());
  return CLDN_OK;
}

static int encode_batch_impl(cldn_encoder_t* e, size_t n_frames, const void* const* clouds, const size_t* cloud_bytes,
                             void* const* outs, const size_t* out_capacities, int write_header, size_t* written_host,
                             int mem) {
  if (!e || (n_frames && (!clouds || !cloud_bytes || !outs || !out_capacities))) {
    set_error("null argument");
    return CLDN_ERR_INVALID_ARGUMENT;
  }
  if (n_frames == 0) return CLDN_OK;
  CUDA_TRY(cudaSetDevice(e->device));
  if (mem == CLDN_MEM_DEVICE && e->stage2 == CLDN_COMP_LZ4) {
    if (int rc = encode_batch_device_lz4(e, n_frames, clouds, cloud_bytes, outs, out_capacities, write_header)) return rc;
    if (written_host) {
      if (int rc = e->h_sizes.reserve(n_frames)) return rc;
      CUDA_TRY(cudaMemcpyAsync(e->h_sizes.p, e->d_sizes.p, n_frames * sizeof(uint64_t), cudaMemcpyDeviceToHost, e->stream));
      if (int rc = check_device_error(e->stream, e->d_err.p, e->h_err.p)) return rc;
      for (size_t f = 0; f < n_frames; ++f) written_host[f] = static_cast<size_t>(e->h_sizes.p[f]);
    }
    return CLDN_OK;
  }
  if (mem == CLDN_MEM_DEVICE && e->stage2 != CLDN_COMP_NONE) {
    set_error("compression_opt %d needs the host-pointer API: ZSTD is delegated to the host library (LZ4 runs on the device)", e->stage2);
    return CLDN_ERR_UNSUPPORTED;
  }
  if (mem == CLDN_MEM_HOST && e->stage2 != CLDN_COMP_NONE) {
    if (!g_stage2.load(e->stage2)) return CLDN_ERR_UNSUPPORTED;
    // stage 1 on the GPU, then every chunk through the system compressor (not pipelined: this is not the product path)
    cldn_info_t full = e->info;
    full.compression_opt = static_cast<uint8_t>(e->stage2);
    const size_t hdr = write_header ? e->header.size() : 0;
    for (size_t f = 0; f < n_frames; ++f) {
      if (e->info.point_step == 0) { set_error("point_step cannot be 0"); return CLDN_ERR_INVALID_ARGUMENT; }
      if (cloud_bytes[f] % e->info.point_step != 0) { set_error("Input cloud_data size is not a multiple of point_step"); return CLDN_ERR_INVALID_ARGUMENT; }
      const size_t n = cloud_bytes[f] / e->info.point_step;
      bool ok;
      const size_t need = max_compressed_size(full, n, false, &ok) + hdr;  // cloudini.cpp:530-534
      if (!ok) return CLDN_ERR_INVALID_ARGUMENT;
      if (out_capacities[f] < need) { set_error("Output buffer too small for worst-case compressed size"); return CLDN_ERR_BUFFER_TOO_SMALL; }
      size_t cap1 = max_compressed_size(e->info, n, false, &ok);
      if (int rc = e->d_in.reserve(cloud_bytes[f] + 256)) return rc;
      if (int rc = e->d_out.reserve(cap1 + 256)) return rc;
      if (cloud_bytes[f]) CUDA_TRY(cudaMemcpyAsync(e->d_in.p, clouds[f], cloud_bytes[f], cudaMemcpyHostToDevice, e->stream));
      const void* din = e->d_in.p;
      void* dout = e->d_out.p;
      if (int rc = encode_batch_device(e, 1, &din, &cloud_bytes[f], &dout, &cap1, 0)) return rc;
      if (int rc = e->h_sizes.reserve(1)) return rc;
      CUDA_TRY(cudaMemcpyAsync(e->h_sizes.p, e->d_sizes.p, sizeof(uint64_t), cudaMemcpyDeviceToHost, e->stream));
      if (int rc = check_device_error(e->stream, e->d_err.p, e->h_err.p)) return rc;
      const size_t s1 = static_cast<size_t>(e->h_sizes.p[0]);
      if (s1 > cap1) { set_error("Output buffer too small for uncompressed chunk"); return CLDN_ERR_BUFFER_TOO_SMALL; }
      e->s2_tmp.resize(s1 + 16);
      if (s1) CUDA_TRY(cudaMemcpy(e->s2_tmp.data(), e->d_out.p, s1, cudaMemcpyDeviceToHost));
      uint8_t* out = static_cast<uint8_t*>(outs[f]);
      if (hdr) memcpy(out, e->header.data(), hdr);
      size_t w = 0;
      if (int rc = stage2_compress_payload(e->stage2, e->s2_tmp.data(), s1, out + hdr, out_capacities[f] - hdr, &w)) return rc;
      if (written_host) written_host[f] = hdr + w;
    }
    return CLDN_OK;
  }
  if (mem == CLDN_MEM_DEVICE) {
    if (int rc = encode_batch_device(e, n_frames, clouds, cloud_bytes, outs, out_capacities, write_header)) return rc;
    if (written_host) {
      if (int rc = e->h_sizes.reserve(n_frames)) return rc;
      CUDA_TRY(cudaMemcpyAsync(e->h_sizes.p, e->d_sizes.p, n_frames * sizeof(uint64_t), cudaMemcpyDeviceToHost, e->stream));
      if (int rc = check_device_error(e->stream, e->d_err.p, e->h_err.p)) return rc;
      for (size_t f = 0; f < n_frames; ++f) written_host[f] = static_cast<size_t>(e->h_sizes.p[f]);
    }
    return CLDN_OK;
  }
  if (mem != CLDN_MEM_HOST) { set_error("bad mem kind"); return CLDN_ERR_INVALID_ARGUMENT; }
  // ---- host path: H2D -> kernels -> D2H (sizes first, then exactly the encoded bytes) ----
  size_t in_total = 0, out_total = 0;
  std::vector<size_t> in_off(n_frames), out_off(n_frames), caps(n_frames);
  const size_t hdr = write_header ? e->header.size() : 0;
  for (size_t f = 0; f < n_frames; ++f) {
    if (e->info.point_step == 0) { set_error("point_step cannot be 0"); return CLDN_ERR_INVALID_ARGUMENT; }
    if (cloud_bytes[f] % e->info.point_step != 0) { set_error("Input cloud_data size is not a multiple of point_step"); return CLDN_ERR_INVALID_ARGUMENT; }
    bool ok;
    const size_t need = max_compressed_size(e->info, cloud_bytes[f] / e->info.point_step, false, &ok) + hdr;
    if (!ok) return CLDN_ERR_INVALID_ARGUMENT;
    if (out_capacities[f] < need) { set_error("Output buffer too small for worst-case compressed size"); return CLDN_ERR_BUFFER_TOO_SMALL; }
    in_off[f] = in_total;
    in_total += (cloud_bytes[f] + 255) & ~size_t(255);
    out_off[f] = out_total;
    // A committed V5 mode can exceed the reference's own worst-case formula (DeltaRle on a 64-bit field: 11 bytes per
    // value against the 10 it budgets); the reference then succeeds iff the caller's buffer is large enough. Stage with
    // room for the real worst case, never more than the caller gave: the kernels check every store against this bound.
    size_t worst = need;
    for (uint32_t s2 = 0; s2 < e->plan.n_sections; ++s2) {
      if (e->plan.sections[s2].bpv == 8) worst += cloud_bytes[f] / e->info.point_step + 16 * ((cloud_bytes[f] / e->info.point_step) / kChunkPoints + 1);
    }
    caps[f] = std::min(out_capacities[f], worst);
    out_total += (caps[f] + 255) & ~size_t(255);
  }
  if (int rc = e->d_in.reserve(in_total + 256)) return rc;
  if (int rc = e->d_out.reserve(out_total + 256)) return rc;
  if (int rc = e->pipe.ensure(2 * n_frames)) return rc;
  if (int rc = e->h_sizes.reserve(n_frames)) return rc;
  // frame f: upload on the h2d stream -> encode on the handle's stream -> 8-byte size read-back
  for (size_t f = 0; f < n_frames; ++f) {
    const void* din = e->d_in.p + in_off[f];
    void* dout = e->d_out.p + out_off[f];
    if (cloud_bytes[f]) CUDA_TRY(cudaMemcpyAsync(e->d_in.p + in_off[f], clouds[f], cloud_bytes[f], cudaMemcpyHostToDevice, e->pipe.h2d));
    CUDA_TRY(cudaEventRecord(e->pipe.ev[2 * f], e->pipe.h2d));
    CUDA_TRY(cudaStreamWaitEvent(e->stream, e->pipe.ev[2 * f], 0));
    if (int rc = encode_batch_device(e, 1, &din, &cloud_bytes[f], &dout, &caps[f], write_header)) return rc;
    CUDA_TRY(cudaMemcpyAsync(e->h_sizes.p + f, e->d_sizes.p, sizeof(uint64_t), cudaMemcpyDeviceToHost, e->stream));
    CUDA_TRY(cudaEventRecord(e->pipe.ev[2 * f + 1], e->stream));
  }
  // as soon as a frame's size is known, its blob goes down on the d2h stream (exactly `size` bytes)
  for (size_t f = 0; f < n_frames; ++f) {
    CUDA_TRY(cudaEventSynchronize(e->pipe.ev[2 * f + 1]));
    const size_t sz = static_cast<size_t>(e->h_sizes.p[f]);
    if (sz > caps[f]) {  // the kernels refused the stores that did not fit and raised the error word
      if (int rc = check_device_error(e->stream, e->d_err.p, e->h_err.p)) return rc;
      set_error("Output buffer too small for uncompressed chunk");
      return CLDN_ERR_BUFFER_TOO_SMALL;
    }
    if (sz) CUDA_TRY(cudaMemcpyAsync(outs[f], e->d_out.p + out_off[f], sz, cudaMemcpyDeviceToHost, e->pipe.d2h));
    if (written_host) written_host[f] = sz;
  }
  CUDA_TRY(cudaStreamSynchronize(e->pipe.d2h));
  return check_device_error(e->stream, e->d_err.p, e->h_err.p);
}

extern "C" {

int cldn_b200_encode_batch(cldn_encoder_t* e, size_t n_frames, const void* const* clouds, const size_t* cloud_bytes,
                           void* const* outs, const size_t* out_capacities, int write_header, size_t* written_host,
                           int mem) {
  return guarded([&] { return encode_batch_impl(e, n_frames, clouds, cloud_bytes, outs, out_capacities, write_header, written_host, mem); });
}

int cldn_b200_encode(cldn_encoder_t* enc, const void* cloud, size_t cloud_bytes, void* out, size_t out_capacity,
                     int write_header, size_t* written, int mem) {
  const void* clouds[1] = {cloud};
  void* outs[1] = {out};
  return cldn_b200_encode_batch(enc, 1, clouds, &cloud_bytes, outs, &out_capacity, write_header, written, mem);
}

}  // extern "C"

// =====================================================================================================================
struct cldn_decoder {
  std::vector<uint8_t> s2_tmp;  // stage-2 decompression scratch (host-pointer API only)
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  cldn_info_t info;  // info of the cached plan
  bool have_plan = false;
  Plan plan;
  bool has_padding = false;
  DevBuf<Plan> d_plan;
  DevBuf<DecFrame> d_frames;
  PinRing<DecFrame> h_frames;
  DevBuf<uint64_t> d_chunk_offsets;
  DevBuf<uint32_t> d_chunk_sizes;
  DevBuf<uint32_t> d_err;
  PinBuf<uint32_t> h_err;
  DevBuf<uint8_t> d_in, d_out;
  DevBuf<uint32_t> d_chunk_tiles, d_chunk_tile_begin, d_stream_end, d_tsums, d_chunk_frame, d_tile_chunk;
  DevBuf<uint64_t> d_tstatus;
  DevBuf<uint64_t> d_trace;
  DevBuf<uint32_t> d_counter, d_redo;
  DevBuf<uint8_t> d_side;         // side mode: section values decoded ahead of the regular stream (DecLaunch::side)
  DevBuf<uint64_t> d_chunk_desc;  // chunk-sequential kernel: self-validating chunk descriptors (see walk_frame_publish)
  uint32_t desc_tag = 0;
  // stage 2 on the device (LZ4)
  DevBuf<uint8_t> d_s1, d_lz_scratch;
  DevBuf<uint32_t> d_lz_chunk_frame, d_lz_chunk_sizes;
  DevBuf<Lz4Frame> d_lz_frames;
  DevBuf<uint64_t> d_lz_sizes;
  PinBuf<uint64_t> h_lz_sizes;
  PinRing<Lz4Frame> h_lz_frames;
  PinRing<uint32_t> h_lz_chunk_frame;
  bool fast_launched = false;  // the last batch went through decode_floatn_fast_kernel (cldn_b200_decoder_last_stats)
  uint32_t fast_chunks = 0;
  bool sections_ahead = false;    // the last batch ran in side mode (cldn_b200_decoder_last_sections_ahead)
  CopyPipeline pipe;
  uint32_t epoch = 0;
};

static bool same_info(const cldn_info_t& a, const cldn_info_t& b) {
  if (a.width != b.width || a.height != b.height || a.point_step != b.point_step || a.encoding_opt != b.encoding_opt ||
      a.compression_opt != b.compression_opt || a.version != b.version || a.n_fields != b.n_fields) return false;
  for (uint32_t i = 0; i < a.n_fields; ++i) {
    const cldn_field_t &x = a.fields[i], &y = b.fields[i];
    if (x.offset != y.offset || x.type != y.type || x.has_resolution != y.has_resolution ||
        (x.has_resolution && memcmp(&x.resolution, &y.resolution, 4) != 0)) return false;
  }
  return true;
}

// PointcloudDecoder::updateDecoders (cloudini.cpp:627-633)
static int decoder_update_plan(cldn_decoder* d, const cldn_info_t& info) {
  if (d->have_plan && same_info(d->info, info)) return CLDN_OK;
  Plan plan;
  if (int rc = build_decode_plan(info, &plan)) return rc;
  if (!plan.supported) {
    set_error("EncodingInfo contains a field decoder this build cannot run");
    return CLDN_ERR_UNSUPPORTED;
  }
  if (info.compression_opt != CLDN_COMP_NONE) {
    set_error("compression_opt %d: stage 2 is undone before the stage-1 decode (LZ4 on the device, ZSTD by the host library)", static_cast<int>(info.compression_opt));
    return CLDN_ERR_UNSUPPORTED;
  }
  if (info.version < 3) { set_error("wire version %d (single unframed chunk) is not supported", info.version); return CLDN_ERR_UNSUPPORTED; }
  if (int rc = d->d_plan.reserve(1)) return rc;
  CUDA_TRY(cudaMemcpyAsync(d->d_plan.p, &plan, sizeof(Plan), cudaMemcpyHostToDevice, d->stream));
  CUDA_TRY(cudaStreamSynchronize(d->stream));  // `plan` is a stack object
  d->plan = plan;
  d->info = info;
  d->have_plan = true;
  // do the declared fields cover every byte of a point? (otherwise host outputs must round-trip their padding)
  // union of the field intervals, clipped to the point: at most CLDN_MAX_FIELDS of them, no point_step-sized table
  // (the header is untrusted: a forged point_step must not size an allocation)
  std::pair<uint64_t, uint64_t> iv[CLDN_MAX_FIELDS];
  uint32_t n_iv = 0;
  for (uint32_t i = 0; i < info.n_fields && n_iv < CLDN_MAX_FIELDS; ++i) {
    const uint64_t lo = info.fields[i].offset, hi = std::min<uint64_t>(lo + static_cast<uint64_t>(size_of_type(info.fields[i].type)), info.point_step);
    if (lo < hi) iv[n_iv++] = {lo, hi};
  }
  std::sort(iv, iv + n_iv);
  uint64_t covered_to = 0;
  bool gap = false;
  for (uint32_t i = 0; i < n_iv; ++i) {
    if (iv[i].first > covered_to) gap = true;
    covered_to = std::max(covered_to, iv[i].second);
  }
  d->has_padding = gap || covered_to < info.point_step;
  return CLDN_OK;
}

static int decode_batch_device(cldn_decoder* d, const cldn_info_t& info, size_t n_frames, const void* const* payloads,
                               const size_t* payload_bytes, void* const* outs, const size_t* out_capacities) {
  d->sections_ahead = false;
  if (int rc = decoder_update_plan(d, info)) return rc;
  const uint64_t n_points = static_cast<uint64_t>(info.width) * info.height;
  if (n_points > 0xFFFFFFFFull) { set_error("too many points"); return CLDN_ERR_UNSUPPORTED; }
  const uint64_t out_need = n_points * info.point_step;
  DecFrame* hf = nullptr;
  int hf_slot = 0;
  if (int rc = d->h_frames.acquire(n_frames, &hf, &hf_slot)) return rc;
  if (int rc = d->d_frames.reserve(n_frames)) return rc;
  uint64_t chunks = 0;
  for (size_t f = 0; f < n_frames; ++f) {
    if (out_capacities[f] < out_need) {
      set_error("Output buffer is too small to hold the decoded data");  // v4_codec.cpp:93-95 / v5_codec.cpp:991-993
      return CLDN_ERR_BUFFER_TOO_SMALL;
    }
    DecFrame& F = hf[f];
    F.payload = static_cast<const uint8_t*>(payloads[f]);
    F.payload_bytes = payload_bytes[f];
    F.out = static_cast<uint8_t*>(outs[f]);
    F.n_points = static_cast<uint32_t>(n_points);
    F.n_chunks = static_cast<uint32_t>((n_points + kChunkPoints - 1) / kChunkPoints);
    F.chunk_begin = static_cast<uint32_t>(chunks);
    F.pad_ = 0;
    chunks += F.n_chunks;
  }
  if (chunks > 0x7FFFFFFFull) { set_error("batch too large"); return CLDN_ERR_UNSUPPORTED; }
  if (int rc = d->d_chunk_offsets.reserve(static_cast<size_t>(chunks) + 1)) return rc;
  if (int rc = d->d_chunk_sizes.reserve(static_cast<size_t>(chunks) + 1)) return rc;
  CUDA_TRY(cudaMemcpyAsync(d->d_frames.p, hf, n_frames * sizeof(DecFrame), cudaMemcpyHostToDevice, d->stream));
  if (int rc = d->h_frames.commit(hf_slot, d->stream)) return rc;
  DecLaunch L;
  L.mix_chase = 0;
  L.par_runs = unmeasured_kernels_enabled() ? 1u : 0u;
  L.frames = d->d_frames.p;
  L.n_frames = static_cast<uint32_t>(n_frames);
  L.n_chunks_total = static_cast<uint32_t>(chunks);
  L.plan = d->d_plan.p;
  L.chunk_offsets = d->d_chunk_offsets.p;
  L.chunk_sizes = d->d_chunk_sizes.p;
  L.err = d->d_err.p;
  L.chunk_frame = nullptr; L.tile_chunk = nullptr;
  L.chunk_tiles = nullptr; L.chunk_tile_begin = nullptr; L.stream_end = nullptr; L.tstatus = nullptr; L.tsums = nullptr;
  L.tile_capacity = 0; L.tile_grid = 0; L.epoch = 0; L.trace = nullptr; L.chunk_counter = nullptr; L.sections_only = 0;
  L.chunk_desc = nullptr; L.desc_tag = 0; L.uniform_chunks = 0; L.redo_list = nullptr; L.redo_mode = 0;
  L.side = nullptr; L.side_mode = 0;
  for (int s = 0; s < kMaxSideFields; ++s) L.side_off[s] = 0;
  if (d->plan.n_sections > 0 && chunks > 0 && (d->plan.regular_overlap || !(d->plan.all_varint || d->plan.n_ops == 0))) {
    // V5 with raw / XOR / Gorilla fields in the regular stream: the per-chunk parser records where the sections start
    if (int rc = d->d_stream_end.reserve(static_cast<size_t>(chunks) + 1)) return rc;
    L.stream_end = d->d_stream_end.p;
  }
  const char* force_chunk = getenv("CLDN_B200_FORCE_CHUNK_DECODE");
  const bool fast_general = !d->plan.floatn_only && chunks > 0 && decode_fast_general_plan(d->plan) &&
                            decode_tiles_sequential(static_cast<uint32_t>(chunks)) && decode_fast_enabled() && !(force_chunk && force_chunk[0] == '1');
  if (fast_general) {
    // float-varint streams with scalar lossy fields (Velodyne XYZIRT ...): fast chunk-sequential reader + redo list
    if (int rc = d->d_stream_end.reserve(static_cast<size_t>(chunks) + 1)) return rc;
    if (int rc = d->d_counter.reserve(4, true)) return rc;
    if (int rc = d->d_redo.reserve(static_cast<size_t>(chunks) + 1)) return rc;
    const size_t had = d->d_chunk_desc.cap;
    if (int rc = d->d_chunk_desc.reserve(2 * static_cast<size_t>(chunks) + 2)) return rc;
    d->desc_tag = (d->desc_tag + 1) & 0xFFFFFFu;
    if (d->d_chunk_desc.cap != had || d->desc_tag == 0) {
      CUDA_TRY(cudaMemsetAsync(d->d_chunk_desc.p, 0, d->d_chunk_desc.cap * sizeof(uint64_t), d->stream));
      if (d->desc_tag == 0) d->desc_tag = 1;
    }
    L.stream_end = d->d_stream_end.p;
    L.chunk_counter = d->d_counter.p;
    L.redo_list = d->d_redo.p;
    L.chunk_desc = d->d_chunk_desc.p;
    L.desc_tag = d->desc_tag;
    bool uniform = n_frames > 0 && hf[0].n_chunks > 0;
    for (size_t f = 1; f < n_frames; ++f) uniform = uniform && hf[f].n_chunks == hf[0].n_chunks;
    L.uniform_chunks = uniform ? hf[0].n_chunks : 0u;
  }
  if (d->plan.floatn_only && chunks > 0 && !(force_chunk && force_chunk[0] == '1')) {
    // tile-parallel path: host upper bound on the tile count (every chunk adds at most 2 tiles of rounding/misalignment)
    uint64_t tiles = 0;
    for (size_t f = 0; f < n_frames; ++f) tiles += payload_bytes[f] / decode_tile_bytes() + 2ull * hf[f].n_chunks + 1;
    if (tiles > 0x7FFFFFFFull) { set_error("batch too large"); return CLDN_ERR_UNSUPPORTED; }
    if (int rc = d->d_chunk_tiles.reserve(static_cast<size_t>(chunks) + 1)) return rc;
    if (int rc = d->d_chunk_tile_begin.reserve(static_cast<size_t>(chunks) + 2)) return rc;
    if (int rc = d->d_stream_end.reserve(static_cast<size_t>(chunks) + 1)) return rc;
    if (int rc = d->d_chunk_frame.reserve(static_cast<size_t>(chunks) + 1)) return rc;
    if (int rc = d->d_tile_chunk.reserve(static_cast<size_t>(tiles) + 1)) return rc;
    const size_t old_cap = d->d_tstatus.cap;
    if (int rc = d->d_tstatus.reserve(static_cast<size_t>(tiles) + 1, true)) return rc;
    if (d->d_tstatus.cap != old_cap || d->d_tsums.cap < d->d_tstatus.cap * 16) {
      if (int rc = d->d_tsums.reserve(d->d_tstatus.cap * 16, true)) return rc;
    }
    d->epoch = (d->epoch + 1) & 0x3FFFFFu;
    if (d->epoch == 0) {
      CUDA_TRY(cudaMemsetAsync(d->d_tstatus.p, 0, d->d_tstatus.cap * sizeof(uint64_t), d->stream));
      CUDA_TRY(cudaMemsetAsync(d->d_tsums.p, 0, d->d_tsums.cap * sizeof(uint32_t), d->stream));
      d->epoch = 1;
    }
    L.chunk_tiles = d->d_chunk_tiles.p;
    L.chunk_tile_begin = d->d_chunk_tile_begin.p;
    L.stream_end = d->d_stream_end.p;
    L.chunk_frame = d->d_chunk_frame.p;
    L.tile_chunk = d->d_tile_chunk.p;
    L.tstatus = d->d_tstatus.p;
    L.tsums = d->d_tsums.p;
    L.tile_capacity = static_cast<uint32_t>(d->d_tstatus.cap);
    L.tile_grid = static_cast<uint32_t>(tiles);
    L.epoch = d->epoch;
    if (int rc = d->d_counter.reserve(4, true)) return rc;
    L.chunk_counter = d->d_counter.p;
    if (int rc = d->d_redo.reserve(static_cast<size_t>(chunks) + 1)) return rc;
    L.redo_list = d->d_redo.p;
    {
      const size_t had = d->d_chunk_desc.cap;
      if (int rc = d->d_chunk_desc.reserve(2 * static_cast<size_t>(chunks) + 2)) return rc;
      d->desc_tag = (d->desc_tag + 1) & 0xFFFFFFu;
      if (d->d_chunk_desc.cap != had || d->desc_tag == 0) {  // fresh memory or tag wrap: no stale word may carry a live tag
        CUDA_TRY(cudaMemsetAsync(d->d_chunk_desc.p, 0, d->d_chunk_desc.cap * sizeof(uint64_t), d->stream));
        if (d->desc_tag == 0) d->desc_tag = 1;
      }
      L.chunk_desc = d->d_chunk_desc.p;
      L.desc_tag = d->desc_tag;
      bool uniform = n_frames > 0 && hf[0].n_chunks > 0;
      for (size_t f = 1; f < n_frames; ++f) uniform = uniform && hf[f].n_chunks == hf[0].n_chunks;
      L.uniform_chunks = uniform ? hf[0].n_chunks : 0u;
    }
    if (getenv("CLDN_B200_TRACE")) {
      if (int rc = d->d_trace.reserve(std::max<size_t>(static_cast<size_t>(tiles), static_cast<size_t>(chunks) * 80) * 8 + 8, true)) return rc;
      L.trace = d->d_trace.p;
    }
  }
  d->fast_launched = L.redo_list != nullptr && decode_tiles_sequential(L.n_chunks_total) && decode_fast_enabled() && !d->plan.regular_overlap;
  if (d->fast_launched && L.chunk_desc && L.stream_end && L.chunk_counter && decode_side_plan(d->plan) && decode_fast_general_plan(d->plan) &&
      !(force_chunk && force_chunk[0] == '1')) {
    // V5 sections ahead of the regular stream: compact per-chunk arrays the fast reader merges into the rows (DecLaunch::side)
    const size_t side_bytes = decode_side_bytes(d->plan, chunks, L.side_off);
    if (int rc = d->d_side.reserve(side_bytes + 256)) return rc;
    L.side = d->d_side.p;
  }
  d->sections_ahead = decode_side_active(d->plan, L);
  (void)fast_general;
  d->fast_chunks = d->fast_launched ? L.n_chunks_total : 0u;
  if (launch_decode(d->plan, L, d->stream) < 0) { set_error("decode kernel launch failed"); return CLDN_ERR_CUDA; }
  CUDA_TRY(cudaGetLastError());
  return CLDN_OK;
}

extern "C" {

int cldn_b200_decoder_create(int device, void* stream, cldn_decoder_t** out) {
  if (!out) { set_error("null argument"); return CLDN_ERR_INVALID_ARGUMENT; }
  *out = nullptr;
  if (int rc = select_device(device)) return rc;
  cldn_decoder* d = new cldn_decoder();
  cudaGetDevice(&d->device);
  if (stream) {
    d->stream = static_cast<cudaStream_t>(stream);
  } else {
    if (cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking) != cudaSuccess) {
      set_error("cudaStreamCreate failed");
      delete d;
      return CLDN_ERR_CUDA;
    }
    d->own_stream = true;
  }
  int rc = d->d_err.reserve(1, true);
  if (!rc) rc = d->h_err.reserve(1);
  if (rc) { cldn_b200_decoder_destroy(d); return rc; }
  *out = d;
  return CLDN_OK;
}

void cldn_b200_decoder_destroy(cldn_decoder_t* d) {
  if (!d) return;
  cudaSetDevice(d->device);
  if (d->stream) cudaStreamSynchronize(d->stream);
  d->d_plan.release(); d->d_frames.release(); d->h_frames.release(); d->d_chunk_offsets.release(); d->d_chunk_sizes.release();
  d->d_err.release(); d->h_err.release(); d->d_in.release(); d->d_out.release();
  d->d_chunk_tiles.release(); d->d_chunk_tile_begin.release(); d->d_stream_end.release(); d->d_tsums.release(); d->d_tstatus.release(); d->d_chunk_frame.release(); d->d_tile_chunk.release(); d->d_trace.release(); d->d_counter.release(); d->d_redo.release(); d->d_side.release(); d->d_chunk_desc.release(); d->pipe.release();
  d->d_s1.release(); d->d_lz_scratch.release(); d->d_lz_chunk_frame.release(); d->d_lz_chunk_sizes.release(); d->d_lz_frames.release();
  d->d_lz_sizes.release(); d->h_lz_sizes.release(); d->h_lz_frames.release(); d->h_lz_chunk_frame.release();
  if (d->own_stream && d->stream) cudaStreamDestroy(d->stream);
  delete d;
}

int cldn_b200_decoder_sync(cldn_decoder_t* d) {
  if (!d) { set_error("null decoder"); return CLDN_ERR_INVALID_ARGUMENT; }
  CUDA_TRY(cudaSetDevice(d->device));
  if (const char* tp = getenv("CLDN_B200_TRACE")) {  // development aid: dump the per-tile phase timestamps
    if (d->d_trace.p) {
      cudaStreamSynchronize(d->stream);
      std::vector<uint64_t> h(d->d_trace.cap);
      cudaMemcpy(h.data(), d->d_trace.p, h.size() * 8, cudaMemcpyDeviceToHost);
      if (FILE* f = fopen(tp, "wb")) { fwrite(h.data(), 8, h.size(), f); fclose(f); }
    }
  }
  return check_device_error(d->stream, d->d_err.p, d->h_err.p);
}

int cldn_b200_decoder_last_stats(cldn_decoder_t* d, uint32_t stats[2]) {
  if (!d || !stats) { set_error("null argument"); return CLDN_ERR_INVALID_ARGUMENT; }
  stats[0] = stats[1] = 0;
  CUDA_TRY(cudaSetDevice(d->device));
  if (!d->d_counter.p || !d->fast_launched) return CLDN_OK;
  uint32_t c[4] = {0, 0, 0, 0};
  CUDA_TRY(cudaStreamSynchronize(d->stream));
  CUDA_TRY(cudaMemcpy(c, d->d_counter.p, sizeof(c), cudaMemcpyDeviceToHost));
  stats[0] = d->fast_chunks;
  stats[1] = c[3];
  return CLDN_OK;
}

int cldn_b200_decoder_last_sections_ahead(cldn_decoder_t* d) {
  if (!d) { set_error("null argument"); return CLDN_ERR_INVALID_ARGUMENT; }
  return d->sections_ahead ? 1 : 0;
}

}  // extern "C"

static int decode_batch_device(cldn_decoder* d, const cldn_info_t& info, size_t n_frames, const void* const* payloads,
                               const size_t* payload_bytes, void* const* outs, const size_t* out_capacities);

// Device pointers, compression LZ4: every chunk's block is decompressed on the device into a slot, the slots are re-framed
// into a stage-1 payload per frame (internal arena) and the ordinary stage-1 decode runs on that. One read-back of the
// frames' stage-1 sizes (8 bytes each) sits between the two.
static int decode_batch_device_lz4(cldn_decoder* d, const cldn_info_t& info, size_t n_frames, const void* const* payloads,
                                   const size_t* payload_bytes, void* const* outs, const size_t* out_capacities) {
  cldn_info_t plain = info;
  plain.compression_opt = CLDN_COMP_NONE;
  const uint64_t n_points = static_cast<uint64_t>(info.width) * info.height;
  if (n_points > 0xFFFFFFFFull) { set_error("too many points"); return CLDN_ERR_UNSUPPORTED; }
  for (size_t f = 0; f < n_frames; ++f) {
    if (out_capacities[f] < n_points * info.point_step) { set_error("Output buffer is too small to hold the decoded data"); return CLDN_ERR_BUFFER_TOO_SMALL; }
  }
  bool ok;
  const size_t chunks_per_frame = static_cast<size_t>((n_points + kChunkPoints - 1) / kChunkPoints);
  const size_t max_chunk = max_compressed_size(plain, std::min<size_t>(n_points, kChunkPoints), false, &ok) + 64;
  if (!ok) return CLDN_ERR_INVALID_ARGUMENT;
  const size_t slot = (max_chunk + 15) & ~size_t(15);
  const size_t s1_cap = ((chunks_per_frame * (slot + 4) + 255) & ~size_t(255));
  const size_t chunks = chunks_per_frame * n_frames;
  if (chunks > 0x7FFFFFFFull || slot > 0xFFFFFFF0ull) { set_error("batch too large"); return CLDN_ERR_UNSUPPORTED; }
  if (int rc = d->d_s1.reserve(std::max<size_t>(s1_cap * n_frames, 256))) return rc;
  if (int rc = d->d_lz_scratch.reserve(std::max<size_t>(chunks, 1) * slot)) return rc;
  if (int rc = d->d_lz_chunk_sizes.reserve(chunks + 1)) return rc;
  if (int rc = d->d_lz_chunk_frame.reserve(chunks + 1)) return rc;
  if (int rc = d->d_lz_frames.reserve(n_frames)) return rc;
  if (int rc = d->d_lz_sizes.reserve(n_frames)) return rc;
  if (int rc = d->h_lz_sizes.reserve(n_frames)) return rc;
  Lz4Frame* hf = nullptr; int hf_slot = 0;
  uint32_t* hcf = nullptr; int hcf_slot = 0;
  if (int rc = d->h_lz_frames.acquire(n_frames, &hf, &hf_slot)) return rc;
  if (int rc = d->h_lz_chunk_frame.acquire(chunks + 1, &hcf, &hcf_slot)) return rc;
  for (size_t f = 0; f < n_frames; ++f) {
    Lz4Frame& F = hf[f];
    memset(&F, 0, sizeof(F));
    F.packed_in = static_cast<const uint8_t*>(payloads[f]);
    F.packed_in_bytes = payload_bytes[f];
    F.plain_out = d->d_s1.p + f * s1_cap;
    F.plain_cap = s1_cap;
    F.n_chunks = static_cast<uint32_t>(chunks_per_frame);
    F.chunk_begin = static_cast<uint32_t>(f * chunks_per_frame);
    for (size_t c = 0; c < chunks_per_frame; ++c) hcf[f * chunks_per_frame + c] = static_cast<uint32_t>(f);
  }
  CUDA_TRY(cudaMemcpyAsync(d->d_lz_frames.p, hf, n_frames * sizeof(Lz4Frame), cudaMemcpyHostToDevice, d->stream));
  if (int rc = d->h_lz_frames.commit(hf_slot, d->stream)) return rc;
  if (chunks) CUDA_TRY(cudaMemcpyAsync(d->d_lz_chunk_frame.p, hcf, chunks * sizeof(uint32_t), cudaMemcpyHostToDevice, d->stream));
  if (int rc = d->h_lz_chunk_frame.commit(hcf_slot, d->stream)) return rc;
  Lz4Launch Z;
  Z.frames = d->d_lz_frames.p;
  Z.n_frames = static_cast<uint32_t>(n_frames);
  Z.n_chunks_total = static_cast<uint32_t>(chunks);
  Z.chunk_frame = d->d_lz_chunk_frame.p;
  Z.scratch = d->d_lz_scratch.p;
  Z.slot_stride = static_cast<uint32_t>(slot);
  Z.chunk_sizes = d->d_lz_chunk_sizes.p;
  Z.sizes = d->d_lz_sizes.p;
  Z.err = d->d_err.p;
  if (launch_lz4_decompress(Z, d->stream) < 0) { set_error("LZ4 kernel launch failed"); return CLDN_ERR_CUDA; }
  CUDA_TRY(cudaMemcpyAsync(d->h_lz_sizes.p, d->d_lz_sizes.p, n_frames * sizeof(uint64_t), cudaMemcpyDeviceToHost, d->stream));
  if (int rc = check_device_error(d->stream, d->d_err.p, d->h_err.p)) return rc;  // synchronises; damaged blocks stop here
  std::vector<const void*> p1(n_frames);
  std::vector<size_t> b1(n_frames);
  for (size_t f = 0; f < n_frames; ++f) { p1[f] = d->d_s1.p + f * s1_cap; b1[f] = static_cast<size_t>(d->h_lz_sizes.p[f]); }
  return decode_batch_device(d, plain, n_frames, p1.data(), b1.data(), outs, out_capacities);
}

static int decode_batch_impl(cldn_decoder_t* d, const cldn_info_t* info, size_t n_frames, const void* const* payloads,
                             const size_t* payload_bytes, void* const* outs, const size_t* out_capacities, int mem, int sync) {
  if (!d || !info || (n_frames && (!payloads || !payload_bytes || !outs || !out_capacities))) {
    set_error("null argument");
    return CLDN_ERR_INVALID_ARGUMENT;
  }
  if (n_frames == 0) return CLDN_OK;
  CUDA_TRY(cudaSetDevice(d->device));
  if (mem == CLDN_MEM_DEVICE) {
    if (info->compression_opt == CLDN_COMP_LZ4) {
      if (int rc = decode_batch_device_lz4(d, *info, n_frames, payloads, payload_bytes, outs, out_capacities)) return rc;
    } else if (int rc = decode_batch_device(d, *info, n_frames, payloads, payload_bytes, outs, out_capacities)) {
      return rc;
    }
    if (sync) return check_device_error(d->stream, d->d_err.p, d->h_err.p);
    return CLDN_OK;
  }
  if (mem != CLDN_MEM_HOST) { set_error("bad mem kind"); return CLDN_ERR_INVALID_ARGUMENT; }
  // the reference rejects a payload that still carries the header (cloudini.cpp:640-643)
  for (size_t f = 0; f < n_frames; ++f) {
    if (payload_bytes[f] >= 10 && memcmp(payloads[f], "CLOUDINI_V", 10) == 0) {
      set_error("compressed_data contains the header. You should use DecodeHeader first");
      return CLDN_ERR_INVALID_ARGUMENT;
    }
  }
  if (info->compression_opt != CLDN_COMP_NONE) {
    // stage 2 first (host libraries), then the ordinary stage-1 decode of the re-framed payload
    if (info->compression_opt > CLDN_COMP_ZSTD) { set_error("Unsupported compression option"); return CLDN_ERR_INVALID_ARGUMENT; }
    if (!g_stage2.load(info->compression_opt)) return CLDN_ERR_UNSUPPORTED;
    cldn_info_t plain = *info;
    plain.compression_opt = CLDN_COMP_NONE;
    bool ok;
    const size_t pts = static_cast<size_t>(info->width) * info->height;
    for (size_t f = 0; f < n_frames; ++f) {  // before any scratch is sized from the (untrusted) header
      if (out_capacities[f] / std::max<uint32_t>(info->point_step, 1u) < pts) {
        set_error("Output buffer is too small to hold the decoded data");
        return CLDN_ERR_BUFFER_TOO_SMALL;
      }
    }
    size_t max_body = max_compressed_size(plain, std::min<size_t>(pts, kChunkPoints), false, &ok);
    if (!ok) return CLDN_ERR_INVALID_ARGUMENT;
    max_body = std::max<size_t>(max_body, pts * info->point_step) + 64;  // DecompressChunk's bound is w*h*step (cloudini.cpp:672-675)
    for (size_t f = 0; f < n_frames; ++f) {
      if (int rc = stage2_decompress_payload(info->compression_opt, static_cast<const uint8_t*>(payloads[f]), payload_bytes[f], d->s2_tmp, max_body)) return rc;
      const void* p1 = d->s2_tmp.data();
      const size_t b1 = d->s2_tmp.size();
      if (int rc = decode_batch_impl(d, &plain, 1, &p1, &b1, &outs[f], &out_capacities[f], CLDN_MEM_HOST, 1)) return rc;
    }
    return CLDN_OK;
  }
  if (int rc = decoder_update_plan(d, *info)) return rc;
  const size_t out_need = static_cast<size_t>(info->width) * info->height * info->point_step;
  size_t in_total = 0;
  std::vector<size_t> in_off(n_frames);
  for (size_t f = 0; f < n_frames; ++f) {
    if (out_capacities[f] < out_need) { set_error("Output buffer is too small to hold the decoded data"); return CLDN_ERR_BUFFER_TOO_SMALL; }
    in_off[f] = in_total;
    in_total += (payload_bytes[f] + 255) & ~size_t(255);
  }
  const size_t out_stride = (out_need + 255) & ~size_t(255);
  if (int rc = d->d_in.reserve(in_total + 256)) return rc;
  if (int rc = d->d_out.reserve(out_stride * n_frames + 256)) return rc;
  if (int rc = d->pipe.ensure(2 * n_frames)) return rc;
  // frame f: upload on the h2d stream -> decode on the handle's stream -> download on the d2h stream
  for (size_t f = 0; f < n_frames; ++f) {
    const void* din = d->d_in.p + in_off[f];
    void* dout = d->d_out.p + out_stride * f;
    if (payload_bytes[f]) CUDA_TRY(cudaMemcpyAsync(d->d_in.p + in_off[f], payloads[f], payload_bytes[f], cudaMemcpyHostToDevice, d->pipe.h2d));
    // only declared field bytes are written by the decoder: keep the caller's padding bytes intact
    if (d->has_padding && out_need) CUDA_TRY(cudaMemcpyAsync(dout, outs[f], out_need, cudaMemcpyHostToDevice, d->pipe.h2d));
    CUDA_TRY(cudaEventRecord(d->pipe.ev[2 * f], d->pipe.h2d));
    CUDA_TRY(cudaStreamWaitEvent(d->stream, d->pipe.ev[2 * f], 0));
    if (int rc = decode_batch_device(d, *info, 1, &din, &payload_bytes[f], &dout, &out_need)) return rc;
    CUDA_TRY(cudaEventRecord(d->pipe.ev[2 * f + 1], d->stream));
    CUDA_TRY(cudaStreamWaitEvent(d->pipe.d2h, d->pipe.ev[2 * f + 1], 0));
    if (out_need) CUDA_TRY(cudaMemcpyAsync(outs[f], dout, out_need, cudaMemcpyDeviceToHost, d->pipe.d2h));
  }
  CUDA_TRY(cudaStreamSynchronize(d->pipe.d2h));
  // a failed decode must not hand back garbage silently: the error word is checked after everything has drained
  return check_device_error(d->stream, d->d_err.p, d->h_err.p);
}

extern "C" {

int cldn_b200_decode_batch(cldn_decoder_t* d, const cldn_info_t* info, size_t n_frames, const void* const* payloads,
                           const size_t* payload_bytes, void* const* outs, const size_t* out_capacities, int mem, int sync) {
  return guarded([&] { return decode_batch_impl(d, info, n_frames, payloads, payload_bytes, outs, out_capacities, mem, sync); });
}

int cldn_b200_decode(cldn_decoder_t* dec, const cldn_info_t* info, const void* payload, size_t payload_bytes, void* out,
                     size_t out_capacity, int mem) {
  const void* payloads[1] = {payload};
  void* outs[1] = {out};
  return cldn_b200_decode_batch(dec, info, 1, payloads, &payload_bytes, outs, &out_capacity, mem, 1);
}

// ---- one-shot helpers shaped like the reference's WASM C ABI ------------------------------------------------------
uint32_t cldn_b200_EncodePointcloudData(const char* header_as_yaml, const void* pc_data, uint32_t pc_data_size,
                                        void* output_data, uint32_t output_capacity) {
  if (!header_as_yaml || !output_data) { set_error("null argument"); return 0; }
  cldn_info_t info;
  if (info_from_yaml(header_as_yaml, strlen(header_as_yaml), &info) != CLDN_OK) return 0;
  const uint64_t expected = static_cast<uint64_t>(info.width) * info.height * info.point_step;
  if (pc_data_size != expected) {  // wasm_functions.cpp:222-228
    set_error("Data size mismatch: expected %llu but got %u", static_cast<unsigned long long>(expected), pc_data_size);
    return 0;
  }
  cldn_encoder_t* enc = nullptr;
  if (cldn_b200_encoder_create(&info, -1, nullptr, &enc) != CLDN_OK) return 0;
  bool ok;
  const size_t cap = max_compressed_size(info, pc_data_size / info.point_step, true, &ok);
  std::vector<uint8_t> tmp(ok ? cap : 0);
  size_t written = 0;
  int rc = ok ? cldn_b200_encode(enc, pc_data, pc_data_size, tmp.data(), tmp.size(), 1, &written, CLDN_MEM_HOST) : CLDN_ERR_INVALID_ARGUMENT;
  cldn_b200_encoder_destroy(enc);
  if (rc != CLDN_OK) return 0;
  if (written > output_capacity) { set_error("Output buffer too small for encoded data. Need %zu bytes, got %u", written, output_capacity); return 0; }
  memcpy(output_data, tmp.data(), written);
  return static_cast<uint32_t>(written);
}

uint32_t cldn_b200_DecodeCompressedData(const void* encoded_data, uint32_t encoded_data_size, void* output_data,
                                        uint32_t output_capacity) {
  if (!encoded_data || !output_data) { set_error("null argument"); return 0; }
  cldn_info_t info;
  size_t hdr = 0;
  if (cldn_b200_decode_header(static_cast<const uint8_t*>(encoded_data), encoded_data_size, &info, &hdr) != CLDN_OK) return 0;
  const uint64_t decoded = static_cast<uint64_t>(info.width) * info.height * info.point_step;
  if (decoded > output_capacity) { set_error("output buffer too small"); return 0; }
  cldn_decoder_t* dec = nullptr;
  if (cldn_b200_decoder_create(-1, nullptr, &dec) != CLDN_OK) return 0;
  const int rc = cldn_b200_decode(dec, &info, static_cast<const uint8_t*>(encoded_data) + hdr, encoded_data_size - hdr,
                                  output_data, output_capacity, CLDN_MEM_HOST);
  cldn_b200_decoder_destroy(dec);
  return rc == CLDN_OK ? static_cast<uint32_t>(decoded) : 0;
}

}  // extern "C"
