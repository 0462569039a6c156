// Host <-> device copies of PAGEABLE caller buffers (numpy arrays, ROS message payloads, std::vector).
//
// cudaMemcpyAsync on pageable memory is staged by the driver through its own pinned buffers with ONE copying thread:
// ~10 GB/s on the GPU boxes, i.e. 1.6 ms for the 16 MB of a 1M-point message -- three times the GPU work of the whole
// converter step. Here the same staging is done by the library: a small ring of pinned slots per calling thread, the
// host-side memcpy of every slot split over a few worker threads (they inherit the caller's CPU affinity, so after
// cldn_b200_bind_host_thread_to_device they run on the GPU's NUMA node), the DMA of slot i overlapping the memcpy of
// slot i + 1. Pinned caller buffers never come here (cudaPointerGetAttributes says what a pointer is).
//
// Semantics match the pageable cudaMemcpyAsync they replace: staged_h2d returns when the source has been consumed (the
// DMA may still be in flight on the stream); staged_d2h returns when the destination is complete.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

#include "cldn_kernels.h"

namespace cldn {
namespace {

constexpr size_t kSlotBytes = 4u << 20;
constexpr int kSlots = 3;
constexpr size_t kMinStaged = 1u << 20;      // below this the driver's path is as good
constexpr size_t kMinPart = 256u << 10;      // a worker's share of one slot

struct CopyTask {
  void* dst;
  const void* src;
  size_t n;
  std::atomic<int>* left;
};

class CopyPool {
 public:
  static CopyPool& get() {
    static CopyPool* pool = new CopyPool();   // never destroyed: worker threads must not outlive their queue at exit
    return *pool;
  }
  // copies n bytes with the caller + up to n_workers_ helpers; returns when all of it is done
  void copy(void* dst, const void* src, size_t n) {
    size_t parts = n / kMinPart;
    if (parts > workers_.size() + 1) parts = workers_.size() + 1;
    if (parts <= 1) { memcpy(dst, src, n); return; }
    const size_t share = ((n / parts) + 63) & ~size_t(63);
    std::atomic<int> left{static_cast<int>(parts) - 1};
    {
      std::lock_guard<std::mutex> g(m_);
      for (size_t p = 1; p < parts; ++p) {
        const size_t off = p * share;
        const size_t len = p + 1 == parts ? n - off : share;
        q_.push_back(CopyTask{static_cast<uint8_t*>(dst) + off, static_cast<const uint8_t*>(src) + off, len, &left});
      }
    }
    cv_.notify_all();
    memcpy(dst, src, share);
    while (left.load(std::memory_order_acquire) > 0) std::this_thread::yield();
  }

 private:
  CopyPool() {
    int n = 4;
    if (const char* e = getenv("CLDN_B200_COPY_THREADS")) n = atoi(e);
    if (n < 0) n = 0;
    if (n > 16) n = 16;
    for (int i = 0; i < n; ++i) {
      workers_.emplace_back([this] { run(); });
      workers_.back().detach();
    }
  }
  void run() {
    for (;;) {
      CopyTask t;
      {
        std::unique_lock<std::mutex> g(m_);
        cv_.wait(g, [this] { return !q_.empty(); });
        t = q_.front();
        q_.pop_front();
      }
      memcpy(t.dst, t.src, t.n);
      t.left->fetch_sub(1, std::memory_order_release);
    }
  }
  std::mutex m_;
  std::condition_variable cv_;
  std::deque<CopyTask> q_;
  std::vector<std::thread> workers_;
};

struct Ring {
  uint8_t* slot[kSlots] = {nullptr, nullptr, nullptr};
  cudaEvent_t ev[kSlots] = {nullptr, nullptr, nullptr};
  bool busy[kSlots] = {false, false, false};
  int device = -1;
  bool ok = false;
  bool ensure() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return false;
    if (ok && dev == device) return true;
    release();
    for (int i = 0; i < kSlots; ++i) {
      if (cudaMallocHost(reinterpret_cast<void**>(&slot[i]), kSlotBytes) != cudaSuccess) { release(); return false; }
      if (cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming) != cudaSuccess) { release(); return false; }
    }
    device = dev;
    ok = true;
    return true;
  }
  void release() {
    for (int i = 0; i < kSlots; ++i) {
      if (ev[i]) { cudaEventDestroy(ev[i]); ev[i] = nullptr; }
      if (slot[i]) { cudaFreeHost(slot[i]); slot[i] = nullptr; }
      busy[i] = false;
    }
    ok = false;
  }
  ~Ring() { /* process / thread exit: the context may be gone already; the driver reclaims pinned memory */ }
};

Ring& ring() {
  static thread_local Ring r;
  return r;
}

bool staging_enabled() {
  static const bool on = [] { const char* e = getenv("CLDN_B200_STAGED_COPIES"); return !(e && e[0] == '0'); }();
  return on;
}

}  // namespace

bool host_is_pageable(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();  // older runtimes report unregistered host memory as an error: clear it
    return true;
  }
  return a.type == cudaMemoryTypeUnregistered;
}

int copy_h2d(void* dst_dev, const void* src_host, size_t bytes, cudaStream_t s) {
  if (bytes == 0) return 0;
  if (bytes < kMinStaged || !staging_enabled() || !host_is_pageable(src_host) || !ring().ensure()) {
    return cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, s) == cudaSuccess ? 0 : -1;
  }
  Ring& r = ring();
  CopyPool& pool = CopyPool::get();
  size_t off = 0;
  for (int i = 0; off < bytes; i = (i + 1) % kSlots) {
    const size_t len = bytes - off < kSlotBytes ? bytes - off : kSlotBytes;
    if (r.busy[i] && cudaEventSynchronize(r.ev[i]) != cudaSuccess) return -1;   // the DMA out of this slot has finished
    pool.copy(r.slot[i], static_cast<const uint8_t*>(src_host) + off, len);
    if (cudaMemcpyAsync(static_cast<uint8_t*>(dst_dev) + off, r.slot[i], len, cudaMemcpyHostToDevice, s) != cudaSuccess) return -1;
    if (cudaEventRecord(r.ev[i], s) != cudaSuccess) return -1;
    r.busy[i] = true;
    off += len;
  }
  return 0;
}

int copy_d2h(void* dst_host, const void* src_dev, size_t bytes, cudaStream_t s) {
  if (bytes == 0) return 0;
  if (bytes < kMinStaged || !staging_enabled() || !host_is_pageable(dst_host) || !ring().ensure()) {
    return cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, s) == cudaSuccess ? 0 : -1;
  }
  Ring& r = ring();
  CopyPool& pool = CopyPool::get();
  const size_t n_pieces = (bytes + kSlotBytes - 1) / kSlotBytes;
  auto piece_len = [&](size_t p) { return p + 1 == n_pieces ? bytes - p * kSlotBytes : kSlotBytes; };
  // every slot may still be the source of an earlier upload on another stream
  for (int i = 0; i < kSlots; ++i) {
    if (r.busy[i]) { if (cudaEventSynchronize(r.ev[i]) != cudaSuccess) return -1; r.busy[i] = false; }
  }
  size_t issued = 0;
  for (size_t p = 0; p < n_pieces; ++p) {
    while (issued < n_pieces && issued < p + kSlots) {     // keep the DMA engine kSlots pieces ahead of the host copies
      const int i = static_cast<int>(issued % kSlots);
      if (cudaMemcpyAsync(r.slot[i], static_cast<const uint8_t*>(src_dev) + issued * kSlotBytes, piece_len(issued), cudaMemcpyDeviceToHost, s) != cudaSuccess) return -1;
      if (cudaEventRecord(r.ev[i], s) != cudaSuccess) return -1;
      ++issued;
    }
    const int i = static_cast<int>(p % kSlots);
    if (cudaEventSynchronize(r.ev[i]) != cudaSuccess) return -1;
    pool.copy(static_cast<uint8_t*>(dst_host) + p * kSlotBytes, r.slot[i], piece_len(p));
  }
  return 0;
}

}  // namespace cldn
