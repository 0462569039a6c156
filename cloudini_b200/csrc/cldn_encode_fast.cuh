// ENCODE of packed XYZI float32x4 clouds (point_step 16): the FAST FloatN kernel.
// Included by cldn_encode.cu (shares find_frame / status words; the build has no relocatable device code).
//
// Same bytes as FieldEncoderFloatN_Lossy::encode (cloudini_lib/src/field_encoder.cpp:42-91) + WriteStage1Chunk's framing
// (chunk_writer.cpp:27-48). Shape:
//  * persistent CTAs: CTA b takes the tiles (1024 points) b, b + G, ... of the launch's global tile order; a tile is four
//    warp quarters of 256 points with warp-private buffers, so the warps meet at ONE barrier per tile (the tile's size
//    and the quarters' offsets); every warp then resolves the decoupled look-back over the frame's tile status words for
//    itself and copies its own bytes out -- no barrier for the prefix, none for buffer reuse. (Measured alternatives:
//    one tile per CTA exposes the load latency at every CTA start; 256-point warp tiles without any barrier make the
//    look-back four times deeper and lose to it.)
//  * thread-blocked points: a lane owns 8 consecutive points, so the previous point is a register (no shuffle / select
//    per value) and the lane's bytes are ONE contiguous run of the output;
//  * the tile after next is already on its way while a tile is processed: cp.async (16 bytes per lane, straight into the
//    XOR-swizzled 16-byte slots that make the transposed reads conflict-free), together with its frame record;
//  * pass 1 (per value: FMUL, F2I, |s| tracking with max.NaN, delta, zigzag + 1, 7-bit groups -> bytes with two
//    add/mask steps, continuation flags from the top bit) keeps the finished LEB128 words in registers and sums their
//    lengths; one warp scan gives every lane its byte offset and the tile its size, which is published at once;
//  * pass 2 streams the words through a 64-bit register window and flushes aligned 32-bit words straight to their final
//    place in the staging buffer: a lane starts its window with the last bytes of its predecessor, so words shared by
//    two lanes are written once, whole;
//  * copy-out with 16-byte stores, chunk prefix back-patch by the chunk's last tile.
// Anything the 4-byte fast path cannot represent -- NaN / inf input, |v * mul| >= 2^25 (so that every delta fits 4
// varint bytes and no product reaches the x86 "integer indefinite" range), a partial tile -- sends the TILE to the exact
// byte-wise path below, which evaluates everything like the reference does.
#pragma once

namespace cldn {

constexpr int kET = 128;                  // threads per CTA (a container of 4 independent warps)
constexpr int kEW = kET / 32;
constexpr int kEP = 8;                    // points per lane and tile
constexpr int kEQuarterPts = 32 * kEP;    // 256 points per warp
constexpr int kETilePts = kEW * kEQuarterPts;       // 1024
constexpr int kEBufBytes = kEQuarterPts * 16 + 16;  // one input buffer = a warp's transposed points + the point in front of them
                                                    // (later: the warp's staged bytes)
constexpr int kEWarpBytes = 2 * kEBufBytes + 128;   // two buffers (the exact path stages up to 20 bytes per point across both) + slack
constexpr int kESmemBytes = kEW * kEWarpBytes;
constexpr int kEFrameCache = 32;          // frame records kept in shared memory per CTA

// cp.async (LDGSTS): 16 bytes global -> shared without a register round trip; groups complete in commit order
__device__ __forceinline__ void async_copy16(void* smem_dst, const void* gmem_src) {
#ifdef CLDN_CUSIM
  *reinterpret_cast<uint4*>(smem_dst) = *reinterpret_cast<const uint4*>(gmem_src);
#else
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
#endif
}
__device__ __forceinline__ void async_copy8(void* smem_dst, const void* gmem_src) {
#ifdef CLDN_CUSIM
  *reinterpret_cast<uint2*>(smem_dst) = *reinterpret_cast<const uint2*>(gmem_src);
#else
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gmem_src) : "memory");
#endif
}
__device__ __forceinline__ void async_commit() {
#ifndef CLDN_CUSIM
  asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
__device__ __forceinline__ void async_wait_all_but_last() {
#ifndef CLDN_CUSIM
  asm volatile("cp.async.wait_group 1;" ::: "memory");
#endif
}
__device__ __forceinline__ void async_wait_all() {
#ifndef CLDN_CUSIM
  asm volatile("cp.async.wait_group 0;" ::: "memory");
#endif
}
__device__ __forceinline__ void prefetch_l2_enc(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
__device__ __forceinline__ float max_nan(float a, float b) {  // NaN if either is NaN (fmaxf would drop it)
  float d;
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
  return d;
}
__device__ __forceinline__ uint32_t top_bit(uint32_t x) {  // index of the most significant set bit; 0xFFFFFFFF for 0
  uint32_t b;
  asm("bfind.u32 %0, %1;" : "=r"(b) : "r"(x));
  return b;
}
__device__ __forceinline__ uint32_t low_mask(uint32_t n) {  // (1 << n) - 1, all ones for n >= 32: one BMSK
  uint32_t d;
  asm("bmsk.clamp.b32 %0, %1, %2;" : "=r"(d) : "r"(0), "r"(n));
  return d;
}
__device__ __forceinline__ uint32_t bitselect_e(uint32_t m, uint32_t a, uint32_t b) {  // (a & m) | (b & ~m), one LOP3
  uint32_t d;
  asm("lop3.b32 %0, %1, %2, %3, 0xCA;" : "=r"(d) : "r"(m), "r"(a), "r"(b));
  return d;
}

// Exact byte-wise evaluation of one warp tile (lane l owns points 8 l .. 8 l + 7), like the reference
// (field_encoder.cpp:42-91). Returns the tile's byte count; the bytes are in `stage`. One full warp calls it.
__device__ __noinline__ uint32_t encode_warp_tile_careful(const EncFrame& F, const FloatNParams& P, uint32_t tile_p0, uint8_t* stage) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t step = P.point_step;
  uint32_t len[kEP];
  uint32_t mine = 0;
#pragma unroll 1
  for (int i = 0; i < kEP; ++i) {
    const uint32_t p = tile_p0 + lane * kEP + i;
    uint32_t l = 0;
    if (p < F.n_points) {
      const uint8_t* pt = F.in + static_cast<size_t>(p) * step;
      const uint8_t* prevp = (p % kChunkPoints) ? pt - step : nullptr;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float x = __uint_as_float(load_u32(pt + P.offset[k]));
        if (isnan(x)) { l += 1; continue; }
        const int32_t q = quant_i32_x86(x, P.mul[k]);
        int32_t pq = 0;
        if (prevp) {
          const float px = __uint_as_float(load_u32(prevp + P.offset[k]));
          if (!isnan(px)) pq = quant_i32_x86(px, P.mul[k]);
        }
        const int32_t d = static_cast<int32_t>(static_cast<uint32_t>(q) - static_cast<uint32_t>(pq));
        l += varint_len(zigzag_plus1(static_cast<int64_t>(d)));
      }
    }
    len[i] = l;
    mine += l;
  }
  uint32_t inc = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t up = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= static_cast<uint32_t>(d)) inc += up;
  }
  const uint32_t total = __shfl_sync(0xffffffffu, inc, 31);
  uint32_t off = inc - mine;
#pragma unroll 1
  for (int i = 0; i < kEP; ++i) {
    const uint32_t p = tile_p0 + lane * kEP + i;
    if (p < F.n_points) {
      const uint8_t* pt = F.in + static_cast<size_t>(p) * step;
      const uint8_t* prevp = (p % kChunkPoints) ? pt - step : nullptr;
      ByteSink bs{stage + off};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float x = __uint_as_float(load_u32(pt + P.offset[k]));
        if (isnan(x)) { bs.put_byte(0); continue; }
        const int32_t q = quant_i32_x86(x, P.mul[k]);
        int32_t pq = 0;
        if (prevp) {
          const float px = __uint_as_float(load_u32(prevp + P.offset[k]));
          if (!isnan(px)) pq = quant_i32_x86(px, P.mul[k]);
        }
        const int32_t d = static_cast<int32_t>(static_cast<uint32_t>(q) - static_cast<uint32_t>(pq));
        bs.put_varint(zigzag_plus1(static_cast<int64_t>(d)));
      }
      off += len[i];
    }
  }
  __syncwarp();
  return total;
}

// Warp version of copy_stage_to_global: `n` staged bytes (16-byte aligned shared memory, readable up to n + 16) to an
// arbitrarily aligned global address: 16-byte stores for the aligned body, byte stores for the <= 15-byte head and tail.
__device__ __forceinline__ void warp_copy_stage_to_global(const uint8_t* stage, uint32_t n, uint8_t* g) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t a = static_cast<uint32_t>(reinterpret_cast<uintptr_t>(g) & 15u);
  uint32_t head = (16u - a) & 15u;
  if (head > n) head = n;
  const uint32_t nvec = (n - head) >> 4;
  const uint32_t tail_begin = head + (nvec << 4);
  if (lane < head) g[lane] = stage[lane];
  if (lane >= 16u && lane - 16u < n - tail_begin) g[tail_begin + lane - 16u] = stage[tail_begin + lane - 16u];
  const uint32_t sh = (head & 3u) * 8u;
  const uint32_t* w = reinterpret_cast<const uint32_t*>(stage + (head & ~3u));
  uint4* gv = reinterpret_cast<uint4*>(g + head);
  for (uint32_t j = lane; j < nvec; j += 32u) {
    const uint32_t* q = w + 4 * j;
    const uint32_t w0 = q[0], w1 = q[1], w2 = q[2], w3 = q[3], w4 = q[4];
    gv[j] = make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh), __funnelshift_r(w3, w4, sh));
  }
}

__device__ __forceinline__ void tile_coords(const EncLaunch& L, uint32_t i, uint32_t* fi, uint32_t* t) {
  if (L.uniform_tiles) {
    *fi = i % L.n_frames;
    *t = i / L.n_frames;
  } else {
    *fi = find_frame(L.frames, L.n_frames, i);
    *t = i - L.frames[*fi].tile_begin;
  }
}

// All CTAs must be co-resident (the grid is sized by the occupancy query): a CTA spins on the sizes of tiles with a smaller
// global index, which belong to CTAs that are running.
struct EncFastShared {
  uint32_t wtot[2][kEW];        // bytes per warp quarter, by tile parity (no barrier separates consecutive tiles)
  unsigned long long excl[2];   // look-back result of the pending tile (by ITS parity), written by warp 0
};

// What a warp remembers of the tile whose bytes are staged but not yet copied out: the copy happens one tile later, behind
// the next tile's barrier, so that the look-back (warp 0) has a whole pass 1 to resolve instead of being waited for.
struct PendingQuarter {
  const uint8_t* stage;
  uint32_t wtot, wbase, total, fi, t, par;
  bool valid;
};

__device__ __forceinline__ void place_quarter(const EncLaunch& L, const EncFrame& F, const PendingQuarter& Q, uint64_t excl) {
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  constexpr uint32_t tiles_per_chunk = kChunkPoints / kETilePts;
  const uint32_t chunk = Q.t / tiles_per_chunk;
  uint8_t* payload = F.out + L.header_bytes;
  const uint64_t at = 4ull * (chunk + 1) + excl + Q.wbase;
  if (L.header_bytes + at + Q.wtot <= F.out_cap) warp_copy_stage_to_global(Q.stage, Q.wtot, payload + at);
  else if (lane == 0) report_error(L.err, DEV_ERR_ENCODE_OUTPUT_SMALL);
  if (Q.t == 0 && warp == 0 && L.header_bytes <= F.out_cap) {
    for (uint32_t k = lane; k < L.header_bytes; k += 32u) F.out[k] = L.header[k];
  }
  const bool last_of_frame = (Q.t + 1 == F.n_tiles);
  const bool last_of_chunk = last_of_frame || ((Q.t + 1) % tiles_per_chunk == 0);
  if (last_of_chunk && warp == kEW - 1 && lane == 0) {
    const uint32_t first = chunk * tiles_per_chunk;
    const uint64_t data_before_chunk = (first == 0) ? 0 : wait_inclusive(L.status, F.tile_begin + first - 1, L.epoch);
    const uint64_t body = (excl + Q.total) - data_before_chunk;
    if (L.header_bytes + 4ull * (chunk + 1) + data_before_chunk <= F.out_cap) {
      store_u32(payload + 4ull * chunk + data_before_chunk, static_cast<uint32_t>(body));  // chunk_writer.cpp:33-40
    }
    if (last_of_frame) L.sizes[Q.fi] = L.header_bytes + 4ull * F.n_chunks + excl + Q.total;
  }
}

__global__ void __launch_bounds__(kET, 6) encode_xyzi_fast_kernel(const EncLaunch L, const FloatNParams P) {
  extern __shared__ __align__(16) uint8_t dyn_smem[];
  __shared__ EncFastShared sh;
  __shared__ EncFrame s_F[kEW][3];             // per warp: records of frames beyond the cache (ring of 3: next, current, pending)
  __shared__ EncFrame s_cache[kEFrameCache];   // the batch's first frames (all of them for the usual batch sizes)
  __shared__ FloatNParams s_P;
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const bool aligned16 = (L.flags & kEncInputsAligned16) != 0u;
  const uint32_t G = gridDim.x;
  uint8_t* const wsm = dyn_smem + warp * kEWarpBytes;

  if (blockIdx.x == 0) handle_empty_frames(L);
  uint32_t i = blockIdx.x;
  uint32_t fi = 0, t = 0;
  if (i < L.n_tiles_total) tile_coords(L, i, &fi, &t);
  if (threadIdx.x == 0) s_P = P;
  if (lane == 0 && i < L.n_tiles_total && fi >= kEFrameCache) s_F[warp][0] = L.frames[fi];
  {
    const uint32_t n_cached = min(L.n_frames, static_cast<uint32_t>(kEFrameCache));
    const unsigned long long* src = reinterpret_cast<const unsigned long long*>(L.frames);
    unsigned long long* dst = reinterpret_cast<unsigned long long*>(s_cache);
    for (uint32_t k = threadIdx.x; k < n_cached * (sizeof(EncFrame) / 8); k += kET) dst[k] = src[k];
  }
  __syncthreads();  // frame records and the field table are in shared memory
  if (i >= L.n_tiles_total) return;

  // asks for this warp's 256 points of tile tt (frame input `in`, `np` points) into `buf` (transposed slots, + the point
  // in front of them in slot 256); false if the tile is not eligible
  auto prefetch = [&](const uint8_t* in, uint32_t np, uint32_t tt, uint8_t* buf) -> bool {
    const uint32_t p0 = tt * kETilePts;
    if (!aligned16 || p0 + kETilePts > np) return false;
    const uint32_t q0 = p0 + warp * kEQuarterPts;
    const uint4* src = reinterpret_cast<const uint4*>(in) + q0 + lane;
    uint4* wsl = reinterpret_cast<uint4*>(buf);
#pragma unroll
    for (int k = 0; k < kEP; ++k) {
      const uint32_t q = 32u * k + lane, ol = q >> 3;
      async_copy16(wsl + 8 * ol + ((q & 7u) ^ (ol & 7u)), src + 32 * k);
    }
    if (lane == 0 && (q0 % kChunkPoints) != 0) async_copy16(wsl + kEQuarterPts, reinterpret_cast<const uint4*>(in) + q0 - 1);
    return true;
  };
  bool have_cur;
  {
    const EncFrame& F0 = fi < kEFrameCache ? s_cache[fi] : s_F[warp][0];
    have_cur = prefetch(F0.in, F0.n_points, t, wsm);
  }
  async_commit();

  PendingQuarter pend;
  pend.valid = false;
  pend.stage = nullptr; pend.wtot = pend.wbase = pend.total = pend.fi = pend.t = pend.par = 0;
  LookbackPoll lb;   // warp 0: the look-back of the pending tile
  lb.done = true; lb.idx = 0; lb.exclusive = 0; lb.s = 0;
  uint32_t pend_tile = 0, pend_first = 0;  // warp 0: status indices of the pending tile

  uint32_t it = 0;
  for (; i < L.n_tiles_total; i += G, ++it) {
    const uint32_t cur = it & 1u, ring = it % 3u;
    uint8_t* buf = wsm + cur * kEBufBytes;                 // this quarter's transposed input, then its staged output
    uint8_t* stage = buf;
    // ---- which tile comes next (its points are asked into L2 now, into shared memory once the other buffer is free) ----
    const uint32_t nxt = i + G;
    uint32_t nfi = 0, nt = 0, np = 0;
    const uint8_t* nin = nullptr;
    if (nxt < L.n_tiles_total) {
      tile_coords(L, nxt, &nfi, &nt);
      if (nfi < kEFrameCache) {
        np = s_cache[nfi].n_points;
        nin = s_cache[nfi].in;
      } else {  // large batches: the record travels with the tile's points, its two words needed now come straight from L2
        const EncFrame* NF = L.frames + nfi;
        np = __ldg(&NF->n_points);
        nin = reinterpret_cast<const uint8_t*>(__ldg(reinterpret_cast<const unsigned long long*>(&NF->in)));
      }
      if (static_cast<uint64_t>(nt) * kETilePts + kETilePts <= np) {
        prefetch_l2_enc(nin + (static_cast<size_t>(nt) * kETilePts + warp * kEQuarterPts) * 16u + lane * 128u);
      }
    }
    async_wait_all();            // this tile's copies have landed
    __syncwarp();
    const EncFrame& F = fi < kEFrameCache ? s_cache[fi] : s_F[warp][ring];

    const uint32_t tile = F.tile_begin + t;
    const uint32_t tile_p0 = t * kETilePts;
    const uint32_t q0 = tile_p0 + warp * kEQuarterPts;     // my warp's first point
    const bool full = tile_p0 + kETilePts <= F.n_points;
    bool fast = full;
    uint32_t X[kEP][4];
    uint32_t mine = 0;
    if (full) {
      uint4* wsl = reinterpret_cast<uint4*>(buf);
      if (!have_cur) {
        // ---- load + transpose: lane l of iteration k loads point 32 k + l of the quarter ----
#pragma unroll
        for (int k = 0; k < kEP; ++k) {
          const uint32_t q = 32u * k + lane;
          uint4 v;
          if (aligned16) {
            v = __ldcs(reinterpret_cast<const uint4*>(F.in) + q0 + q);
          } else {
            const uint8_t* pt = F.in + static_cast<size_t>(q0 + q) * 16u;
            v = make_uint4(load_u32(pt), load_u32(pt + 4), load_u32(pt + 8), load_u32(pt + 12));
          }
          const uint32_t ol = q >> 3;  // owner lane; slot of point j of lane l: 8 l + (j ^ (l & 7))
          wsl[8 * ol + ((q & 7u) ^ (ol & 7u))] = v;
        }
        if (lane == 0 && (q0 % kChunkPoints) != 0) {
          const uint8_t* pt = F.in + static_cast<size_t>(q0 - 1) * 16u;
          wsl[kEQuarterPts] = make_uint4(load_u32(pt), load_u32(pt + 4), load_u32(pt + 8), load_u32(pt + 12));
        }
        __syncwarp();
      }
      // previous point of my first point: 0 at a chunk start, else quantised like any point
      uint4 pvu = make_uint4(0, 0, 0, 0);
      if (lane != 0) {
        const uint32_t pl = lane - 1;
        pvu = wsl[8 * pl + (7u ^ (pl & 7u))];
      } else if ((q0 % kChunkPoints) != 0) {
        pvu = wsl[kEQuarterPts];
      }
      float trk = 0.0f;
      int32_t prev[4];
      {
        const float pf[4] = {__uint_as_float(pvu.x), __uint_as_float(pvu.y), __uint_as_float(pvu.z), __uint_as_float(pvu.w)};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float s = __fmul_rn(pf[k], P.mul[k]);
          trk = max_nan(trk, fabsf(s));
          prev[k] = __float2int_rn(s);
        }
      }
      // ---- pass 1: LEB128 words of my 32 values + their total length ----
      const uint4* my_slots = wsl + 8 * lane;
      const uint32_t lx = lane & 7u;
#pragma unroll
      for (int j = 0; j < kEP; ++j) {
        const uint4 u = my_slots[j ^ lx];
        const float pf[4] = {__uint_as_float(u.x), __uint_as_float(u.y), __uint_as_float(u.z), __uint_as_float(u.w)};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float s = __fmul_rn(pf[k], P.mul[k]);   // _mm_mul_ps: IEEE RN, never contracted
          trk = max_nan(trk, fabsf(s));
          const int32_t q = __float2int_rn(s);          // cvtps2dq under the default MXCSR: ties to even
          const uint32_t d = static_cast<uint32_t>(q) - static_cast<uint32_t>(prev[k]);
          prev[k] = q;
          const uint32_t zz1 = ((d << 1) ^ static_cast<uint32_t>(static_cast<int32_t>(d) >> 31)) + 1u;  // < 2^28 on the fast path
          // 7-bit groups -> bytes: 14-bit halves into 16-bit lanes, then the upper 7 bits of each lane one bit up
          uint32_t x = (zz1 & 0xFFFFC000u) * 3u + zz1;           // lo14 + hi14 * 2^16 (one LOP3 + one IMAD)
          x = x + (x & 0x3F803F80u);
          const uint32_t b = top_bit(x);                           // inside the value's last byte (garbage tiles: x may be 0)
          x |= low_mask(b) & 0x00808080u;                          // continuation flags on every byte below it
          X[j][k] = x;
          mine += b >> 3;
        }
      }
      mine += kEP * 4;
      fast = trk < 33554432.0f;  // 2^25; false for NaN
    }
    // ---- offsets inside the quarter (warp scan); the quarters' sizes meet at the tile's one barrier ----
    uint32_t inc = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t up = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= static_cast<uint32_t>(d)) inc += up;
    }
    uint32_t wtot = __shfl_sync(0xffffffffu, inc, 31);
    if (lane == 0) sh.wtot[cur][warp] = wtot;
    // warp 0: the pending tile's look-back has had this whole pass to resolve; its result travels through the barrier
    if (warp == 0 && pend.valid) {
      const uint64_t e = lb.finish(L.status, pend_first, pend_tile, L.epoch, pend.total);
      if (lane == 0) sh.excl[pend.par] = e;
    }
    const int any_slow = __syncthreads_or(fast ? 0 : 1);
    // ---- the pending tile's quarter leaves now: its buffer is the one the next tile's points go into ----
    if (pend.valid) {
      const EncFrame& PF = pend.fi < kEFrameCache ? s_cache[pend.fi] : s_F[warp][(it + 2u) % 3u];
      place_quarter(L, PF, pend, sh.excl[pend.par]);
      pend.valid = false;
      __syncwarp();
    }
    bool have_next = false;
    if (nxt < L.n_tiles_total && nfi >= kEFrameCache && lane < sizeof(EncFrame) / 8) {  // the next tile's frame record
      async_copy8(reinterpret_cast<uint8_t*>(&s_F[warp][(it + 1u) % 3u]) + 8 * lane, reinterpret_cast<const uint8_t*>(L.frames + nfi) + 8 * lane);
    }
    if (any_slow) {
      // exact path: up to 20 bytes per point, staged across BOTH of the warp's buffers; the next tile is loaded
      // synchronously when its turn comes
      stage = wsm;
      wtot = encode_warp_tile_careful(F, s_P, q0, stage);
      if (lane == 0) sh.wtot[cur][warp] = wtot;   // (the fast sizes are not read on this path)
      __syncthreads();
    } else if (nxt < L.n_tiles_total) {
      have_next = prefetch(nin, np, nt, wsm + (cur ^ 1u) * kEBufBytes);
    }
    async_commit();
    uint32_t wbase = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kEW; ++w) {
      const uint32_t c = sh.wtot[cur][w];
      if (static_cast<uint32_t>(w) < warp) wbase += c;
      total += c;
    }
    // the tile's size is final: warp 0 publishes it before the bytes are packed and starts the look-back
    if (warp == 0) {
      lb.begin(L.status, F.tile_begin, tile, L.epoch, total);
      lb.issue(L.status, F.tile_begin, L.epoch);
      pend_tile = tile;
      pend_first = F.tile_begin;
    }
    if (!any_slow) {
      const uint32_t off = inc - mine;
      __syncwarp();  // every lane is done with the transposed input the staged bytes are about to overwrite
      // ---- pass 2: 64-bit window, aligned word flushes. The quarter's bytes start at staging offset 0. ----
      // `bit` = bit position of the next byte in the quarter's stream; the window's low word is the aligned word holding it
      uint32_t bit = 8u * off;
      // a lane starts its window with the last off % 4 bytes of its predecessor: the last 4 bytes of every lane's run
      // (top byte = most recent) travel one lane up
      uint32_t lo;
      {
        uint32_t t4 = 0;
        t4 = __funnelshift_rc(t4, X[kEP - 1][1], (top_bit(X[kEP - 1][1]) & 0x18u) + 8u);
        t4 = __funnelshift_rc(t4, X[kEP - 1][2], (top_bit(X[kEP - 1][2]) & 0x18u) + 8u);
        t4 = __funnelshift_rc(t4, X[kEP - 1][3], (top_bit(X[kEP - 1][3]) & 0x18u) + 8u);
        uint32_t ptail = __shfl_up_sync(0xffffffffu, t4, 1);
        if (lane == 0) ptail = 0u;
        lo = __funnelshift_rc(ptail, 0u, 32u - (bit & 31u));
      }
      uint32_t wa = (bit >> 3) & ~3u;                                 // byte address of that word in the staging buffer
#pragma unroll
      for (int j = 0; j < kEP; ++j) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t x = X[j][k];
          const uint32_t b = top_bit(x);
          lo |= __funnelshift_l(0u, x, bit);                 // x << (bit % 32)
          const uint32_t hi = __funnelshift_l(x, 0u, bit);   // x >> (32 - bit % 32), 0 for bit % 32 == 0
          bit += (b & 0x18u) + 8u;
          const uint32_t wn = (bit >> 3) & ~3u;
          if (wn != wa) { *reinterpret_cast<uint32_t*>(stage + wa) = lo; lo = hi; }
          wa = wn;
        }
      }
      if (lane == 31u && (bit & 31u) != 0u) *reinterpret_cast<uint32_t*>(stage + wa) = lo;  // nobody follows the quarter's last lane
    }
    __syncwarp();  // staged bytes complete
    pend.valid = true;
    pend.stage = stage; pend.wtot = wtot; pend.wbase = wbase; pend.total = total; pend.fi = fi; pend.t = t; pend.par = cur;
    if (any_slow) {
      // the exact path's bytes lie across both buffers: they leave at once (the next tile needs one of them)
      if (warp == 0) {
        const uint64_t e = lb.finish(L.status, pend_first, pend_tile, L.epoch, total);
        if (lane == 0) sh.excl[cur] = e;
      }
      __syncthreads();
      place_quarter(L, F, pend, sh.excl[cur]);
      pend.valid = false;
      __syncwarp();
    }
    have_cur = have_next;
    fi = nfi;
    t = nt;
  }
  // ---- the last tile of this CTA is still pending ----
  async_wait_all();
  if (warp == 0 && pend.valid) {
    const uint64_t e = lb.finish(L.status, pend_first, pend_tile, L.epoch, pend.total);
    if (lane == 0) sh.excl[pend.par] = e;
  }
  __syncthreads();
  if (pend.valid) {
    const EncFrame& PF = pend.fi < kEFrameCache ? s_cache[pend.fi] : s_F[warp][(it + 2u) % 3u];
    place_quarter(L, PF, pend, sh.excl[pend.par]);
  }
}

static bool encode_fast_enabled() {
  const char* e = getenv("CLDN_B200_ENC_FAST");
  return !(e && e[0] == '0');
}
// Plans / batches the fast kernel takes: exactly one FloatN op over packed float32x4 at offsets 0,4,8,12 of a 16-byte point.
bool encode_fast_applies(const Plan& plan) {
  if (!plan.floatn_only || !encode_fast_enabled()) return false;
  const RegOp& op = plan.ops[0];
  if (!(op.lanes == 4 && plan.point_step == 16 && op.offset[0] == 0 && op.offset[1] == 4 && op.offset[2] == 8 && op.offset[3] == 12)) return false;
  for (int k = 0; k < 4; ++k) {
    const float m = op.enc_mul_f[k];
    if (!(m > 0.0f) || m > 3.0e38f) return false;  // the |s| < 2^25 test must imply a finite, non-NaN input
  }
  return true;
}
uint32_t encode_fast_tile_points() { return kETilePts; }

static int launch_encode_fast(const Plan& plan, const EncLaunch& L, cudaStream_t stream) {
  FloatNParams P;
  for (int k = 0; k < 4; ++k) {
    P.offset[k] = plan.ops[0].offset[k];
    P.mul[k] = plan.ops[0].enc_mul_f[k];
  }
  P.point_step = plan.point_step;
  const size_t smem = kESmemBytes;
  auto k = encode_xyzi_fast_kernel;
  if (set_smem(k, smem) != cudaSuccess) return -1;
  // persistent grid: every CTA must be resident (they wait for each other's tile sizes)
  static int resident = 0;
  if (resident == 0) {
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, kET, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    resident = std::max(1, sms) * per_sm;
  }
  const uint32_t grid = std::min<uint32_t>(L.n_tiles_total, static_cast<uint32_t>(resident));
  k<<<grid, kET, smem, stream>>>(L, P);
  count_launch();
  return 1;
}

}  // namespace cldn
