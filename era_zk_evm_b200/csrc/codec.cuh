// Transport encoder of the witness streams (include/zkb_codec.h): XOR-with-prediction + presence bitmap, one warp per VM.
//
// The canonical streams stay in HBM (they are what the device-side consumers read); this kernel is the last step before
// the PCIe link: it reads every record once per pass (HBM-read bound: two coalesced 128-byte loads per cycle row) and
// writes ~30 % of the bytes.  Two passes over the same code (template WRITE): pass 1 only counts the encoded bytes of
// every (VM, stream), a scan turns the counts into byte offsets, pass 2 writes -- so the blob layout is deterministic
// (bit-identical to the scalar encoder in zkb_codec.h, which the tests check) and needs no atomics.
#pragma once
#include <stdint.h>

#define ZKB_CODEC_NO_HOST
#include "../../include/zkb_codec.h"
#include "vm.cuh"

namespace zkb {

struct EncArgs {
  uint32_t* sizes;          // [6][n_vms] encoded bytes per VM and stream (pass 1 out, scan in)
  uint64_t* offsets;        // [6][n_vms + 1] exclusive prefix sums (scan out, pass 2 in)
  uint64_t* totals;         // [8] mapped pinned host memory: payload bytes per stream, [6] = raw canonical bytes
  uint8_t* blob;            // pass 2: the device blob
  uint64_t counts_offset, offsets_offset;
  uint64_t payload_offset[ZKB_N_STREAMS];
  uint32_t kinds_mask;      // streams outside the mask are left out of the blob (their counts are still reported)
};

__device__ __forceinline__ uint32_t lanemask_lt() {
  uint32_t m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

// cycle rows: lane l owns words l and l + 32 of the row (two coalesced 128-byte loads per row)
template <bool WRITE>
__device__ __forceinline__ uint64_t encode_rows(const uint32_t* __restrict__ rows, uint32_t n, uint32_t* __restrict__ out, uint32_t lane) {
  uint32_t prev_lo = 0, prev_hi = 0;
  uint64_t words = 0;
  const uint32_t lt = lanemask_lt();
  uint32_t nx_lo = 0, nx_hi = 0;
  if (n) {
    nx_lo = __ldcs(rows + lane);
    nx_hi = __ldcs(rows + 32 + lane);
  }
  for (uint32_t r = 0; r < n; r++) {
    const uint32_t w_lo = nx_lo, w_hi = nx_hi;
    if (r + 1 < n) {  // next row's loads in flight while this one is encoded
      nx_lo = __ldcs(rows + (size_t)(r + 1) * 64 + lane);
      nx_hi = __ldcs(rows + (size_t)(r + 1) * 64 + 32 + lane);
    }
    const uint32_t w2 = __shfl_sync(0xffffffffu, w_lo, 2);
    const uint32_t vidx = w2 & ((1u << ZK_VARIANT_BITS) - 1u);
    uint32_t pred_lo = 0;
    if (lane == 0) pred_lo = prev_lo + 1u;
    else if (lane == 1) pred_lo = prev_lo + ZK_TIME_DELTA_PER_CYCLE;
    else if (lane == 4) pred_lo = vidx | 1u << 16;
    else if (lane == 5) {
      const uint32_t p = prev_lo >> 16;
      pred_lo = p | ((p + 1u) & 0xFFFFu) << 16;
    } else if (lane == 6) pred_lo = prev_lo;
    else if (lane == 7) pred_lo = prev_lo - ZK_OPCODE_PRICES[vidx];
    const uint32_t pred_hi = (lane < 8 || lane == 11) ? 0u : prev_hi;  // words 32..39 operands, 43 per-cycle counts
    const uint32_t x_lo = w_lo ^ pred_lo, x_hi = w_hi ^ pred_hi;
    const uint32_t m_lo = __ballot_sync(0xffffffffu, x_lo != 0), m_hi = __ballot_sync(0xffffffffu, x_hi != 0);
    const uint32_t n_lo = __popc(m_lo), n_hi = __popc(m_hi);
    if (WRITE) {
      uint32_t* o = out + words;
      if (lane == 0) o[0] = m_lo;
      if (lane == 1) o[1] = m_hi;
      if (x_lo) o[2 + __popc(m_lo & lt)] = x_lo;
      if (x_hi) o[2 + n_lo + __popc(m_hi & lt)] = x_hi;
    }
    words += 2u + n_lo + n_hi;
    prev_lo = w_lo;
    prev_hi = w_hi;
  }
  return words * 4;
}

// 12-word records (MEM, DECOMMIT): two records per step, one per half-warp
template <bool WRITE, int KIND>
__device__ __forceinline__ uint64_t encode_rec12(const uint32_t* __restrict__ recs, uint32_t n, uint32_t* __restrict__ out, uint32_t lane) {
  const uint32_t g = lane >> 4, li = lane & 15u;
  uint32_t w_last = 0;
  uint64_t words = 0;
  for (uint32_t r0 = 0; r0 < n; r0 += 2) {
    const uint32_t r = r0 + g;
    const bool valid = r < n && li < 12;
    const uint32_t w = valid ? __ldcs(recs + (size_t)r * 12 + li) : 0u;
    const uint32_t a = __shfl_sync(0xffffffffu, w, li), b = __shfl_sync(0xffffffffu, w_last, 16 + li);
    const uint32_t p = g ? a : b;  // the same word of the previous record
    uint32_t pred;
    if (KIND == ZKB_STREAM_MEM) pred = li == 2 ? p + 1u : li < 4 ? p : 0u;
    else pred = li < 4 ? p : 0u;
    const uint32_t x = valid ? (w ^ pred) : 0u;
    const uint32_t m = __ballot_sync(0xffffffffu, x != 0);
    const uint32_t m0 = m & 0xFFFFu, m1 = m >> 16;
    const uint32_t s0 = 1u + __popc(m0), s1 = (r0 + 1 < n) ? 1u + __popc(m1) : 0u;
    if (WRITE) {
      uint32_t* o = out + words + (g ? s0 : 0u);
      const uint32_t mine = g ? m1 : m0;
      if (li == 0 && r < n) o[0] = mine;
      if (x) o[1 + __popc(mine & ((1u << li) - 1u))] = x;
    }
    words += s0 + s1;
    w_last = w;
  }
  return words * 4;
}

// 32-word records (LOG, FRAME): lane l owns word l
template <bool WRITE, int KIND>
__device__ __forceinline__ uint64_t encode_rec32(const uint32_t* __restrict__ recs, uint32_t n, uint32_t* __restrict__ out, uint32_t lane) {
  uint32_t prev = 0;
  uint64_t words = 0;
  const uint32_t lt = lanemask_lt();
  for (uint32_t r = 0; r < n; r++) {
    const uint32_t w = __ldcs(recs + (size_t)r * 32 + lane);
    const uint32_t pred = (KIND == ZKB_STREAM_FRAME || lane < 8) ? prev : 0u;
    const uint32_t x = w ^ pred;
    const uint32_t m = __ballot_sync(0xffffffffu, x != 0);
    if (WRITE) {
      uint32_t* o = out + words;
      if (lane == 0) o[0] = m;
      if (x) o[1 + __popc(m & lt)] = x;
    }
    words += 1u + __popc(m);
    prev = w;
  }
  return words * 4;
}

template <bool WRITE>
__global__ void __launch_bounds__(256) zkb_encode_kernel(const DevBatch B, const EncArgs A) {
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t n_vms = B.n_vms;
  for (uint32_t vm = blockIdx.x * 8 + warp; vm < n_vms; vm += gridDim.x * 8) {
    const uint32_t* x = B.hot[vm].x;
    uint32_t cnt[ZKB_N_STREAMS];
#pragma unroll
    for (int k = 0; k < ZKB_N_STREAMS; k++) cnt[k] = ((A.kinds_mask >> k) & 1u) ? x[X_COUNT0 + k] : 0u;
    uint32_t* outp[ZKB_N_STREAMS];
#pragma unroll
    for (int k = 0; k < ZKB_N_STREAMS; k++) outp[k] = nullptr;
    if (WRITE) {
      // the blob's tables: per-VM summary (same 8 words the host mirror holds) and this VM's offsets
      uint32_t* counts = reinterpret_cast<uint32_t*>(A.blob + A.counts_offset) + (size_t)vm * 8;
      const uint32_t status = x[X_STATUS];
      const uint32_t st_out = (status == ZKB_VM_RUNNING && B.hot[vm].live[L_DEPTH - 40] == 0 && x[X_CYCLE] > 0) ? (uint32_t)ZKB_VM_ENDED : status;
      if (lane < 8) counts[lane] = lane < 6 ? x[X_COUNT0 + lane] : lane == 6 ? st_out : x[X_CYCLE];
      uint64_t* offs = reinterpret_cast<uint64_t*>(A.blob + A.offsets_offset);
      if (lane < ZKB_N_STREAMS) {
        offs[(size_t)lane * (n_vms + 1) + vm] = A.offsets[(size_t)lane * (n_vms + 1) + vm];
        if (vm == n_vms - 1) offs[(size_t)lane * (n_vms + 1) + n_vms] = A.offsets[(size_t)lane * (n_vms + 1) + n_vms];
      }
#pragma unroll
      for (int k = 0; k < ZKB_N_STREAMS; k++)
        outp[k] = reinterpret_cast<uint32_t*>(A.blob + A.payload_offset[k] + A.offsets[(size_t)k * (n_vms + 1) + vm]);
    }
    uint64_t sz[ZKB_N_STREAMS];
    const uint32_t* s0 = reinterpret_cast<const uint32_t*>(B.streams[ZKB_STREAM_ROWS] + (size_t)vm * B.cap[ZKB_STREAM_ROWS] * ZKB_ROW_BYTES);
    sz[0] = encode_rows<WRITE>(s0, cnt[0], outp[0], lane);
    const uint32_t* s1 = reinterpret_cast<const uint32_t*>(B.streams[ZKB_STREAM_MEM] + (size_t)vm * B.cap[ZKB_STREAM_MEM] * ZKB_MEM_BYTES);
    sz[1] = encode_rec12<WRITE, ZKB_STREAM_MEM>(s1, cnt[1], outp[1], lane);
    const uint32_t* s2 = reinterpret_cast<const uint32_t*>(B.streams[ZKB_STREAM_LOG] + (size_t)vm * B.cap[ZKB_STREAM_LOG] * ZKB_LOG_BYTES);
    sz[2] = encode_rec32<WRITE, ZKB_STREAM_LOG>(s2, cnt[2], outp[2], lane);
    const uint32_t* s3 = reinterpret_cast<const uint32_t*>(B.streams[ZKB_STREAM_DECOMMIT] + (size_t)vm * B.cap[ZKB_STREAM_DECOMMIT] * ZKB_DECOMMIT_BYTES);
    sz[3] = encode_rec12<WRITE, ZKB_STREAM_DECOMMIT>(s3, cnt[3], outp[3], lane);
    const uint32_t* s4 = reinterpret_cast<const uint32_t*>(B.streams[ZKB_STREAM_FRAME] + (size_t)vm * B.cap[ZKB_STREAM_FRAME] * ZKB_FRAME_BYTES);
    sz[4] = encode_rec32<WRITE, ZKB_STREAM_FRAME>(s4, cnt[4], outp[4], lane);
    // REFUND: raw 8-byte records
    sz[5] = (uint64_t)cnt[5] * ZKB_REFUND_BYTES;
    if (WRITE) {
      const uint32_t* s5 = reinterpret_cast<const uint32_t*>(B.streams[ZKB_STREAM_REFUND] + (size_t)vm * B.cap[ZKB_STREAM_REFUND] * ZKB_REFUND_BYTES);
      for (uint32_t i = lane; i < cnt[5] * 2; i += 32) outp[5][i] = s5[i];
    } else {
#pragma unroll
      for (int k = 0; k < ZKB_N_STREAMS; k++)
        if (lane == (uint32_t)k) A.sizes[(size_t)k * n_vms + vm] = (uint32_t)sz[k];
    }
  }
}

// exclusive prefix sums of the per-VM sizes, one block per stream; totals go to mapped host memory
__global__ void __launch_bounds__(1024) zkb_encode_scan_kernel(const DevBatch B, const EncArgs A) {
  __shared__ uint64_t s_sum[1024];
  const uint32_t k = blockIdx.x, n = B.n_vms, t = threadIdx.x;
  const uint32_t per = (n + 1023) / 1024;
  const uint32_t lo = min(n, t * per), hi = min(n, lo + per);
  const uint32_t* sizes = A.sizes + (size_t)k * n;
  uint64_t local = 0;
  for (uint32_t i = lo; i < hi; i++) local += sizes[i];
  s_sum[t] = local;
  __syncthreads();
  for (uint32_t o = 1; o < 1024; o <<= 1) {  // Hillis-Steele inclusive scan
    uint64_t v = t >= o ? s_sum[t - o] : 0;
    __syncthreads();
    s_sum[t] += v;
    __syncthreads();
  }
  uint64_t run = s_sum[t] - local;
  uint64_t* offs = A.offsets + (size_t)k * (n + 1);
  for (uint32_t i = lo; i < hi; i++) {
    offs[i] = run;
    run += sizes[i];
  }
  if (t == 1023) {
    offs[n] = s_sum[1023];
    A.totals[k] = s_sum[1023];
  }
}

__global__ void zkb_encode_header_kernel(ZkbEncodedHeader h, uint8_t* blob) {
  if (threadIdx.x == 0) *reinterpret_cast<ZkbEncodedHeader*>(blob) = h;
}

}  // namespace zkb
