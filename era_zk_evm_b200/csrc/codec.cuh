// Transport encoder of the witness streams (include/zkb_codec.h, format v2): XOR-with-prediction + presence bitmap, one warp
// per VM; cycle rows and memory queries are coded jointly (encode_joint), the other streams record by record.
//
// The canonical streams stay in HBM (they are what the device-side consumers read); this kernel is the last step before
// the PCIe link.  ONE encoding pass: every VM's payloads go to a staging area at worst-case offsets (known from the
// record counts alone: record index x the longest encoding of a record), the pass also records their true sizes, a scan
// turns those into the blob's byte offsets, and a copy kernel compacts staging -> blob and writes the blob's tables.  The
// walk over a VM is a latency-bound dependent chain (~2 600 cycles per cycle row on 16 resident warps per SM), so it
// runs once, not twice (round 2's first encoder counted in one pass and wrote in a second: 14.4 ms per 14 208-VM
// sub-batch; this: see profiles/).  The blob layout is deterministic (bit-identical to the scalar encoder in
// zkb_codec.h, which the tests check) and needs no atomics.
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
  uint8_t* blob;            // the device blob (compaction target)
  uint8_t* stage;           // staging area of the encoding pass
  uint64_t stage_base[ZKB_N_STREAMS];   // start of every stream's staging region
  const uint64_t* canon;    // [6][n_vms + 1] canonical (packed) byte offsets of every VM: its first record's index x record size
  uint64_t counts_offset, offsets_offset;
  uint64_t payload_offset[ZKB_N_STREAMS];
  uint32_t kinds_mask;      // streams outside the mask are left out of the blob (their counts are still reported)
};

// longest encoding of one record of each stream (mask words + every word present) and the canonical record size
__device__ __forceinline__ uint32_t enc_worst_bytes(int k) {
  return k == ZKB_STREAM_ROWS ? 8u + ZKB_ROW_TX_WORDS * 4u : k == ZKB_STREAM_MEM ? 52u : k == ZKB_STREAM_LOG ? 132u : k == ZKB_STREAM_DECOMMIT ? 52u
         : k == ZKB_STREAM_FRAME ? 132u : 8u;
}
__device__ __forceinline__ uint32_t* stage_slot(const EncArgs& A, int k, uint32_t vm, uint32_t n_vms) {
  const uint64_t first_record = A.canon[(size_t)k * (n_vms + 1) + vm] / rec_bytes(k);
  return reinterpret_cast<uint32_t*>(A.stage + A.stage_base[k] + first_record * enc_worst_bytes(k));
}

__device__ __forceinline__ uint32_t lanemask_lt() {
  uint32_t m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

__device__ __constant__ uint32_t c_mem_flag_codes[16] = {0x00000004u, 0x00000001u, 0x00000101u, 0x00000003u, 0x00000000u, 0x00000100u, 0x01000003u, 0x02000101u,
                                                         0x00010000u, 0x00010100u, 0x00000002u, 0x00000102u, 0x01000001u, 0x02000102u, 0x00010003u, 0xFFFFFFFFu};  // == ZKB_MEM_FLAG_CODES_INIT

// ---- format v2: cycle rows and memory queries of one VM, coded jointly by one warp (zkb_codec.h JointCoder is the scalar
// statement of the same walk; the blobs are bit-identical) -------------------------------------------------------------
// Lane l owns words l and l + 32 of the current row (two coalesced 128-byte loads per row) and word l of the current
// memory query (lanes 0..11); lanes 0..7 keep the current code word, lane s the FIFO pointer of cache set s; the
// 32-set x 4-way code-word cache of the VM sits in shared memory (4 KB per warp).
struct JointWarp {
  uint32_t prev_lo, prev_hi, pm, cw, rr, prevprev50;
  uint32_t* cache;            // [32 sets][4 ways][8] of this warp
  const uint32_t* mems;       // the VM's memory queries
  uint32_t n_mem, mi;
  uint32_t m_next;            // prefetched word `lane` of query mi
  uint32_t* out_rows;
  uint32_t* out_mem;
  uint64_t rwords, mwords;
};

__device__ __forceinline__ uint32_t first_min8(const uint32_t* c) {
  uint32_t best = c[0], sel = 0;
#pragma unroll
  for (uint32_t v = 1; v < 8; v++)
    if (c[v] < best) {
      best = c[v];
      sel = v;
    }
  return sel;
}

// one memory query (JointCoder::mem_record).  w_lo / w_hi: the cycle's row; post: w3 (the raw opcode's immediates) may be used
template <bool WRITE>
__device__ __forceinline__ void joint_mem(JointWarp& J, uint32_t lane, uint32_t lt, uint32_t w_lo, uint32_t w_hi, bool post, bool fe, uint32_t j) {
  const uint32_t w = J.m_next;
  J.mi++;
  if (J.mi < J.n_mem) J.m_next = lane < 12 ? __ldcs(J.mems + (size_t)J.mi * 12 + lane) : 0u;   // next query in flight
  const uint32_t flags = __shfl_sync(0xffffffffu, w, 3), index = __shfl_sync(0xffffffffu, w, 2);
  const uint32_t type = flags & 0xFFu, rw = (flags >> 8) & 1u;
  const uint32_t fm = __ballot_sync(0xffffffffu, lane < 15 && c_mem_flag_codes[lane & 15u] == flags);
  const uint32_t fcode = fm ? (uint32_t)__ffs(fm) - 1u : 15u;
  const uint32_t pm3 = __shfl_sync(0xffffffffu, J.pm, 3), pm2 = __shfl_sync(0xffffffffu, J.pm, 2);
  const uint32_t row_ts = __shfl_sync(0xffffffffu, w_lo, 1), row_w3 = __shfl_sync(0xffffffffu, w_lo, 3), pc_before = __shfl_sync(0xffffffffu, w_lo, 5) & 0xFFFFu;
  const uint32_t row_w8 = __shfl_sync(0xffffffffu, w_lo, 8), row_w9 = __shfl_sync(0xffffffffu, w_lo, 9), row_w10 = __shfl_sync(0xffffffffu, w_lo, 10);
  const uint32_t prev50 = __shfl_sync(0xffffffffu, J.prev_hi, 18), prev51 = __shfl_sync(0xffffffffu, J.prev_hi, 19);
  const uint32_t p0 = row_ts + (rw ? 3u : 0u);
  const uint32_t p1 = type == 4u ? prev50 : type <= 2u ? prev51 + 1u + type : row_w9;
  uint32_t p2;
  if (j > 0 && pm3 == flags) p2 = pm2 + 1u;
  else if (type == 4u) p2 = ((j == 0 && fe) || !post) ? pc_before >> 2 : row_w3 & 0xFFFFu;
  else if (type == 0u) p2 = post ? (rw ? row_w3 >> 16 : row_w3 & 0xFFFFu) : 0u;
  else if (type <= 2u) p2 = row_w8 >> 5;
  else p2 = (row_w8 + row_w10) >> 5;
  // value candidates, on lanes 4..11 (limb k = lane - 4)
  const bool vl = lane >= 4 && lane < 12;
  const uint32_t k = (lane - 4u) & 7u, set = index % ZKB_CW_SETS;
  uint32_t cand[8];
  cand[0] = 0u;
  cand[1] = __shfl_sync(0xffffffffu, w_lo, (lane + 4u) & 31u);
  cand[2] = __shfl_sync(0xffffffffu, w_lo, (lane + 12u) & 31u);
  cand[3] = __shfl_sync(0xffffffffu, w_lo, (lane + 20u) & 31u);
  const uint32_t d1 = __shfl_sync(0xffffffffu, w_hi, (lane - 4u) & 31u);
  const uint32_t* cset = J.cache + set * (ZKB_CW_WAYS * 8u);
#pragma unroll
  for (uint32_t v = 0; v < 4; v++) cand[4 + v] = type == 4u ? cset[v * 8u + k] : (v == 0 ? d1 : 0u);
  uint32_t cnt[8];
#pragma unroll
  for (uint32_t v = 0; v < 8; v++) cnt[v] = __popc(__ballot_sync(0xffffffffu, vl && w != cand[v]));
  const uint32_t vsel = first_min8(cnt);
  uint32_t pv = cand[0];
#pragma unroll
  for (uint32_t v = 1; v < 8; v++) pv = vsel == v ? cand[v] : pv;
  const uint32_t pred = lane == 0 ? p0 : lane == 1 ? p1 : lane == 2 ? p2 : lane == 3 ? J.pm : pv;
  uint32_t x = lane < 12 ? (w ^ pred) : 0u;
  if (lane == 3 && fcode < 15u) x = 0u;
  const uint32_t pres = __ballot_sync(0xffffffffu, x != 0);
  if (WRITE) {
    uint32_t* o = J.out_mem + J.mwords;
    if (lane == 0) o[0] = pres | vsel << 12 | fcode << 16;
    if (x) o[1 + __popc(pres & lt)] = x;
  }
  J.mwords += 1u + __popc(pres);
  // state: code-word cache (FIFO per set), the current code word, the previous query
  if (type == 4u) {
    const bool hit = cnt[4] == 0 || cnt[5] == 0 || cnt[6] == 0 || cnt[7] == 0;
    if (!hit) {
      const uint32_t way = __shfl_sync(0xffffffffu, J.rr, set);
      __syncwarp();
      if (vl) J.cache[set * (ZKB_CW_WAYS * 8u) + way * 8u + k] = w;
      if (lane == set) J.rr = (J.rr + 1u) % ZKB_CW_WAYS;
      __syncwarp();
    }
    const uint32_t v = __shfl_sync(0xffffffffu, w, (lane + 4u) & 31u);
    if (j == 0 && rw == 0 && index == pc_before >> 2 && lane < 8) J.cw = v;
  }
  J.pm = w;
}

template <bool WRITE>
__device__ __forceinline__ void encode_joint(JointWarp& J, const uint32_t* __restrict__ rows, uint32_t n_rows, uint32_t lane) {
  const uint32_t lt = lanemask_lt();
  J.prev_lo = J.prev_hi = J.pm = J.cw = J.rr = J.prevprev50 = 0u;
  J.mi = 0;
  J.rwords = J.mwords = 0;
  __syncwarp();
  for (uint32_t i = lane; i < ZKB_CW_SETS * ZKB_CW_WAYS * 8u; i += 32) J.cache[i] = 0u;
  __syncwarp();
  J.m_next = (J.n_mem && lane < 12) ? __ldcs(J.mems + lane) : 0u;
  uint32_t nx_lo = 0, nx_hi = 0;
  if (n_rows) {
    nx_lo = __ldcs(rows + lane);
    nx_hi = __ldcs(rows + 32 + lane);
  }
  for (uint32_t r = 0; r < n_rows; r++) {
    const uint32_t w_lo = nx_lo, w_hi = nx_hi;
    if (r + 1 < n_rows) {  // next row's loads in flight while this one is encoded
      nx_lo = __ldcs(rows + (size_t)(r + 1) * 64 + lane);
      nx_hi = __ldcs(rows + (size_t)(r + 1) * 64 + 32 + lane);
    }
    const uint32_t pc_before = __shfl_sync(0xffffffffu, w_lo, 5) & 0xFFFFu;
    const uint32_t prev48 = __shfl_sync(0xffffffffu, J.prev_hi, 16), prev50 = __shfl_sync(0xffffffffu, J.prev_hi, 18);
    const uint32_t w43 = __shfl_sync(0xffffffffu, w_hi, 11);
    // dst0 predictor (lanes 24..31 hold dst0; src0 sits 16 lanes, src1 8 lanes below)
    const uint32_t s0 = __shfl_sync(0xffffffffu, w_lo, (lane - 16u) & 31u), s1 = __shfl_sync(0xffffffffu, w_lo, (lane - 8u) & 31u);
    const bool dl = lane >= 24;
    const uint32_t sum = s0 + s1, dif = s0 - s1;
    const uint32_t Ga = __ballot_sync(0xffffffffu, sum < s0) >> 24, Pa = __ballot_sync(0xffffffffu, sum == 0xFFFFFFFFu) >> 24;
    const uint32_t Gs = __ballot_sync(0xffffffffu, s0 < s1) >> 24, Ps = __ballot_sync(0xffffffffu, s0 == s1) >> 24;
    const uint32_t Ka = carry_chain(Ga, Pa), Ks = carry_chain(Gs, Ps);
    const uint32_t addv = sum + ((Ka >> (lane & 7u)) & 1u), subv = dif - ((Ks >> (lane & 7u)) & 1u);
    const uint32_t c0 = __popc(__ballot_sync(0xffffffffu, dl && w_lo != 0u)), c1 = __popc(__ballot_sync(0xffffffffu, dl && w_lo != s0));
    const uint32_t c2 = __popc(__ballot_sync(0xffffffffu, dl && w_lo != addv)), c3 = __popc(__ballot_sync(0xffffffffu, dl && w_lo != subv));
    uint32_t dsel = 0, best = c0;
    if (c1 < best) { best = c1; dsel = 1; }
    if (c2 < best) { best = c2; dsel = 2; }
    if (c3 < best) { best = c3; dsel = 3; }
    const uint32_t dpred = dsel == 0 ? 0u : dsel == 1 ? s0 : dsel == 2 ? addv : subv;
    const uint32_t ncode = w43 < 7u ? w43 : 7u;
    // the cycle's memory queries: the first one ahead of the opcode when an instruction fetch is expected
    const bool fe = r == 0 || (pc_before >> 2) != (prev48 >> 16) || prev50 != J.prevprev50;
    const uint32_t nm = min(w43 & 0xFFFFu, J.n_mem - J.mi);
    uint32_t j = 0;
    if (fe && nm >= 1) {
      joint_mem<WRITE>(J, lane, lt, w_lo, w_hi, false, fe, 0);
      j = 1;
    }
    const uint32_t sub = pc_before & 3u;
    const uint32_t cw_a = __shfl_sync(0xffffffffu, J.cw, 6u - 2u * sub), cw_b = __shfl_sync(0xffffffffu, J.cw, 7u - 2u * sub);
    const uint32_t vidx = __shfl_sync(0xffffffffu, w_lo, 2) & ((1u << ZK_VARIANT_BITS) - 1u);
    uint32_t pred_lo = 0;
    if (lane == 0) pred_lo = J.prev_lo + 1u;
    else if (lane == 1) pred_lo = J.prev_lo + ZK_TIME_DELTA_PER_CYCLE;
    else if (lane == 2) pred_lo = cw_a;
    else if (lane == 3) pred_lo = cw_b;
    else if (lane == 4) pred_lo = vidx | 1u << 16;
    else if (lane == 5) {
      const uint32_t p = J.prev_lo >> 16;
      pred_lo = p | ((p + 1u) & 0xFFFFu) << 16;
    } else if (lane == 6) pred_lo = J.prev_lo;
    else if (lane == 7) pred_lo = J.prev_lo - ZK_OPCODE_PRICES[vidx];
    else if (dl) pred_lo = dpred;
    uint32_t pred_hi = (lane < 8 || lane == 11) ? 0u : J.prev_hi;  // words 32..39 dst1, 43 per-cycle counts
    if (lane == 16) pred_hi = (J.prev_hi & 0xFFFFu) | (pc_before >> 2) << 16;
    const uint32_t x_lo = w_lo ^ pred_lo;
    uint32_t x_hi = lane < (ZKB_ROW_TX_WORDS - 32) ? (w_hi ^ pred_hi) : 0u;   // words 55..63 are not transmitted
    if (lane == 11 && ncode < 7u) x_hi = 0u;
    const uint32_t m_lo = __ballot_sync(0xffffffffu, x_lo != 0), p_hi = __ballot_sync(0xffffffffu, x_hi != 0);
    const uint32_t n_lo = __popc(m_lo), n_hi = __popc(p_hi);
    if (WRITE) {
      uint32_t* o = J.out_rows + J.rwords;
      if (lane == 0) o[0] = m_lo;
      if (lane == 1) o[1] = p_hi | dsel << 23 | ncode << 25;
      if (x_lo) o[2 + __popc(m_lo & lt)] = x_lo;
      if (x_hi) o[2 + n_lo + __popc(p_hi & lt)] = x_hi;
    }
    J.rwords += 2u + n_lo + n_hi;
    for (; j < nm; j++) joint_mem<WRITE>(J, lane, lt, w_lo, w_hi, true, fe, j);
    J.prevprev50 = prev50;
    J.prev_lo = w_lo;
    J.prev_hi = w_hi;
  }
  // memory queries no row announces (a cycle that stopped the VM emits its queries but no row): zero row context
  for (uint32_t j = 0; J.mi < J.n_mem; j++) joint_mem<WRITE>(J, lane, lt, 0u, 0u, true, false, j);
}

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
    pred = li < 4 ? p : 0u;
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
    const uint32_t rd = __shfl_sync(0xffffffffu, w, (lane - 8u) & 31u);   // LOG: written_value is predicted by the read value
    const uint32_t pred = (KIND == ZKB_STREAM_FRAME || lane < 8) ? prev : (KIND == ZKB_STREAM_LOG && lane >= 24) ? rd : 0u;
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

// the encoding pass: one warp per VM, payloads into the staging area, true sizes into A.sizes.  THREE CTAs per SM (80
// registers): four would own the whole register file, and the small populate / memset kernels of the NEXT sub-batch (<= 32
// registers x 128 threads) could not start underneath it -- the host thread that populates then sits in its stream syncs until
// the encoder drains and the interpreter launch it queues afterwards leaves the GPU idle for the rest of its host work.
__global__ void __launch_bounds__(256, 3) zkb_encode_kernel(const DevBatch B, const EncArgs A) {
  __shared__ uint32_t s_cache[8][ZKB_CW_SETS * ZKB_CW_WAYS * 8];   // the code-word cache of each warp's VM
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t n_vms = B.n_vms;
  for (uint32_t vm = blockIdx.x * 8 + warp; vm < n_vms; vm += gridDim.x * 8) {
    const uint32_t* x = B.hot[vm].x;
    uint32_t cnt[ZKB_N_STREAMS];
#pragma unroll
    for (int k = 0; k < ZKB_N_STREAMS; k++) cnt[k] = ((A.kinds_mask >> k) & 1u) ? x[X_COUNT0 + k] : 0u;
    uint64_t sz[ZKB_N_STREAMS];
    const uint32_t* s0 = reinterpret_cast<const uint32_t*>(B.streams[ZKB_STREAM_ROWS] + (size_t)vm * B.cap[ZKB_STREAM_ROWS] * ZKB_ROW_BYTES);
    const uint32_t* s1 = reinterpret_cast<const uint32_t*>(B.streams[ZKB_STREAM_MEM] + (size_t)vm * B.cap[ZKB_STREAM_MEM] * ZKB_MEM_BYTES);
    JointWarp J;
    J.cache = s_cache[warp];
    J.mems = s1;
    J.n_mem = cnt[1];
    J.out_rows = stage_slot(A, ZKB_STREAM_ROWS, vm, n_vms);
    J.out_mem = stage_slot(A, ZKB_STREAM_MEM, vm, n_vms);
    encode_joint<true>(J, s0, cnt[0], lane);   // (a subset blob carries ROWS and MEM together or not at all)
    sz[0] = J.rwords * 4;
    sz[1] = J.mwords * 4;
    const uint32_t* s2 = reinterpret_cast<const uint32_t*>(B.streams[ZKB_STREAM_LOG] + (size_t)vm * B.cap[ZKB_STREAM_LOG] * ZKB_LOG_BYTES);
    sz[2] = encode_rec32<true, ZKB_STREAM_LOG>(s2, cnt[2], stage_slot(A, ZKB_STREAM_LOG, vm, n_vms), lane);
    const uint32_t* s3 = reinterpret_cast<const uint32_t*>(B.streams[ZKB_STREAM_DECOMMIT] + (size_t)vm * B.cap[ZKB_STREAM_DECOMMIT] * ZKB_DECOMMIT_BYTES);
    sz[3] = encode_rec12<true, ZKB_STREAM_DECOMMIT>(s3, cnt[3], stage_slot(A, ZKB_STREAM_DECOMMIT, vm, n_vms), lane);
    const uint32_t* s4 = reinterpret_cast<const uint32_t*>(B.streams[ZKB_STREAM_FRAME] + (size_t)vm * B.cap[ZKB_STREAM_FRAME] * ZKB_FRAME_BYTES);
    sz[4] = encode_rec32<true, ZKB_STREAM_FRAME>(s4, cnt[4], stage_slot(A, ZKB_STREAM_FRAME, vm, n_vms), lane);
    sz[5] = (uint64_t)cnt[5] * ZKB_REFUND_BYTES;   // REFUND: raw 8-byte records, copied by the compaction straight from the stream
#pragma unroll
    for (int k = 0; k < ZKB_N_STREAMS; k++)
      if (lane == (uint32_t)k) A.sizes[(size_t)k * n_vms + vm] = (uint32_t)sz[k];
  }
}

// compaction: staging -> blob at the scanned offsets, plus the blob's per-VM tables.  One CTA per VM (grid-stride).
__global__ void __launch_bounds__(128) zkb_encode_compact_kernel(const DevBatch B, const EncArgs A) {
  const uint32_t n_vms = B.n_vms, t = threadIdx.x;
  for (uint32_t vm = blockIdx.x; vm < n_vms; vm += gridDim.x) {
    const uint32_t* x = B.hot[vm].x;
    // the blob's tables: per-VM summary (same 8 words the host mirror holds) and this VM's offsets
    uint32_t* counts = reinterpret_cast<uint32_t*>(A.blob + A.counts_offset) + (size_t)vm * 8;
    const uint32_t status = x[X_STATUS];
    const uint32_t st_out = (status == ZKB_VM_RUNNING && B.hot[vm].live[L_DEPTH - 40] == 0 && x[X_CYCLE] > 0) ? (uint32_t)ZKB_VM_ENDED : status;
    if (t < 8) counts[t] = t < 6 ? x[X_COUNT0 + t] : t == 6 ? st_out : x[X_CYCLE];
    uint64_t* offs = reinterpret_cast<uint64_t*>(A.blob + A.offsets_offset);
    if (t < ZKB_N_STREAMS) {
      offs[(size_t)t * (n_vms + 1) + vm] = A.offsets[(size_t)t * (n_vms + 1) + vm];
      if (vm == n_vms - 1) offs[(size_t)t * (n_vms + 1) + n_vms] = A.offsets[(size_t)t * (n_vms + 1) + n_vms];
    }
#pragma unroll
    for (int k = 0; k < ZKB_N_STREAMS; k++) {
      const uint32_t words = A.sizes[(size_t)k * n_vms + vm] / 4u;
      const uint32_t* src = k == ZKB_STREAM_REFUND
                                ? reinterpret_cast<const uint32_t*>(B.streams[ZKB_STREAM_REFUND] + (size_t)vm * B.cap[ZKB_STREAM_REFUND] * ZKB_REFUND_BYTES)
                                : stage_slot(A, k, vm, n_vms);
      uint32_t* dst = reinterpret_cast<uint32_t*>(A.blob + A.payload_offset[k] + A.offsets[(size_t)k * (n_vms + 1) + vm]);
      for (uint32_t i = t; i < words; i += 128) dst[i] = src[i];
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

// header + the alignment padding after every payload (each payload starts 16-byte aligned; the bytes in between are part
// of the blob and must be deterministic: the blob is compared byte for byte with the scalar encoder's)
__global__ void zkb_encode_header_kernel(ZkbEncodedHeader h, uint8_t* blob) {
  const uint32_t t = threadIdx.x;
  if (t == 0) *reinterpret_cast<ZkbEncodedHeader*>(blob) = h;
  if (t < ZKB_N_STREAMS) {
    const uint64_t lo = h.payload_offset[t] + h.payload_bytes[t], hi = t + 1 < ZKB_N_STREAMS ? h.payload_offset[t + 1] : h.total_bytes;
    for (uint64_t i = lo; i < hi; i++) blob[i] = 0;
  }
}

}  // namespace zkb
