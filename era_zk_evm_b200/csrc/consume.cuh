// K12 / K13: the device-side CONSUMER of the witness streams -- what the downstream-shaped
// VmWitnessTracer adapter produces for circuit synthesis, computed where the streams already are:
//   * per-circuit-batch snapshots: the full VmLocalState handed to start_new_execution_cycle every K-th cycle (and after
//     the last one), rebuilt from the cycle rows + frame records + the instruction-fetch memory queries exactly as the host
//     replay does (include/zkb_host.hpp replay_records; reference contract: src/witness_trace/mod.rs:11-72, state layout
//     src/vm_state/mod.rs:54-73) -- registers follow the dst0 / dst1 write-backs and the register ABI of far_call / ret
//     (far_call.rs:573-610, ret.rs:213-236);
//   * queue-state commitments: a running hash over the memory, log and decommitment queues, recorded at every snapshot
//     boundary (so that the circuits of batch j start from the queue tail of batch j - 1) and finalised per VM.
//     The downstream harness uses a Poseidon2 sponge over Goldilocks whose parameters are not part of /root/reference; the
//     stand-in is the sha256 compression chain  state' = compress(state, record zero-padded to 64-byte blocks)  from the
//     sha256 IV, finalised with the standard padding -- i.e. final = sha256(padded records), which hashlib pins.
// With these a host loop needs only snapshots + commitments + the (encoded) query logs across PCIe; the full-stream
// download stays available.
//   K12 zkb_consume_state_kernel   one warp per VM, walks the rows once (HBM-read bound: 256 B per cycle)
//   K13 zkb_consume_hash_kernel    one THREAD per (queue, VM): sequential chain per queue, parallel across VMs; a warp = one queue kind
#pragma once
#include <stdint.h>

#include "vm.cuh"

namespace zkb {

#define ZKB_SNAP_WORDS 198u   // sizeof(ZkbSnapshot) / 4
static_assert(sizeof(ZkbSnapshot) == ZKB_SNAP_WORDS * 4 && sizeof(ZkbLocalState) == 680, "ZkbSnapshot layout");
enum {  // word offsets inside ZkbSnapshot (ZkbLocalState first, include/zkb.h)
  SW_PCW = 0, SW_PCMP = 8, SW_REGS = 9, SW_BITS = 129, SW_TS = 130, SW_CYCLE = 131, SW_SPENT = 132, SW_PAGECTR = 133, SW_ABS = 134,
  SW_EPP = 135, SW_TX_PSP = 136, SW_CTX = 137, SW_DEPTH = 141, SW_FRAME = 142,
  SW_F_BASE = 157, SW_F_CODE_PAGE = 158, SW_F_SP_PC = 159, SW_F_EH = 160, SW_F_ERGS = 161, SW_F_SHARDS = 162, SW_F_LOCAL = 163, SW_F_CTX = 164,
  SW_F_HEAP = 168, SW_F_AUX = 169, SW_SNAP_CYCLE = 170, SW_N_MEM = 171, SW_N_LOG = 172, SW_N_DEC = 173, SW_QUEUES = 174
};

struct ConsumeOut {
  uint32_t* snaps;       // [vm][max_snaps][ZKB_SNAP_WORDS]
  uint32_t* n_snaps;     // [vm]
  uint32_t* finals;      // [vm][3][8] final queue digests (big-endian words as sha256 prints them)
  uint32_t max_snaps, period;
  const VmHot* hot_init; // the VMs' state before their first cycle
  uint32_t* fstack;      // [vm][max_depth + 1] frame-record index of every live frame
};

// per-warp tracked state
struct ConsumeSmem {
  uint32_t regs[15][8];
  uint32_t pcw[8];
  uint32_t ptr_mask;     // bit i = registers[i].is_pointer
  uint32_t prev_code_page;
  uint32_t pad[6];
};

__global__ void __launch_bounds__(256) zkb_consume_state_kernel(const DevBatch B, const ConsumeOut O) {
  __shared__ ConsumeSmem sm[8];
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  ConsumeSmem& S = sm[warp];
  for (uint32_t vm = blockIdx.x * 8 + warp; vm < B.n_vms; vm += gridDim.x * 8) {
    const VmHot* h0 = O.hot_init + vm;
    const uint32_t n_rows = B.hot[vm].x[X_COUNT0 + ZKB_STREAM_ROWS], n_frames = B.hot[vm].x[X_COUNT0 + ZKB_STREAM_FRAME];
    const uint32_t* rows = reinterpret_cast<const uint32_t*>(B.streams[ZKB_STREAM_ROWS] + (size_t)vm * B.cap[ZKB_STREAM_ROWS] * ZKB_ROW_BYTES);
    const uint32_t* mems = reinterpret_cast<const uint32_t*>(B.streams[ZKB_STREAM_MEM] + (size_t)vm * B.cap[ZKB_STREAM_MEM] * ZKB_MEM_BYTES);
    const uint32_t* frames = reinterpret_cast<const uint32_t*>(B.streams[ZKB_STREAM_FRAME] + (size_t)vm * B.cap[ZKB_STREAM_FRAME] * ZKB_FRAME_BYTES);
    uint32_t* fstack = O.fstack + (size_t)vm * (B.max_depth + 1);
    uint32_t* out = O.snaps + (size_t)vm * O.max_snaps * ZKB_SNAP_WORDS;
    // ---- state before the first recorded cycle (what the host populated) ----
    for (uint32_t i = lane; i < 120; i += 32) S.regs[i >> 3][i & 7u] = h0->regs[1 + (i >> 3)][i & 7u];
    if (lane < 8) S.pcw[lane] = h0->prev_word[lane];
    if (lane == 0) {
      S.ptr_mask = h0->x[X_PTRMASK] >> 1;
      S.prev_code_page = h0->x[X_PREV_CODE_PAGE];
    }
    // scalars, warp-uniform registers: everything end_execution_cycle of the previous cycle left behind
    uint32_t flags_pending = (h0->x[X_FLAGS] & 7u) | (h0->x[X_PENDING] ? 0x100u : 0u);
    uint32_t spent = h0->live[L_SPENT_PUBDATA - 40], pagectr = h0->live[L_PAGE_COUNTER - 40], epp = h0->live[L_EPP - 40], tx_psp = h0->live[L_TX_PSP - 40];
    uint32_t ctx = lane < 4 ? h0->live[L_CTX - 40 + lane] : 0u;   // lanes 0..3
    uint32_t depth = h0->live[L_DEPTH - 40];
    uint32_t pc = h0->F[F_SP_PC] >> 16, sp = h0->F[F_SP_PC] & 0xFFFFu, ergs = h0->F[F_ERGS], heap_bound = h0->F[F_HEAP_BOUND], aux_bound = h0->F[F_AUX_BOUND];
    uint32_t eh = h0->F[F_EH_SHARDS] & 0xFFFFu;
    // frame stack: the bootloader push (helpers.rs:289-316) is frame record 0 when present
    uint32_t in_rows = 0;
    for (uint32_t base = 0; base < n_rows; base += 32) {
      const uint32_t w = base + lane < n_rows ? rows[(size_t)(base + lane) * 64 + 43] : 0u;
      in_rows += (w >> 26) & 3u;
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) in_rows += __shfl_xor_sync(0xffffffffu, in_rows, o);
    uint32_t ifr = 0, fsp = 0;   // next frame record; live frames above the root
    if (n_frames == in_rows + 1) {
      if (lane == 0) fstack[0] = 0u;
      fsp = 1;
      ifr = 1;
    }
    __syncwarp();
    uint32_t im = 0, il = 0, idc = 0, n_snap = 0;

    // writes the snapshot of the state handed to start_new_execution_cycle(cycle, timestamp)
    auto emit = [&](uint32_t cycle, uint32_t timestamp) {
      if (n_snap >= O.max_snaps) return;
      uint32_t* s = out + (size_t)n_snap * ZKB_SNAP_WORDS;
      __syncwarp();
      for (uint32_t i = lane; i < 120; i += 32) s[SW_REGS + i] = S.regs[i >> 3][i & 7u];
      if (lane < 8) s[SW_PCW + lane] = S.pcw[lane];
      if (lane < 4) {
        s[SW_CTX + lane] = ctx;
      }
      // current frame: static part from the frame record that opened it (the root frame: all zero), dynamic part tracked
      const bool root = fsp == 0;
      const uint32_t* fr = frames + (size_t)(root ? 0u : fstack[fsp - 1]) * 32;
      if (lane < 15) s[SW_FRAME + lane] = root ? 0u : fr[2 + lane];   // this / sender / code addresses
      if (lane >= 16 && lane < 20) s[SW_F_CTX + lane - 16] = root ? 0u : fr[23 + lane - 16];
      if (lane == 20) {
        const uint32_t w20 = root ? 0u : fr[20], w22 = root ? 0u : fr[22];
        s[SW_PCMP] = S.prev_code_page;
        s[SW_BITS] = (S.ptr_mask & 0x7FFFu) | (flags_pending & 7u) << 16 | ((flags_pending >> 8) & 1u) << 24;
        s[SW_TS] = timestamp;
        s[SW_CYCLE] = cycle;
        s[SW_SPENT] = spent;
        s[SW_PAGECTR] = pagectr;
        s[SW_ABS] = 0u;
        s[SW_EPP] = epp;
        s[SW_TX_PSP] = tx_psp;
        s[SW_DEPTH] = depth;
        s[SW_F_BASE] = root ? 0u : fr[17];
        s[SW_F_CODE_PAGE] = root ? 0u : fr[18];
        s[SW_F_SP_PC] = sp | pc << 16;
        s[SW_F_EH] = eh;
        s[SW_F_ERGS] = ergs;
        s[SW_F_SHARDS] = ((w20 >> 16) & 0xFFu) | (w20 >> 24) << 8 | (w22 & 0xFFu) << 16 | ((w22 >> 8) & 1u) << 24;
        s[SW_F_LOCAL] = (w22 >> 16) & 1u;
        s[SW_F_HEAP] = heap_bound;
        s[SW_F_AUX] = aux_bound;
        s[SW_SNAP_CYCLE] = cycle;
        s[SW_N_MEM] = im;
        s[SW_N_LOG] = il;
        s[SW_N_DEC] = idc;
      }
      n_snap++;
      __syncwarp();
    };

    uint32_t last_cycle = 0, last_ts = 0;
    bool any = false;
    for (uint32_t r = 0; r < n_rows; r++) {
      const uint32_t w_lo = __ldcs(rows + (size_t)r * 64 + lane), w_hi = __ldcs(rows + (size_t)r * 64 + 32 + lane);
      const uint32_t cycle = __shfl_sync(0xffffffffu, w_lo, 0), ts = __shfl_sync(0xffffffffu, w_lo, 1);
      if (cycle != 0 && cycle % O.period == 0) emit(cycle, ts);
      const uint32_t raw_lo = __shfl_sync(0xffffffffu, w_lo, 2);
      const uint32_t w4 = __shfl_sync(0xffffffffu, w_lo, 4), w5 = __shfl_sync(0xffffffffu, w_lo, 5), w6 = __shfl_sync(0xffffffffu, w_lo, 6);
      const uint32_t w43 = __shfl_sync(0xffffffffu, w_hi, 43 - 32);
      const uint32_t n_mem = w43 & 0xFFFFu, n_log = (w43 >> 16) & 0xFFu, n_dec = (w43 >> 24) & 3u, n_frame = (w43 >> 26) & 3u;
      const uint32_t bits = w6 >> 24;
      const uint32_t entry = ZK_OPCODE_TABLE[w4 & 0xFFFFu];
      const uint32_t family = entry & 15u, dst_mode = (entry >> ZK_E_DST_SHIFT) & 3u;
      const bool masked = (w4 >> 24) != 0 || ((w4 >> 16) & 0xFFu) == 0;
      const uint32_t ops = masked ? 0u : raw_lo;
      const uint32_t dst0_reg = (ops >> 24) & 15u, dst1_reg = ops >> 28;
      // previous_code_word follows the instruction fetch (cycle.rs:59-100)
      const uint32_t pc_before = w5 & 0xFFFFu, super_pc = pc_before >> 2;
      const uint32_t cur_code_page = fsp == 0 ? 0u : frames[(size_t)fstack[fsp - 1] * 32 + 18];
      const bool was_pending = (flags_pending >> 8) & 1u;
      if (!was_pending && !(bits & ZKB_ROWBIT_SKIP) && (cur_code_page != S.prev_code_page || (tx_psp >> 16) != super_pc)) {
        if (n_mem > 0 && lane < 8) S.pcw[lane] = mems[(size_t)im * 12 + 4 + lane];
      }
      __syncwarp();
      if (lane == 0) S.prev_code_page = cur_code_page;
      // frames of the cycle, in emission order (helpers.rs:225-264)
      for (uint32_t j = 0; j < n_frame; j++, ifr++) {
        const uint32_t head = frames[(size_t)ifr * 32];
        if ((head & 0xFFu) == ZKB_FRAMEKIND_START) {
          if (fsp <= B.max_depth && lane == 0) fstack[fsp] = ifr;
          fsp++;
        } else if (fsp > 0) {
          fsp--;
        }
        __syncwarp();
      }
      // registers (far_call.rs:573-610, ret.rs:213-236, helpers.rs:266-287)
      const bool far_call_done = family == ZK_OP_FAR_CALL && n_frame == 1;
      const bool far_ret_done = family == ZK_OP_RET && n_frame == 1 && (bits & ZKB_ROWBIT_DST0_VALID);
      const uint32_t dst0_limb = __shfl_sync(0xffffffffu, w_lo, 24 + (lane & 7u));   // lanes 0..7: limb of dst0
      const uint32_t dst1_limb = __shfl_sync(0xffffffffu, w_hi, lane & 7u);          // lanes 0..7: limb of dst1 (row words 32..39)
      if (far_call_done) {
        const bool to_system = (__shfl_sync(0xffffffffu, w_hi, 0) & 2u) != 0;
        if (lane < 8) {
          S.regs[0][lane] = dst0_limb;
          S.regs[1][lane] = dst1_limb;
          if (!to_system)
            for (int i = 2; i < 12; i++) S.regs[i][lane] = 0u;
          for (int i = 12; i < 15; i++) S.regs[i][lane] = 0u;
        }
        if (lane == 0) S.ptr_mask = 1u;   // r1 is the calldata pointer; every other marker is cleared or its register zeroed
      } else if (far_ret_done) {
        if (lane < 8) {
          S.regs[0][lane] = dst0_limb;
          for (int i = 1; i < 15; i++) S.regs[i][lane] = 0u;
        }
        if (lane == 0) S.ptr_mask = 1u;
      } else {
        if ((bits & ZKB_ROWBIT_DST0_VALID) && dst_mode == ZK_DST_REG && dst0_reg != 0) {
          if (lane < 8) S.regs[dst0_reg - 1][lane] = dst0_limb;
          if (lane == 0) S.ptr_mask = (S.ptr_mask & ~(1u << (dst0_reg - 1))) | ((bits & ZKB_ROWBIT_DST0_PTR) ? 1u << (dst0_reg - 1) : 0u);
        }
        __syncwarp();
        if ((bits & ZKB_ROWBIT_DST1_VALID) && dst1_reg != 0) {
          if (lane < 8) S.regs[dst1_reg - 1][lane] = dst1_limb;
          if (lane == 0) S.ptr_mask = (S.ptr_mask & ~(1u << (dst1_reg - 1))) | ((bits & ZKB_ROWBIT_DST1_PTR) ? 1u << (dst1_reg - 1) : 0u);
        }
      }
      __syncwarp();
      // scalars after the cycle: the row carries them
      flags_pending = ((w6 >> 16) & 7u) | ((bits & ZKB_ROWBIT_PENDING) ? 0x100u : 0u);
      spent = __shfl_sync(0xffffffffu, w_hi, 41 - 32);
      pagectr = __shfl_sync(0xffffffffu, w_hi, 42 - 32);
      epp = __shfl_sync(0xffffffffu, w_hi, 49 - 32);
      tx_psp = __shfl_sync(0xffffffffu, w_hi, 48 - 32);
      ctx = __shfl_sync(0xffffffffu, w_hi, 44 - 32 + (lane & 3u));
      depth = __shfl_sync(0xffffffffu, w_hi, 40 - 32);
      pc = w5 >> 16;
      sp = w6 & 0xFFFFu;
      ergs = __shfl_sync(0xffffffffu, w_lo, 7);
      heap_bound = __shfl_sync(0xffffffffu, w_hi, 52 - 32);
      aux_bound = __shfl_sync(0xffffffffu, w_hi, 53 - 32);
      eh = __shfl_sync(0xffffffffu, w_hi, 54 - 32) & 0xFFFFu;
      im += n_mem;
      il += n_log;
      idc += n_dec;
      last_cycle = cycle + 1;
      last_ts = (bits & ZKB_ROWBIT_SKIP) ? ts : ts + ZK_TIME_DELTA_PER_CYCLE;
      any = true;
    }
    // the state after the last recorded cycle (also when it falls on a period boundary: emitted once, here)
    if (any) emit(last_cycle, last_ts);
    if (lane == 0) O.n_snaps[vm] = n_snap;
    __syncwarp();
  }
}

// ---- sha256 compression, one thread (registers only) ----
__device__ __forceinline__ void sha256_compress_thread(uint32_t st[8], uint32_t w[16]) {
  uint32_t a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
#pragma unroll
  for (int t = 0; t < 64; t++) {
    if (t >= 16) {
      const uint32_t w15 = w[(t - 15) & 15], w2 = w[(t - 2) & 15];
      const uint32_t s0 = rotr32(w15, 7) ^ rotr32(w15, 18) ^ (w15 >> 3);
      const uint32_t s1 = rotr32(w2, 17) ^ rotr32(w2, 19) ^ (w2 >> 10);
      w[t & 15] = w[t & 15] + s0 + w[(t - 7) & 15] + s1;
    }
    const uint32_t S1 = rotr32(e, 6) ^ rotr32(e, 11) ^ rotr32(e, 25);
    const uint32_t ch = (e & f) ^ (~e & g);
    const uint32_t t1 = h + S1 + ch + c_sha256_k[t] + w[t & 15];
    const uint32_t S0 = rotr32(a, 2) ^ rotr32(a, 13) ^ rotr32(a, 22);
    const uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
    h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + S0 + mj;
  }
  st[0] += a; st[1] += b; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
}

// One thread per (VM, queue q = 0 memory / 1 log / 2 decommit): absorbs the queue's records in order (bytes as stored,
// i.e. message words are the byte-swapped little-endian record words), drops the chaining value into every snapshot
// whose boundary count is reached, finalises with the standard sha256 padding.
__global__ void __launch_bounds__(128) zkb_consume_hash_kernel(const DevBatch B, const ConsumeOut O) {
  // thread t -> (queue t / n_vms, VM t % n_vms): the lanes of a warp hash the SAME queue of 32 neighbouring VMs, whose chains
  // are about equally long (a (vm, queue) interleave left two thirds of every warp waiting for its memory-queue lanes)
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= B.n_vms * 3u) return;
  const uint32_t q = t / B.n_vms, vm = t % B.n_vms;
  const int kind = q == 0 ? ZKB_STREAM_MEM : q == 1 ? ZKB_STREAM_LOG : ZKB_STREAM_DECOMMIT;
  const uint32_t rec_words = q == 1 ? 32u : 12u;
  const uint32_t n = B.hot[vm].x[X_COUNT0 + kind];
  const uint32_t* recs = reinterpret_cast<const uint32_t*>(B.streams[kind] + (size_t)vm * B.cap[kind] * (rec_words * 4));
  uint32_t* snaps = O.snaps + (size_t)vm * O.max_snaps * ZKB_SNAP_WORDS;
  const uint32_t n_snap = O.n_snaps[vm];
  uint32_t st[8];
#pragma unroll
  for (int i = 0; i < 8; i++) st[i] = c_sha256_iv[i];
  uint32_t next = 0;
  uint32_t bound = n_snap ? snaps[SW_N_MEM + q] : 0xFFFFFFFFu;
  for (uint32_t r = 0;; r++) {
    while (next < n_snap && bound == r) {   // chaining value in front of record r
#pragma unroll
      for (int i = 0; i < 8; i++) snaps[(size_t)next * ZKB_SNAP_WORDS + SW_QUEUES + q * 8 + i] = st[i];
      next++;
      bound = next < n_snap ? snaps[(size_t)next * ZKB_SNAP_WORDS + SW_N_MEM + q] : 0xFFFFFFFFu;
    }
    if (r >= n) break;
    const uint32_t* p = recs + (size_t)r * rec_words;
    uint32_t w[16];
    for (uint32_t blk = 0; blk * 16 < rec_words; blk++) {
#pragma unroll
      for (uint32_t i = 0; i < 16; i++) {
        const uint32_t wi = blk * 16 + i;
        w[i] = wi < rec_words ? bswap32(p[wi]) : 0u;
      }
      sha256_compress_thread(st, w);
    }
  }
  // finalisation: 0x80, zeros, 64-bit bit length of the padded stream (a whole number of blocks)
  uint32_t w[16];
#pragma unroll
  for (int i = 0; i < 16; i++) w[i] = 0u;
  w[0] = 0x80000000u;
  const uint64_t bits = (uint64_t)n * (q == 1 ? 128u : 64u) * 8u;
  w[14] = (uint32_t)(bits >> 32);
  w[15] = (uint32_t)bits;
  sha256_compress_thread(st, w);
  uint32_t* fin = O.finals + ((size_t)vm * 3 + q) * 8;
#pragma unroll
  for (int i = 0; i < 8; i++) fin[i] = st[i];
}

}  // namespace zkb
