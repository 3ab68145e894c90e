// zkb.cu — kernels + the C ABI of include/zkb.h (host side) for the B200 batched EraVM witness generator.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC (see build.py).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "vm.cuh"
#include "codec.cuh"
#include "logsort.cuh"
#include "consume.cuh"
#include "alubench.cuh"

using namespace zkb;

#ifndef ZKB_WARPS_PER_CTA
#define ZKB_WARPS_PER_CTA 24
#endif
#ifndef ZKB_MIN_CTAS_PER_SM
#define ZKB_MIN_CTAS_PER_SM 1
#endif

// ---------------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------------
// K1: persistent interpreter.  Replaces the caller loop `while !vm.execution_has_ended() { vm.cycle() }`.
// One VM per octet (8 lanes), four VMs per warp, ZKB_VMS_PER_CTA per CTA (vm.cuh).
// LOCKSTEP = false: each warp pulls four consecutive VM indices from an atomic queue and runs them to completion.
// LOCKSTEP = true : each CTA pulls ZKB_VMS_PER_CTA consecutive VMs and steps them together (run_vm_group).
#define ZKB_VMS_PER_CTA (ZKB_WARPS_PER_CTA * ZK_VMS_PER_WARP)
#define ZKB_RUN_SMEM_BYTES (ZKB_VMS_PER_CTA * sizeof(VmSmem))
template <bool LOCKSTEP, bool KD>
__global__ void __launch_bounds__(ZKB_WARPS_PER_CTA * 32, ZKB_MIN_CTAS_PER_SM) zkb_run_kernel(const DevBatch B, uint32_t max_cycles) {
  extern __shared__ uint4 smem_raw[];
  __shared__ uint32_t s_base;
  __shared__ uint32_t s_kc_flags[3];
  VmSmem* smem = reinterpret_cast<VmSmem*>(smem_raw);
  if (threadIdx.x < 3) s_kc_flags[threadIdx.x] = 0u;
  if ((threadIdx.x & 7u) == 0) smem[threadIdx.x >> 3].kc[KB_KC_PENDING] = 0u;   // slots that never load a VM must not look pending
  __syncthreads();
  uint32_t lane = oct_lane();
  const uint32_t warp = threadIdx.x >> 5, oct = threadIdx.x >> 3;  // oct = VM slot within the CTA
#ifdef ZKB_PINNED_BASE
  // Experiment (off by default): the VM's shared-memory window and the lane index are used by almost every instruction;
  // left to itself ptxas re-derives them from %tid / the shared window base again and again (7 % of the executed
  // instructions).  Passing them through an empty volatile asm pins them in registers: +3.5 % on the register-only ALU
  // loop, -2.7 % on ERC-20 (two more registers of pressure in the far-call / UMA handlers -> spills).
  uint32_t s_addr = (uint32_t)__cvta_generic_to_shared(&smem[oct]);
  asm volatile("" : "+r"(s_addr), "+r"(lane));
  VmSmem& S = *reinterpret_cast<VmSmem*>(__cvta_shared_to_generic(s_addr));
#elif defined(ZKB_PINNED_LANE)
  asm volatile("" : "+r"(lane));
  VmSmem& S = smem[oct];
#elif defined(ZKB_PINNED_SADDR)
  uint32_t s_addr = (uint32_t)__cvta_generic_to_shared(&smem[oct]);
  asm volatile("" : "+r"(s_addr));
  VmSmem& S = *reinterpret_cast<VmSmem*>(__cvta_shared_to_generic(s_addr));
#else
  VmSmem& S = smem[oct];
#endif
  uint8_t* mem_all = B.witness ? B.streams[ZKB_STREAM_MEM] : nullptr;
  if (!LOCKSTEP) {
    while (true) {
      uint32_t vm_base = 0;
      if ((threadIdx.x & 31u) == 0) vm_base = atomicAdd(B.queue + (KD ? 1 : 0), (uint32_t)ZK_VMS_PER_WARP);
      vm_base = __shfl_sync(ZK_FULL, vm_base, 0);
      if (vm_base >= B.n_vms) break;
      uint32_t n = 0;
      while (run_vm_group<false, KD>(B, S, s_kc_flags, vm_base + oct_index(), lane, max_cycles, n)) {
        // deferred keccak256 of the warp's yielded VMs: lanes 0..3 take its four slots (one thread per state)
        __syncwarp();
        const uint32_t wl = threadIdx.x & 31u;
        if constexpr (KD) {
          if (wl < ZK_VMS_PER_WARP) run_deferred_keccak(smem[warp * ZK_VMS_PER_WARP + wl].kc, B.heap_mem, B.n_slabs, B.heap_words, mem_all, B.cap[ZKB_STREAM_MEM]);
        }
        __syncwarp();
      }
    }
  } else {
    while (true) {
      if (threadIdx.x == 0) s_base = atomicAdd(B.queue + (KD ? 1 : 0), B.chunk);
      __syncthreads();
      const uint32_t base = s_base;
      if (base >= B.n_vms) break;
      uint32_t n = 0;
      // slots beyond the chunk stay empty this turn (B.chunk < VMs per CTA evens out the last wave, see zkb_run)
      const uint32_t vm_idx = oct < B.chunk ? base + oct : 0xFFFFFFFFu;
      while (run_vm_group<true, KD>(B, S, s_kc_flags, vm_idx, lane, max_cycles, n)) {
        // deferred keccak256 of every yielded VM of the CTA: thread t < VMs per CTA takes slot t (one thread per state)
        __syncthreads();
        if constexpr (KD) {
          if (threadIdx.x < ZKB_VMS_PER_CTA) run_deferred_keccak(smem[threadIdx.x].kc, B.heap_mem, B.n_slabs, B.heap_words, mem_all, B.cap[ZKB_STREAM_MEM]);
        }
        if (threadIdx.x < 3) s_kc_flags[threadIdx.x] = 0u;
        __syncthreads();
      }
      __syncthreads();
    }
  }
  (void)warp;
}

// K6a: per-VM cold-state initialisation (level table, page indirections)
__global__ void zkb_init_kernel(const DevBatch B) {
  uint32_t vm = blockIdx.x * blockDim.x + threadIdx.x;
  if (vm >= B.n_vms) return;
  uint32_t* lvl = B.lvl + (size_t)vm * (B.max_far_depth + 1) * 4;
  for (uint32_t l = 0; l <= B.max_far_depth; l++) {
    lvl[l * 4 + 0] = ZKB_NO_SLAB;
    lvl[l * 4 + 1] = ZKB_NO_SLAB;
    lvl[l * 4 + 2] = 0;
    lvl[l * 4 + 3] = 0;
  }
  uint32_t* pt = B.pt + (size_t)vm * ZKB_PT_ENTRIES * 2;
  for (uint32_t e = 0; e < ZKB_PT_ENTRIES; e++) {
    pt[e * 2] = ZKB_PT_FREE;
    pt[e * 2 + 1] = 0;
  }
}

struct DevStorageInit {
  uint32_t shard;
  uint32_t addr[5];
  uint32_t key[8];
  uint32_t value[8];
};

// K6b: InMemoryStorage::populate (storage.rs:26-32): one octet per VM inserts n entries through the same
// open-addressed insert the interpreter uses (no journal).
__global__ void zkb_populate_storage_kernel(const DevBatch B, uint32_t vm_lo, uint32_t vm_hi, const DevStorageInit* entries, uint32_t n,
                                            uint32_t per_vm, uint32_t* fail_flag) {
  __shared__ VmSmem smem[16];  // not touched by storage_access; the Vm object only wants a reference
  const uint32_t lane = oct_lane(), oct = threadIdx.x >> 3;
  uint32_t vm = vm_lo + blockIdx.x * 16 + oct;
  if (vm >= vm_hi) return;  // octet-uniform: whole octets leave together
  Vm<false> v(B, smem[oct], vm, lane);
  v.status = ZKB_VM_RUNNING;
  v.journal_len() = 0;
  const DevStorageInit* e = per_vm ? entries + (size_t)(vm - vm_lo) * n : entries;
  for (uint32_t i = 0; i < n; i++) {
    uint32_t aw = lane < 5 ? e[i].addr[lane] : 0u;
    u256l key = e[i].key[lane];
    u256l val = e[i].value[lane];
    v.storage_access(e[i].shard, aw, key, Vm<false>::ST_POPULATE, val);
  }
  if (v.status != ZKB_VM_RUNNING && lane == 0) *fail_flag = v.status;
}

// K6c: SimpleMemory::populate_heap (memory.rs:287-291) into slab 0 (the bootloader frame's heap).
// bytes: big-endian memory image; word w limb l = BE bytes [32w + 28 - 4l, +4)
__global__ void zkb_populate_heap_kernel(const DevBatch B, uint32_t vm_lo, uint32_t vm_hi, const uint8_t* bytes, uint32_t n_bytes,
                                         uint32_t per_vm, uint32_t level) {
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  uint32_t vm = vm_lo + blockIdx.x * 4 + warp;
  if (vm >= vm_hi) return;
  const uint8_t* src = per_vm ? bytes + (size_t)(vm - vm_lo) * n_bytes : bytes;
  uint32_t n_words = (n_bytes + 31) / 32;
  uint32_t* heap = B.heap_mem + (size_t)vm * B.n_slabs * B.heap_words * 8;  // slab 0
  for (uint32_t i = lane; i < n_words * 8; i += 32) {
    uint32_t w = i >> 3, l = i & 7u;
    uint32_t off = 32 * w + 28 - 4 * l;
    uint32_t v = 0;
    for (int k = 0; k < 4; k++) {
      uint32_t a = off + k;
      uint32_t byte = a < n_bytes ? src[a] : 0u;
      v = (v << 8) | byte;
    }
    heap[i] = v;
  }
  if (lane == 0) {
    B.slab_hwm[(size_t)vm * B.n_slabs] = n_words;
    B.lvl[((size_t)vm * (B.max_far_depth + 1) + level) * 4 + 0] = 0;
  }
}

// K5: pack one stream kind of every VM contiguously (VM order) for a single D2H / NCCL send.
__global__ void zkb_pack_kernel(const uint8_t* __restrict__ src, uint64_t stride, const uint64_t* __restrict__ offsets, uint8_t* __restrict__ dst,
                                uint32_t n_vms) {
  uint32_t vm = blockIdx.x;
  if (vm >= n_vms) return;
  uint64_t begin = offsets[vm], n = offsets[vm + 1] - begin;
  const uint8_t* sp = src + (size_t)vm * stride;
  if (((stride | begin | n) & 15u) == 0) {  // every record kind but the 8-byte RefundRec: 16-byte copies
    const uint4* s = reinterpret_cast<const uint4*>(sp);
    uint4* d = reinterpret_cast<uint4*>(dst + begin);
    for (uint64_t i = threadIdx.x; i < n / 16; i += blockDim.x) d[i] = s[i];
  } else {
    const uint2* s = reinterpret_cast<const uint2*>(sp);
    uint2* d = reinterpret_cast<uint2*>(dst + begin);
    for (uint64_t i = threadIdx.x; i < n / 8; i += blockDim.x) d[i] = s[i];
  }
}


// K7: per-VM flattening of the finished batch's storage / event logs: rebuilds what the reference's
// backends hold after the run -- InMemoryStorage::flatten_and_net_history().0 (src/testing/storage.rs:34-76) and
// InMemoryEventSink::flatten() (src/reference_impls/event_sink.rs:66-131) -- from the LOG + FRAME streams, i.e. the
// chronological query history with the rollback queries that finish_frame(panicked) appends in reverse
// (storage.rs:156-180, event_sink.rs:166-170), and the net (never rolled back) events / L1 messages in timestamp order.
// One warp per VM: lanes scan 32 cycle rows at a time for their record counts, records are replayed in program order
// against a stack of rollback marks (the same scheme as the interpreter's storage journal).
struct FlatOut {
  uint32_t* hist[2];   // [vm][cap_hist][32] storage history, event history (LogQueryRec words)
  uint32_t* net[2];    // [vm][cap_net][32]  net events, net L1 messages
  uint32_t* rb[2];     // [vm][cap_net]      indices of not-yet-rolled-back writes / events in hist[k]
  uint32_t* counts;    // [vm][4] records in hist[0], hist[1], net[0], net[1]
  uint32_t* status;    // [vm] 0 ok, 1 VM not ended, 2 capacity, 3 frame stack too deep
  uint32_t cap_hist, cap_net;
};
#define ZKB_FLAT_MAX_DEPTH 256

__global__ void __launch_bounds__(128) zkb_flatten_kernel(const DevBatch B, const FlatOut F) {
  __shared__ uint32_t s_marks[4][ZKB_FLAT_MAX_DEPTH][2];
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  for (uint32_t vm = blockIdx.x * 4 + warp; vm < B.n_vms; vm += gridDim.x * 4) {
    const VmHot* hot = B.hot + vm;
    const uint32_t n_rows = hot->x[X_COUNT0 + ZKB_STREAM_ROWS], n_frames = hot->x[X_COUNT0 + ZKB_STREAM_FRAME];
    const bool ended = hot->x[X_STATUS] == ZKB_VM_ENDED || (hot->x[X_STATUS] == ZKB_VM_RUNNING && hot->live[L_DEPTH - 40] == 0 && hot->x[X_CYCLE] > 0);
    uint32_t* cnt = F.counts + (size_t)vm * 4;
    if (!ended) {
      if (lane < 4) cnt[lane] = 0;
      if (lane == 0) F.status[vm] = 1;
      continue;
    }
    const uint32_t* rows = reinterpret_cast<const uint32_t*>(B.streams[ZKB_STREAM_ROWS] + (size_t)vm * B.cap[ZKB_STREAM_ROWS] * ZKB_ROW_BYTES);
    const uint32_t* logs = reinterpret_cast<const uint32_t*>(B.streams[ZKB_STREAM_LOG] + (size_t)vm * B.cap[ZKB_STREAM_LOG] * ZKB_LOG_BYTES);
    const uint32_t* frames = reinterpret_cast<const uint32_t*>(B.streams[ZKB_STREAM_FRAME] + (size_t)vm * B.cap[ZKB_STREAM_FRAME] * ZKB_FRAME_BYTES);
    uint32_t* hist[2] = {F.hist[0] + (size_t)vm * F.cap_hist * 32, F.hist[1] + (size_t)vm * F.cap_hist * 32};
    uint32_t* rb[2] = {F.rb[0] + (size_t)vm * F.cap_net, F.rb[1] + (size_t)vm * F.cap_net};
    uint32_t n_hist[2] = {0, 0}, n_rb[2] = {0, 0};
    uint32_t sp = 0, il = 0, ifr = 0, st = 0;
    // frames recorded inside cycles; one more record = the bootloader push (helpers.rs:289-316), which opens frame 0
    uint32_t in_rows = 0;
    for (uint32_t base = 0; base < n_rows; base += 32) {
      uint32_t w = base + lane < n_rows ? rows[(size_t)(base + lane) * 64 + 43] : 0u;
      in_rows += (w >> 26) & 3u;
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) in_rows += __shfl_xor_sync(0xffffffffu, in_rows, o);
    if (n_frames == in_rows + 1) {
      if (lane == 0) s_marks[warp][0][0] = s_marks[warp][0][1] = 0;
      sp = 1;
      ifr = 1;
    }
    __syncwarp();
    for (uint32_t base = 0; base < n_rows && st == 0; base += 32) {
      const uint32_t w = base + lane < n_rows ? rows[(size_t)(base + lane) * 64 + 43] : 0u;
      const uint32_t nl = (w >> 16) & 0xFFu, nf = (w >> 26) & 3u;
      uint32_t busy = __ballot_sync(0xffffffffu, (nl | nf) != 0);
      while (busy && st == 0) {
        const int r = __ffs(busy) - 1;
        busy &= busy - 1;
        const uint32_t nl_r = __shfl_sync(0xffffffffu, nl, r), nf_r = __shfl_sync(0xffffffffu, nf, r);
        for (uint32_t j = 0; j < nl_r; j++, il++) {  // log queries of the cycle, in emission order
          uint32_t word = logs[(size_t)il * 32 + lane];
          const uint32_t aux = (__shfl_sync(0xffffffffu, word, 1) >> 16) & 0xFFu;
          const uint32_t rw = __shfl_sync(0xffffffffu, word, 7) & 0xFFu;
          if (aux == ZK_PRECOMPILE_AUX_BYTE) continue;  // precompile calls touch neither backend's log
          const int k = aux == ZK_STORAGE_AUX_BYTE ? 0 : 1;
          if (k == 0 && !rw && lane >= 24) word = 0u;   // the storage's own record of a read keeps written_value = 0 (storage.rs:134-135)
          if (n_hist[k] >= F.cap_hist || (rw && n_rb[k] >= F.cap_net)) {
            st = 2;
            break;
          }
          hist[k][(size_t)n_hist[k] * 32 + lane] = word;
          if (rw) {
            if (lane == 0) rb[k][n_rb[k]] = n_hist[k];
            n_rb[k]++;
          }
          n_hist[k]++;
        }
        for (uint32_t j = 0; j < nf_r && st == 0; j++, ifr++) {  // then the frame start / finish of the cycle
          const uint32_t head = frames[(size_t)ifr * 32];
          if ((head & 0xFFu) == ZKB_FRAMEKIND_START) {
            if (sp >= ZKB_FLAT_MAX_DEPTH) {
              st = 3;
              break;
            }
            if (lane == 0) {
              s_marks[warp][sp][0] = n_rb[0];
              s_marks[warp][sp][1] = n_rb[1];
            }
            sp++;
            __syncwarp();
          } else {
            const bool panicked = ((head >> 8) & 0xFFu) != 0;
            if (sp == 0) {
              st = 3;
              break;
            }
            sp--;
            __syncwarp();
            if (panicked) {
#pragma unroll
              for (int k = 0; k < 2; k++) {
                const uint32_t mark = s_marks[warp][sp][k];
                __syncwarp();
                while (n_rb[k] > mark && st == 0) {  // rollbacks.into_iter().rev() appended to the forward log
                  n_rb[k]--;
                  const uint32_t idx = rb[k][n_rb[k]];
                  uint32_t word = hist[k][(size_t)idx * 32 + lane];
                  if (lane == 7) word |= 1u << 8;  // LogQuery.rollback = true
                  if (n_hist[k] >= F.cap_hist) {
                    st = 2;
                    break;
                  }
                  hist[k][(size_t)n_hist[k] * 32 + lane] = word;
                  n_hist[k]++;
                }
              }
            }
          }
        }
      }
    }
    // net events / L1 messages = the events never rolled back, already in timestamp order (event_sink.rs:82-131)
    uint32_t n_net[2] = {0, 0};
    if (st == 0) {
      __syncwarp();
      for (uint32_t i = 0; i < n_rb[1]; i++) {
        const uint32_t word = hist[1][(size_t)rb[1][i] * 32 + lane];
        const int k = ((__shfl_sync(0xffffffffu, word, 1) >> 16) & 0xFFu) == ZK_EVENT_AUX_BYTE ? 0 : 1;
        F.net[k][((size_t)vm * F.cap_net + n_net[k]) * 32 + lane] = word;
        n_net[k]++;
      }
    }
    if (lane == 0) {
      cnt[0] = st ? 0 : n_hist[0];
      cnt[1] = st ? 0 : n_hist[1];
      cnt[2] = n_net[0];
      cnt[3] = n_net[1];
      F.status[vm] = st;
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
// K8: versioned bytecode hashes (row f-4; ContractCodeSha256 layout parsed at far_call.rs:169-252).  One thread per
// bytecode: sha256 is a serial chain per message, so the parallelism of an ingest batch is across contracts; each thread
// streams its own code as 16-byte loads and keeps the 16-word schedule in registers.  Integer-ALU bound (~1 300
// instructions per 64-byte block), not HBM bound.
__device__ __forceinline__ void sha256_block_regs(uint32_t h[8], uint32_t w[16]) {
  uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
#pragma unroll
  for (int t = 0; t < 64; t++) {
    if (t >= 16) {
      uint32_t w15 = w[(t - 15) & 15], w2 = w[(t - 2) & 15];
      uint32_t s0 = rotr32(w15, 7) ^ rotr32(w15, 18) ^ (w15 >> 3);
      uint32_t s1 = rotr32(w2, 17) ^ rotr32(w2, 19) ^ (w2 >> 10);
      w[t & 15] = w[t & 15] + s0 + w[(t - 7) & 15] + s1;
    }
    uint32_t S1 = rotr32(e, 6) ^ rotr32(e, 11) ^ rotr32(e, 25);
    uint32_t ch = (e & f) ^ (~e & g);
    uint32_t t1 = hh + S1 + ch + c_sha256_k[t] + w[t & 15];
    uint32_t S0 = rotr32(a, 2) ^ rotr32(a, 13) ^ rotr32(a, 22);
    uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
    hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + S0 + mj;
  }
  h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}

__global__ void __launch_bounds__(128) zkb_hash_bytecodes_kernel(const uint8_t* __restrict__ words_be, const uint64_t* __restrict__ offsets_words,
                                                                 uint32_t n, uint32_t marker, uint8_t* __restrict__ hashes_be) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t w0 = offsets_words[i], n_words = offsets_words[i + 1] - w0, n_bytes = n_words * 32;
  const uint4* src = reinterpret_cast<const uint4*>(words_be + w0 * 32);
  uint32_t h[8], w[16];
#pragma unroll
  for (int k = 0; k < 8; k++) h[k] = c_sha256_iv[k];
  const uint64_t full = n_bytes / 64;
  for (uint64_t blk = 0; blk < full; blk++) {
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const uint4 v = __ldg(src + blk * 4 + q);
      w[4 * q + 0] = bswap32(v.x); w[4 * q + 1] = bswap32(v.y); w[4 * q + 2] = bswap32(v.z); w[4 * q + 3] = bswap32(v.w);
    }
    sha256_block_regs(h, w);
  }
  // padding: the code is a whole number of 32-byte words, so the tail is 0 or 32 bytes and always fits one block
#pragma unroll
  for (int k = 0; k < 16; k++) w[k] = 0u;
  if (n_bytes % 64) {
#pragma unroll
    for (int q = 0; q < 2; q++) {
      const uint4 v = __ldg(src + full * 4 + q);
      w[4 * q + 0] = bswap32(v.x); w[4 * q + 1] = bswap32(v.y); w[4 * q + 2] = bswap32(v.z); w[4 * q + 3] = bswap32(v.w);
    }
    w[8] = 0x80000000u;
  } else {
    w[0] = 0x80000000u;
  }
  const uint64_t n_bits = n_bytes * 8;
  w[14] = (uint32_t)(n_bits >> 32);
  w[15] = (uint32_t)n_bits;
  sha256_block_regs(h, w);
  uint32_t* out = reinterpret_cast<uint32_t*>(hashes_be + (size_t)i * 32);
  out[0] = bswap32((uint32_t)ZK_CODE_HASH_VERSION_BYTE << 24 | (marker & 0xFFu) << 16 | ((uint32_t)n_words & 0xFFFFu));
#pragma unroll
  for (int k = 1; k < 8; k++) out[k] = bswap32(h[k]);
}

// K9: sparse part of zkb_restore.  The stack pages and heap slabs are 80 % of a batch's mutable state but a VM touches
// only a prefix of each: [0, high-water mark) -- everything beyond it is zero in the live arrays AND in the snapshot
// (slabs / pages are cleared up to their mark when they are released, memory.rs:181-188).  So restoring
// max(live mark, snapshot mark) words per page is a full restore; the marks themselves (DevBatch.lvl / slab_hwm) are
// restored by the dense copies that follow on the same stream.  One warp per VM, 16-byte copies.
struct SnapView {
  const uint32_t* stack_mem;
  const uint8_t* stack_ptr;
  const uint32_t* heap_mem;
  const uint32_t* lvl;
  const uint32_t* slab_hwm;
};
__global__ void __launch_bounds__(128) zkb_restore_sparse_kernel(const DevBatch B, const SnapView Sn) {
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t levels = B.max_far_depth + 1;
  for (uint32_t vm = blockIdx.x * 4 + warp; vm < B.n_vms; vm += gridDim.x * 4) {
    for (uint32_t l = 0; l < levels; l++) {
      const size_t page = (size_t)vm * levels + l;
      const uint32_t n = min(B.stack_words, max(B.lvl[page * 4 + 2], Sn.lvl[page * 4 + 2]));
      const uint4* src = reinterpret_cast<const uint4*>(Sn.stack_mem + page * B.stack_words * 8);
      uint4* dst = reinterpret_cast<uint4*>(B.stack_mem + page * B.stack_words * 8);
      for (uint32_t i = lane; i < n * 2; i += 32) dst[i] = src[i];
      const uint8_t* psrc = Sn.stack_ptr + page * B.stack_words;
      uint8_t* pdst = B.stack_ptr + page * B.stack_words;
      for (uint32_t i = lane; i < n; i += 32) pdst[i] = psrc[i];
    }
    for (uint32_t sl = 0; sl < B.n_slabs; sl++) {
      const size_t slab = (size_t)vm * B.n_slabs + sl;
      const uint32_t n = min(B.heap_words, max(B.slab_hwm[slab], Sn.slab_hwm[slab]));
      const uint4* src = reinterpret_cast<const uint4*>(Sn.heap_mem + slab * B.heap_words * 8);
      uint4* dst = reinterpret_cast<uint4*>(B.heap_mem + slab * B.heap_words * 8);
      for (uint32_t i = lane; i < n * 2; i += 32) dst[i] = src[i];
    }
  }
}

// K10: one-sided push of a packed stream into a peer GPU's memory (multi-GPU concat, row e): grid-stride 16-byte copies
// whose stores travel over NVLink.  Launched with a handful of 128-thread CTAs so that it co-resides with the persistent
// interpreter (which leaves ~4 K registers and ~40 KB of shared memory per SM unused).
#define ZKB_PUSH_THREADS 128  // 4 warps x 32 registers = 4 K registers: what the interpreter CTA leaves free on an SM (61 440 of 65 536)
template <typename V>  // uint4, or uint2 when a rank's share starts on an 8-byte boundary (RefundRec is 8 bytes)
__global__ void __launch_bounds__(ZKB_PUSH_THREADS) zkb_peer_push_kernel(const V* __restrict__ src, V* __restrict__ dst, uint64_t n16,
                                                            const uint8_t* __restrict__ tail_src, uint8_t* __restrict__ tail_dst, uint32_t n_tail) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  // four independent vector transfers in flight per thread
  for (; i + 3 * stride < n16; i += 4 * stride) {
    const V a = __ldg(src + i), b = __ldg(src + i + stride), c = __ldg(src + i + 2 * stride), d = __ldg(src + i + 3 * stride);
    dst[i] = a;
    dst[i + stride] = b;
    dst[i + 2 * stride] = c;
    dst[i + 3 * stride] = d;
  }
  for (; i < n16; i += stride) dst[i] = __ldg(src + i);
  if (blockIdx.x == 0 && threadIdx.x < n_tail) tail_dst[threadIdx.x] = tail_src[threadIdx.x];
}

static thread_local std::string g_err;
static int32_t set_err(int32_t code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define CUDA_OK(expr)                                                                                         \
  do {                                                                                                        \
    cudaError_t _e = (expr);                                                                                  \
    if (_e != cudaSuccess) return set_err(ZKB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
  } while (0)

static const uint32_t REC_BYTES[ZKB_N_STREAMS] = {ZKB_ROW_BYTES, ZKB_MEM_BYTES, ZKB_LOG_BYTES, ZKB_DECOMMIT_BYTES, ZKB_FRAME_BYTES, ZKB_REFUND_BYTES};

struct Bytecode {
  uint32_t hash[8];
  uint32_t offset_words, len_words;
};

struct ZkbBatch {
  ZkbConfig cfg;
  DevBatch d;
  std::vector<VmHot> h_hot;
  std::vector<uint32_t> h_root;  // [n_vms][32] root frames (callstack slot 0)
  std::vector<uint32_t> h_bootrec;  // [n_vms][32] FrameRec of the bootloader push (start_new_execution_context)
  bool hot_dirty = true;         // host mirror newer than device
  bool hot_stale = false;        // device newer than host mirror
  bool root_dirty = true;
  std::vector<Bytecode> codes;
  std::vector<uint32_t> h_code_words;
  bool codes_dirty = true;
  uint32_t* d_code_words = nullptr;
  uint32_t* d_code_meta = nullptr;
  uint32_t* d_code_index = nullptr;   // open-addressed hash index over the loaded bytecodes (DevBatch.code_index)
  std::vector<int32_t> boot_code;  // per VM: code id bound by populate_code (page in boot_page)
  uint32_t* d_order = nullptr;     // schedule slot -> VM, VMs grouped by boot_code (nullptr while the identity is already grouped)
  bool order_dirty = true;
  bool regroup = true;             // ZKB_REGROUP=0 turns the grouping off (experiment)
  std::vector<uint32_t> boot_page;
  uint32_t* d_fail = nullptr;   // device pointer of h_fail (mapped pinned: reading it never queues behind D2H copies)
  uint32_t* h_fail = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaStream_t last_stream = nullptr;
  uint32_t n_launches = 0;
  int grid = 0;
  bool lockstep = false;
  bool balance_waves = false;   // lockstep: even out the last wave of VMs over all SMs (ZKB_BALANCE=1; see zkb_run)
  uint8_t* d_pack = nullptr;
  uint64_t pack_capacity = 0;
  uint64_t* d_offsets = nullptr;                  // [ZKB_N_STREAMS][n_vms + 1]
  uint64_t* h_offsets[ZKB_N_STREAMS] = {};         // PINNED host copies (a pageable source would make the "async" upload
                                                  // wait for everything already queued on the stream)
  bool offsets_valid = false;
  uint32_t* h_counts = nullptr;                   // mapped pinned mirror of DevBatch.host_counts
  bool counts_valid = false;                      // h_counts reflects the device state (after upload / a finished run)
  uint8_t* d_stage = nullptr;                     // grow-only H2D staging buffer of the populate_* calls
  uint64_t stage_capacity = 0;
  cudaEvent_t ev_setup = nullptr;
  bool launched = false;
  std::vector<void*> allocs;
  uint64_t h2d_bytes = 0, d2h_bytes = 0;
  // checkpoint (VmLocalState: Clone, vm_state/mod.rs:53): device copies of every mutable per-VM array
  struct Region {
    void* live;
    void* saved;
    size_t bytes;
    bool sparse;  // stack pages / heap slabs: restored by zkb_restore_sparse_kernel (touched prefixes only)
  };
  FlatOut flat{};
  bool flat_allocated = false, flat_valid = false;
  std::vector<uint32_t> h_flat_counts, h_flat_status;
  std::vector<Region> snap;
  std::vector<uint32_t> snap_counts;
  bool has_snapshot = false;
  // pages_with_extended_lifetime[BOOTLOADER_CALLDATA_PAGE] (memory.rs:230-231,293-298): host-resident, because the
  // reference registers no indirection for it -- no VM instruction can read it (see zkb_set_calldata in zkb.h)
  std::vector<std::vector<uint8_t>> calldata;
  // transport encoder (zkb_codec.h): per-(VM, stream) sizes, their prefix sums, the blob, totals in mapped host memory
  uint32_t* d_enc_sizes = nullptr;
  uint64_t* d_enc_offsets = nullptr;
  uint64_t* d_enc_canon = nullptr;      // [6][n_vms + 1] canonical byte offsets (device copy of h_offsets)
  uint8_t* d_enc_stage = nullptr;       // staging area of the encoding pass (grow-only)
  uint64_t enc_stage_capacity = 0;
  uint64_t* h_enc_totals = nullptr;
  uint64_t* d_enc_totals = nullptr;
  uint8_t* d_enc = nullptr;
  uint64_t enc_capacity = 0;
  cudaEvent_t ev_enc = nullptr, ev_enc_done = nullptr, ev_blob_read = nullptr;
  // device-side consumer (consume.cuh): the VMs' pre-run state, snapshots, queue digests
  VmHot* d_hot_init = nullptr;
  ConsumeOut consume{};
  bool consume_valid = false;
  std::vector<uint32_t> h_n_snaps;
  uint8_t* d_snap_pack = nullptr;
  uint64_t snap_pack_capacity = 0;
  cudaStream_t enc_stream = nullptr;   // the encoder's own stream: its passes never queue behind a D2H copy in flight
  bool blob_read_pending = false;
};

static void be32_to_limbs(const uint8_t* be, uint32_t* limbs) {
  for (int l = 0; l < 8; l++) {
    const uint8_t* p = be + 28 - 4 * l;
    limbs[l] = (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3];
  }
}
static void limbs_to_be32(const uint32_t* limbs, uint8_t* be) {
  for (int l = 0; l < 8; l++) {
    uint8_t* p = be + 28 - 4 * l;
    p[0] = (uint8_t)(limbs[l] >> 24);
    p[1] = (uint8_t)(limbs[l] >> 16);
    p[2] = (uint8_t)(limbs[l] >> 8);
    p[3] = (uint8_t)limbs[l];
  }
}

static void init_hot(const ZkbBatch* b, VmHot& h) {
  memset(&h, 0, sizeof(h));
  // CallStackEntry::empty_context (execution_stack.rs:35-55)
  h.F[F_SP_PC] = ZK_INITIAL_SP_ON_FAR_CALL;
  h.F[F_ERGS] = ZK_VM_INITIAL_FRAME_ERGS;
  h.F[F_CODE_ID] = ZKB_NO_CODE;
  h.F[F_FAR_LEVEL] = 0;
  // VmLocalState::empty_state (vm_state/mod.rs:76-94)
  h.live[L_PAGE_COUNTER - 40] = ZK_STARTING_BASE_PAGE;
  h.live[L_EH_BITS - 40] = ZKB_FRAMEBIT_KERNEL << 16;  // address 0 is a kernel address
  h.x[X_TIMESTAMP] = ZK_STARTING_TIMESTAMP;
  h.x[X_STATUS] = ZKB_VM_RUNNING;
  h.x[X_SLAB_FREE] = b->cfg.n_heap_slabs >= 32 ? 0xFFFFFFFFu : ((1u << b->cfg.n_heap_slabs) - 1u);
}

template <class T>
static cudaError_t dalloc(ZkbBatch* b, T** p, size_t count, bool zero) {
  void* q = nullptr;
  size_t bytes = std::max<size_t>(count * sizeof(T), 16);
  cudaError_t e = cudaMalloc(&q, bytes);
  if (e != cudaSuccess) return e;
  b->allocs.push_back(q);
  if (zero) e = cudaMemset(q, 0, bytes);
  *p = (T*)q;
  return e;
}

// device staging for host inputs: grow-only, so the steady state of a pipelined host loop never calls
// cudaMalloc/cudaFree (both synchronise the whole device and would serialise against in-flight D2H copies)
static int32_t stage_h2d(ZkbBatch* b, const void* src, size_t bytes, uint8_t** out) {
  if (bytes > b->stage_capacity) {
    if (b->d_stage) CUDA_OK(cudaFree(b->d_stage));
    b->d_stage = nullptr;
    b->stage_capacity = 0;
    size_t cap = std::max<size_t>(bytes + bytes / 4, 1 << 20);
    CUDA_OK(cudaMalloc(&b->d_stage, cap));
    b->stage_capacity = cap;
  }
  CUDA_OK(cudaMemcpy(b->d_stage, src, bytes, cudaMemcpyHostToDevice));
  b->h2d_bytes += bytes;
  *out = b->d_stage;
  return ZKB_OK;
}

// wait for this batch's last launch only (not the device): other batches' kernels and copies keep running
static int32_t wait_last_run(ZkbBatch* b) {
  if (b->launched) CUDA_OK(cudaEventSynchronize(b->ev1));
  return ZKB_OK;
}

static int32_t clear_cold_state(ZkbBatch* b) {
  const ZkbConfig& c = b->cfg;
  size_t n = c.n_vms, levels = c.max_far_depth + 1;
  CUDA_OK(cudaMemset(b->d.stack_mem, 0, n * levels * c.stack_words * 32));
  CUDA_OK(cudaMemset(b->d.stack_ptr, 0, n * levels * c.stack_words));
  CUDA_OK(cudaMemset(b->d.heap_mem, 0, n * c.n_heap_slabs * (size_t)c.heap_bytes));
  CUDA_OK(cudaMemset(b->d.slab_hwm, 0, n * c.n_heap_slabs * 4));
  CUDA_OK(cudaMemset(b->d.st_tags, 0, n * c.storage_slots * 4));
  CUDA_OK(cudaMemset(b->d.st_vals, 0, n * c.storage_slots * 32));
  zkb_init_kernel<<<(c.n_vms + 127) / 128, 128>>>(b->d);
  CUDA_OK(cudaGetLastError());
  return ZKB_OK;
}

static int32_t upload(ZkbBatch* b) {
  if (b->codes_dirty) {
    if (b->d_code_words) cudaFree(b->d_code_words);
    if (b->d_code_meta) cudaFree(b->d_code_meta);
    b->d_code_words = b->d_code_meta = nullptr;
    size_t nw = std::max<size_t>(b->h_code_words.size(), 8);
    CUDA_OK(cudaMalloc(&b->d_code_words, nw * 4));
    CUDA_OK(cudaMalloc(&b->d_code_meta, std::max<size_t>(b->codes.size(), 1) * 40));
    if (!b->h_code_words.empty()) CUDA_OK(cudaMemcpy(b->d_code_words, b->h_code_words.data(), b->h_code_words.size() * 4, cudaMemcpyHostToDevice));
    std::vector<uint32_t> meta(b->codes.size() * 10);
    for (size_t i = 0; i < b->codes.size(); i++) {
      meta[i * 10] = b->codes[i].offset_words;
      meta[i * 10 + 1] = b->codes[i].len_words;
      memcpy(&meta[i * 10 + 2], b->codes[i].hash, 32);
    }
    if (!meta.empty()) CUDA_OK(cudaMemcpy(b->d_code_meta, meta.data(), meta.size() * 4, cudaMemcpyHostToDevice));
    b->d.code_words = b->d_code_words;
    b->d.code_meta = b->d_code_meta;
    b->d.n_codes = (uint32_t)b->codes.size();
    // hash index for the decommitter's lookup (vm.cuh op_far_call): at most half full, linear probing from the low limb
    uint32_t slots = 16;
    while (slots < 2 * b->codes.size()) slots *= 2;
    std::vector<uint32_t> index(slots, ZKB_NO_CODE);
    for (size_t i = 0; i < b->codes.size(); i++) {
      uint32_t s = b->codes[i].hash[0] & (slots - 1);
      while (index[s] != ZKB_NO_CODE) s = (s + 1) & (slots - 1);
      index[s] = (uint32_t)i;
    }
    if (b->d_code_index) cudaFree(b->d_code_index);
    b->d_code_index = nullptr;
    CUDA_OK(cudaMalloc(&b->d_code_index, (size_t)slots * 4));
    CUDA_OK(cudaMemcpy(b->d_code_index, index.data(), (size_t)slots * 4, cudaMemcpyHostToDevice));
    b->d.code_index = b->d_code_index;
    b->d.code_index_mask = slots - 1;
    b->codes_dirty = false;
  }
  if (b->hot_dirty) {
    CUDA_OK(cudaMemcpy(b->d.hot, b->h_hot.data(), b->h_hot.size() * sizeof(VmHot), cudaMemcpyHostToDevice));
    b->h2d_bytes += b->h_hot.size() * sizeof(VmHot);
    // the consumer (zkb_consume) rebuilds every VmLocalState from the state in front of the first recorded cycle
    CUDA_OK(cudaMemcpyAsync(b->d_hot_init, b->d.hot, b->h_hot.size() * sizeof(VmHot), cudaMemcpyDeviceToDevice, 0));
    b->hot_dirty = false;
    for (size_t v = 0; v < b->h_hot.size(); v++) {
      const VmHot& h = b->h_hot[v];
      for (int k = 0; k < ZKB_N_STREAMS; k++) b->h_counts[v * 8 + k] = h.x[X_COUNT0 + k];
      b->h_counts[v * 8 + 6] = h.x[X_STATUS];
      b->h_counts[v * 8 + 7] = h.x[X_CYCLE];
    }
    b->counts_valid = true;
    b->offsets_valid = false;
  }
  if (b->root_dirty) {
    CUDA_OK(cudaMemcpy2D(b->d.callstack, (size_t)b->cfg.max_depth * 128, b->h_root.data(), 128, 128, b->cfg.n_vms, cudaMemcpyHostToDevice));
    if (b->cfg.witness_mode && b->cfg.cap_records[ZKB_STREAM_FRAME] > 0)
      CUDA_OK(cudaMemcpy2D(b->d.streams[ZKB_STREAM_FRAME], (size_t)b->cfg.cap_records[ZKB_STREAM_FRAME] * ZKB_FRAME_BYTES, b->h_bootrec.data(), 128,
                           128, b->cfg.n_vms, cudaMemcpyHostToDevice));
    b->h2d_bytes += (size_t)b->cfg.n_vms * 256;
    b->root_dirty = false;
  }
  return ZKB_OK;
}

static int32_t download_hot(ZkbBatch* b) {
  if (b->hot_stale) {
    CUDA_OK(cudaSetDevice(b->cfg.device));
    if (wait_last_run(b) != ZKB_OK) return ZKB_ERR_CUDA;
    CUDA_OK(cudaMemcpy(b->h_hot.data(), b->d.hot, b->h_hot.size() * sizeof(VmHot), cudaMemcpyDeviceToHost));
    b->d2h_bytes += b->h_hot.size() * sizeof(VmHot);
    b->hot_stale = false;
  }
  return ZKB_OK;
}

// per-VM summary (stream counts, status, cycles) without a device->host copy: valid after upload or a finished run
static int32_t summary(ZkbBatch* b, const uint32_t** out) {
  if (b->hot_dirty) {  // host-side edits not uploaded yet: the host mirror is authoritative
    int32_t rc = upload(b);
    if (rc != ZKB_OK) return rc;
  }
  if (wait_last_run(b) != ZKB_OK) return ZKB_ERR_CUDA;
  *out = b->h_counts;
  return ZKB_OK;
}

static bool range_ok(ZkbBatch* b, uint32_t lo, uint32_t hi) { return b && lo <= hi && hi <= b->cfg.n_vms; }

static int find_code(ZkbBatch* b, const uint8_t hash_be[32]) {
  uint32_t limbs[8];
  be32_to_limbs(hash_be, limbs);
  for (size_t i = 0; i < b->codes.size(); i++)
    if (memcmp(b->codes[i].hash, limbs, 32) == 0) return (int)i;
  return -1;
}

extern "C" {

const char* zkb_last_error(void) { return g_err.c_str(); }

int32_t zkb_create(const ZkbConfig* cfg, ZkbBatch** out) {
  if (!cfg || !out || cfg->n_vms == 0) return set_err(ZKB_ERR_INVALID_ARGUMENT, "null config / zero VMs");
  if (cfg->heap_bytes % 32 || cfg->n_heap_slabs == 0 || cfg->n_heap_slabs > 32 || cfg->storage_slots < 32 ||
      (cfg->storage_slots & (cfg->storage_slots - 1)) || cfg->max_far_depth == 0 || cfg->max_depth < 2 || cfg->stack_words == 0 ||
      cfg->max_far_depth > 250)
    return set_err(ZKB_ERR_INVALID_ARGUMENT, "bad capacity in ZkbConfig");
  if (cfg->reserved[1] > ZK_INITIAL_STORAGE_WRITE_PUBDATA_BYTES)
    return set_err(ZKB_ERR_INVALID_ARGUMENT, "warm_write_refund_bytes above INITIAL_STORAGE_WRITE_PUBDATA_BYTES (log.rs:110 asserts refund <= net cost)");
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) return set_err(ZKB_ERR_NO_DEVICE, "no CUDA device: this library has no CPU fallback");
  CUDA_OK(cudaSetDevice(cfg->device));
  ZkbBatch* b = new ZkbBatch();
  b->cfg = *cfg;
  const ZkbConfig& c = b->cfg;
  DevBatch& d = b->d;
  memset(&d, 0, sizeof(d));
  d.n_vms = c.n_vms;
  d.witness = c.witness_mode;
  for (int k = 0; k < ZKB_N_STREAMS; k++) d.cap[k] = c.cap_records[k];
  d.stack_words = c.stack_words;
  d.heap_words = c.heap_bytes / 32;
  d.n_slabs = c.n_heap_slabs;
  d.warm_refund_bytes = c.reserved[1];
  d.max_far_depth = c.max_far_depth;
  d.max_depth = c.max_depth;
  d.storage_slots = c.storage_slots;
  d.journal_entries = std::max(1u, c.journal_entries);
  size_t n = c.n_vms, levels = c.max_far_depth + 1;
  cudaError_t e = cudaSuccess;
#define ALLOC(ptr, count, zero)                                                                    \
  if (e == cudaSuccess) e = dalloc(b, &(ptr), (count), (zero));
  ALLOC(d.hot, n, false);
  ALLOC(d.callstack, n * c.max_depth * 32, false);
  ALLOC(d.stack_mem, n * levels * c.stack_words * 8, false);
  ALLOC(d.stack_ptr, n * levels * c.stack_words, false);
  ALLOC(d.heap_mem, n * c.n_heap_slabs * (size_t)d.heap_words * 8, false);
  ALLOC(d.lvl, n * levels * 4, false);
  ALLOC(d.slab_hwm, n * c.n_heap_slabs, false);
  ALLOC(d.pt, n * ZKB_PT_ENTRIES * 2, false);
  ALLOC(d.dec, n * ZKB_DEC_ENTRIES * 2, true);
  ALLOC(d.st_tags, n * c.storage_slots, false);
  ALLOC(d.st_keys, n * c.storage_slots * 8, true);
  ALLOC(d.st_addr, n * c.storage_slots * 8, true);
  ALLOC(d.st_vals, n * c.storage_slots * 8, false);
  ALLOC(d.j_slot, n * d.journal_entries, true);
  ALLOC(d.j_val, n * d.journal_entries * 8, true);
  if (c.witness_mode)
    for (int k = 0; k < ZKB_N_STREAMS; k++) ALLOC(d.streams[k], n * (size_t)c.cap_records[k] * REC_BYTES[k], false);
  ALLOC(d.queue, 4, true);
  ALLOC(d.defer, n * ZKB_DEFER_WORDS, true);
  ALLOC(b->d_hot_init, n, false);
  ALLOC(b->d_offsets, (n + 1) * ZKB_N_STREAMS, false);
#undef ALLOC
  if (e != cudaSuccess) {
    std::string msg = std::string("cudaMalloc: ") + cudaGetErrorString(e);
    for (void* p : b->allocs) cudaFree(p);
    delete b;
    cudaGetLastError();
    return set_err(ZKB_ERR_OUT_OF_MEMORY, msg);
  }
  {
    void* hp = nullptr;
    cudaError_t he = cudaHostAlloc(&hp, n * 8 * sizeof(uint32_t), cudaHostAllocMapped);
    void* dp = nullptr;
    if (he == cudaSuccess) he = cudaHostGetDevicePointer(&dp, hp, 0);
    if (he != cudaSuccess) {
      for (void* p : b->allocs) cudaFree(p);
      delete b;
      return set_err(ZKB_ERR_OUT_OF_MEMORY, std::string("cudaHostAlloc: ") + cudaGetErrorString(he));
    }
    void* fp = nullptr;
    void* fdp = nullptr;
    if (cudaHostAlloc(&fp, 64, cudaHostAllocMapped) == cudaSuccess && cudaHostGetDevicePointer(&fdp, fp, 0) == cudaSuccess) {
      b->h_fail = (uint32_t*)fp;
      b->d_fail = (uint32_t*)fdp;
      *b->h_fail = 0;
    }
    b->h_counts = (uint32_t*)hp;
    d.host_counts = (uint32_t*)dp;
    memset(hp, 0, n * 8 * sizeof(uint32_t));
    void* op = nullptr;
    if (cudaHostAlloc(&op, (n + 1) * ZKB_N_STREAMS * sizeof(uint64_t), cudaHostAllocDefault) != cudaSuccess) {
      cudaFreeHost(hp);
      for (void* p : b->allocs) cudaFree(p);
      delete b;
      return set_err(ZKB_ERR_OUT_OF_MEMORY, "cudaHostAlloc (stream offsets)");
    }
    for (int k = 0; k < ZKB_N_STREAMS; k++) b->h_offsets[k] = (uint64_t*)op + (size_t)k * (n + 1);
  }
  b->h_hot.resize(n);
  b->h_root.assign(n * 32, 0);
  b->h_bootrec.assign(n * 32, 0);
  b->boot_code.assign(n, -1);
  b->boot_page.assign(n, 0);
  for (auto& h : b->h_hot) init_hot(b, h);
  memset(d.default_aa, 0, sizeof(d.default_aa));
  int32_t rc = clear_cold_state(b);
  if (rc != ZKB_OK) return rc;
  CUDA_OK(cudaEventCreate(&b->ev0));
  CUDA_OK(cudaEventCreate(&b->ev1));
  CUDA_OK(cudaEventCreateWithFlags(&b->ev_setup, cudaEventDisableTiming));
  int per_sm = 0, n_sm = 0;
  CUDA_OK(cudaFuncSetAttribute(zkb_run_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ZKB_RUN_SMEM_BYTES));
  CUDA_OK(cudaFuncSetAttribute(zkb_run_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ZKB_RUN_SMEM_BYTES));
  CUDA_OK(cudaFuncSetAttribute(zkb_run_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ZKB_RUN_SMEM_BYTES));
  CUDA_OK(cudaFuncSetAttribute(zkb_run_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ZKB_RUN_SMEM_BYTES));
  CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, zkb_run_kernel<false, false>, ZKB_WARPS_PER_CTA * 32, ZKB_RUN_SMEM_BYTES));
  CUDA_OK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, cfg->device));
  b->grid = std::max(1, per_sm * n_sm);  // persistent grid: a multiple of the SM count (148 on B200)
  if (cfg->reserved[0] > 0 && (int)cfg->reserved[0] < n_sm) b->grid = per_sm * (n_sm - (int)cfg->reserved[0]);
  uint32_t sched = cfg->schedule;
  if (const char* env = getenv("ZKB_SCHEDULE")) sched = (uint32_t)atoi(env);  // experiment override
  b->lockstep = sched != ZKB_SCHED_FREE;
  if (const char* env = getenv("ZKB_BALANCE")) b->balance_waves = atoi(env) != 0;
  if (const char* env = getenv("ZKB_REGROUP")) b->regroup = atoi(env) != 0;
  *out = b;
  return ZKB_OK;
}

int32_t zkb_destroy(ZkbBatch* b) {
  if (!b) return ZKB_OK;
  cudaSetDevice(b->cfg.device);
  cudaDeviceSynchronize();
  for (void* p : b->allocs) cudaFree(p);
  if (b->d_code_words) cudaFree(b->d_code_words);
  if (b->d_code_meta) cudaFree(b->d_code_meta);
  if (b->d_code_index) cudaFree(b->d_code_index);
  if (b->d_pack) cudaFree(b->d_pack);
  for (auto& r : b->snap)
    if (r.saved) cudaFree(r.saved);
  if (b->ev0) cudaEventDestroy(b->ev0);
  if (b->ev1) cudaEventDestroy(b->ev1);
  if (b->ev_setup) cudaEventDestroy(b->ev_setup);
  if (b->d_stage) cudaFree(b->d_stage);
  if (b->h_counts) cudaFreeHost(b->h_counts);
  if (b->h_fail) cudaFreeHost(b->h_fail);
  if (b->h_offsets[0]) cudaFreeHost(b->h_offsets[0]);
  if (b->d_enc) cudaFree(b->d_enc);
  if (b->d_enc_stage) cudaFree(b->d_enc_stage);
  if (b->d_snap_pack) cudaFree(b->d_snap_pack);
  if (b->h_enc_totals) cudaFreeHost(b->h_enc_totals);
  if (b->ev_enc) cudaEventDestroy(b->ev_enc);
  if (b->ev_enc_done) cudaEventDestroy(b->ev_enc_done);
  if (b->ev_blob_read) cudaEventDestroy(b->ev_blob_read);
  if (b->enc_stream) cudaStreamDestroy(b->enc_stream);
  delete b;
  return ZKB_OK;
}

int32_t zkb_reset(ZkbBatch* b) {
  if (!b) return ZKB_ERR_INVALID_ARGUMENT;
  CUDA_OK(cudaSetDevice(b->cfg.device));
  if (wait_last_run(b) != ZKB_OK) return ZKB_ERR_CUDA;
  for (auto& h : b->h_hot) init_hot(b, h);
  std::fill(b->h_root.begin(), b->h_root.end(), 0u);
  std::fill(b->h_bootrec.begin(), b->h_bootrec.end(), 0u);
  std::fill(b->boot_code.begin(), b->boot_code.end(), -1);
  b->order_dirty = true;
  b->calldata.clear();
  b->hot_dirty = b->root_dirty = true;
  b->hot_stale = false;
  return clear_cold_state(b);
}

int32_t zkb_load_bytecode(ZkbBatch* b, const uint8_t hash_be[32], const uint8_t* words_be, uint32_t n_words) {
  if (!b || !hash_be || (!words_be && n_words)) return ZKB_ERR_INVALID_ARGUMENT;
  if (find_code(b, hash_be) >= 0) return set_err(ZKB_ERR_INVALID_ARGUMENT, "bytecode hash already loaded (decommitter.rs:25)");
  Bytecode bc;
  be32_to_limbs(hash_be, bc.hash);
  bc.offset_words = (uint32_t)(b->h_code_words.size() / 8);
  bc.len_words = n_words;
  b->h_code_words.resize(b->h_code_words.size() + (size_t)n_words * 8);
  for (uint32_t i = 0; i < n_words; i++) be32_to_limbs(words_be + 32 * (size_t)i, &b->h_code_words[((size_t)bc.offset_words + i) * 8]);
  b->codes.push_back(bc);
  b->codes_dirty = true;
  return ZKB_OK;
}

int32_t zkb_read_bytecode(ZkbBatch* b, const uint8_t hash_be[32], uint8_t* words_be_out, uint32_t max_words, uint32_t* n_words_out) {
  if (!b || !hash_be || (!words_be_out && max_words)) return ZKB_ERR_INVALID_ARGUMENT;
  int id = find_code(b, hash_be);
  if (id < 0) return set_err(ZKB_ERR_UNKNOWN_BYTECODE, "read_bytecode: bytecode hash not loaded");
  const Bytecode& bc = b->codes[id];
  if (n_words_out) *n_words_out = bc.len_words;
  for (uint32_t i = 0; i < std::min(max_words, bc.len_words); i++)
    limbs_to_be32(&b->h_code_words[((size_t)bc.offset_words + i) * 8], words_be_out + 32 * (size_t)i);
  return ZKB_OK;
}

int32_t zkb_set_calldata(ZkbBatch* b, uint32_t vm_lo, uint32_t vm_hi, const uint8_t* words_be, uint32_t n_words, uint32_t per_vm) {
  if (!range_ok(b, vm_lo, vm_hi) || (!words_be && n_words)) return ZKB_ERR_INVALID_ARGUMENT;
  if (b->calldata.empty()) b->calldata.resize(b->cfg.n_vms);
  for (uint32_t v = vm_lo; v < vm_hi; v++) {
    const uint8_t* src = per_vm ? words_be + (size_t)(v - vm_lo) * n_words * 32 : words_be;
    b->calldata[v].assign(src, src + (size_t)n_words * 32);
  }
  return ZKB_OK;
}

int32_t zkb_read_calldata(ZkbBatch* b, uint32_t vm, uint32_t word_lo, uint32_t n_words, uint8_t* words_be_out) {
  if (!b || vm >= b->cfg.n_vms || (!words_be_out && n_words)) return ZKB_ERR_INVALID_ARGUMENT;
  memset(words_be_out, 0, (size_t)n_words * 32);
  if (b->calldata.empty()) return ZKB_OK;
  const std::vector<uint8_t>& page = b->calldata[vm];
  for (uint32_t i = 0; i < n_words; i++) {
    const size_t at = ((size_t)word_lo + i) * 32;
    if (at + 32 <= page.size()) memcpy(words_be_out + 32 * (size_t)i, page.data() + at, 32);
  }
  return ZKB_OK;
}

int32_t zkb_set_block_properties(ZkbBatch* b, const uint8_t default_aa_code_hash_be[32], uint8_t zkporter_is_available) {
  if (!b || !default_aa_code_hash_be) return ZKB_ERR_INVALID_ARGUMENT;
  be32_to_limbs(default_aa_code_hash_be, b->d.default_aa);
  b->d.zkporter = zkporter_is_available ? 1u : 0u;
  return ZKB_OK;
}

int32_t zkb_populate_storage(ZkbBatch* b, uint32_t vm_lo, uint32_t vm_hi, const ZkbStorageInit* entries, uint32_t n, uint32_t per_vm) {
  if (!range_ok(b, vm_lo, vm_hi) || (!entries && n)) return ZKB_ERR_INVALID_ARGUMENT;
  if (n == 0 || vm_lo == vm_hi) return ZKB_OK;
  CUDA_OK(cudaSetDevice(b->cfg.device));
  size_t total = per_vm ? (size_t)(vm_hi - vm_lo) * n : n;
  std::vector<DevStorageInit> h(total);
  for (size_t i = 0; i < total; i++) {
    h[i].shard = entries[i].shard_id;
    memcpy(h[i].addr, entries[i].address, 20);
    be32_to_limbs(entries[i].key_be, h[i].key);
    be32_to_limbs(entries[i].value_be, h[i].value);
  }
  uint8_t* staged = nullptr;
  int32_t src = stage_h2d(b, h.data(), total * sizeof(DevStorageInit), &staged);
  if (src != ZKB_OK) return src;
  DevStorageInit* d_e = reinterpret_cast<DevStorageInit*>(staged);
  if (!b->h_fail) return set_err(ZKB_ERR_OUT_OF_MEMORY, "no mapped host flag");
  *b->h_fail = 0;
  zkb_populate_storage_kernel<<<(vm_hi - vm_lo + 15) / 16, 128>>>(b->d, vm_lo, vm_hi, d_e, n, per_vm, b->d_fail);
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaStreamSynchronize(0));  // orders the kernel before the staging buffer is reused; the flag is host-mapped
  uint32_t fail = *(volatile uint32_t*)b->h_fail;
  if (fail) return set_err(ZKB_ERR_INVALID_ARGUMENT, "storage table capacity exceeded while populating (raise ZkbConfig.storage_slots)");
  return ZKB_OK;
}

int32_t zkb_populate_code(ZkbBatch* b, uint32_t vm_lo, uint32_t vm_hi, uint32_t page, const uint8_t hash_be[32]) {
  if (!range_ok(b, vm_lo, vm_hi) || !hash_be) return ZKB_ERR_INVALID_ARGUMENT;
  int id = find_code(b, hash_be);
  if (id < 0) return set_err(ZKB_ERR_UNKNOWN_BYTECODE, "populate_code: bytecode hash not loaded");
  for (uint32_t v = vm_lo; v < vm_hi; v++) {
    b->boot_code[v] = id;
    b->order_dirty = true;
    b->boot_page[v] = page;
  }
  return ZKB_OK;
}

int32_t zkb_push_bootloader_context(ZkbBatch* b, uint32_t vm_lo, uint32_t vm_hi, const ZkbFrame* f) {
  if (!range_ok(b, vm_lo, vm_hi) || !f) return ZKB_ERR_INVALID_ARGUMENT;
  if (b->cfg.max_depth < 2 || b->cfg.max_far_depth < 1) return ZKB_ERR_INVALID_ARGUMENT;
  if (download_hot(b) != ZKB_OK) return ZKB_ERR_CUDA;
  for (uint32_t v = vm_lo; v < vm_hi; v++) {
    VmHot& h = b->h_hot[v];
    if (h.live[L_DEPTH - 40] != 0) return set_err(ZKB_ERR_INVALID_ARGUMENT, "bootloader context already pushed");
    // helpers.rs:295-303: the root frame keeps the difference
    if (h.F[F_ERGS] < f->ergs_remaining) return set_err(ZKB_ERR_INVALID_ARGUMENT, "trying to create bootloader frame with more ergs than VM has available");
    h.F[F_ERGS] -= f->ergs_remaining;
    h.F[F_JOURNAL_MARK] = 0;
    memcpy(&b->h_root[(size_t)v * 32], h.F, 128);
    uint32_t* F = h.F;
    memset(F, 0, 128);
    memcpy(&F[F_THIS], f->this_address, 20);
    memcpy(&F[F_SENDER], f->msg_sender, 20);
    memcpy(&F[F_CODE_ADDR], f->code_address, 20);
    F[F_BASE_PAGE] = f->base_memory_page;
    F[F_CODE_PAGE] = f->code_page;
    F[F_SP_PC] = (uint32_t)f->sp | (uint32_t)f->pc << 16;
    F[F_EH_SHARDS] = (uint32_t)f->exception_handler_location | (uint32_t)f->this_shard_id << 16 | (uint32_t)f->caller_shard_id << 24;
    F[F_ERGS] = f->ergs_remaining;
    F[F_MISC] = (uint32_t)f->code_shard_id | (uint32_t)(f->is_static ? 1 : 0) << 8 | (uint32_t)(f->is_local_frame ? 1 : 0) << 16;
    memcpy(&F[F_CTX], f->context_u128_value, 16);
    F[F_HEAP_BOUND] = f->heap_bound;
    F[F_AUX_BOUND] = f->aux_heap_bound;
    F[F_CODE_ID] = (b->boot_code[v] >= 0 && b->boot_page[v] == f->code_page) ? (uint32_t)b->boot_code[v] : ZKB_NO_CODE;
    F[F_JOURNAL_MARK] = 0;
    F[F_FAR_LEVEL] = 1;  // memory.start_global_frame(UNMAPPED_PAGE, base, empty ptr) (helpers.rs:308-315)
    bool kernel = F[0] == 0 && F[1] == 0 && F[2] == 0 && F[3] == 0 && (F[4] & 0xFFFFu) == 0;
    uint32_t bits = (f->is_static ? ZKB_FRAMEBIT_STATIC : 0u) | (f->is_local_frame ? ZKB_FRAMEBIT_LOCAL : 0u) | (kernel ? ZKB_FRAMEBIT_KERNEL : 0u);
    h.live[L_DEPTH - 40] = 1;
    h.live[L_CODE_PAGE - 40] = f->code_page;
    h.live[L_BASE_PAGE - 40] = f->base_memory_page;
    h.live[L_HEAP_BOUND - 40] = f->heap_bound;
    h.live[L_AUX_BOUND - 40] = f->aux_heap_bound;
    h.live[L_EH_BITS - 40] = (uint32_t)f->exception_handler_location | bits << 16;
    h.x[X_FAR_DEPTH] = 1;
    // start_frame -> witness_tracer.start_new_execution_context (helpers.rs:237-241): the VM's first frame record
    uint32_t* rec = &b->h_bootrec[(size_t)v * 32];
    memset(rec, 0, 128);
    rec[0] = ZKB_FRAMEKIND_START;
    rec[1] = h.x[X_CYCLE];
    memcpy(&rec[2], F, 27 * 4);
    rec[29] = b->h_root[(size_t)v * 32 + F_ERGS];
    rec[30] = b->h_root[(size_t)v * 32 + F_SP_PC] >> 16 | (b->h_root[(size_t)v * 32 + F_SP_PC] & 0xFFFFu) << 16;
    if (h.x[X_COUNT0 + ZKB_STREAM_FRAME] == 0) h.x[X_COUNT0 + ZKB_STREAM_FRAME] = 1;
  }
  b->hot_dirty = b->root_dirty = true;
  return ZKB_OK;
}

int32_t zkb_populate_heap(ZkbBatch* b, uint32_t vm_lo, uint32_t vm_hi, const uint8_t* bytes, uint32_t n_bytes, uint32_t per_vm) {
  if (!range_ok(b, vm_lo, vm_hi) || (!bytes && n_bytes)) return ZKB_ERR_INVALID_ARGUMENT;
  if (n_bytes > b->cfg.heap_bytes) return set_err(ZKB_ERR_INVALID_ARGUMENT, "populate_heap: image larger than ZkbConfig.heap_bytes");
  if (n_bytes == 0 || vm_lo == vm_hi) return ZKB_OK;
  CUDA_OK(cudaSetDevice(b->cfg.device));
  if (download_hot(b) != ZKB_OK) return ZKB_ERR_CUDA;
  for (uint32_t v = vm_lo; v < vm_hi; v++) {
    VmHot& h = b->h_hot[v];
    if (h.x[X_FAR_DEPTH] != 1) return set_err(ZKB_ERR_INVALID_ARGUMENT, "populate_heap: push the bootloader context first");
    h.x[X_SLAB_FREE] &= ~1u;  // slab 0 = heap of the bootloader frame
  }
  b->hot_dirty = true;
  size_t total = per_vm ? (size_t)(vm_hi - vm_lo) * n_bytes : n_bytes;
  uint8_t* d_bytes = nullptr;
  int32_t src = stage_h2d(b, bytes, total, &d_bytes);
  if (src != ZKB_OK) return src;
  zkb_populate_heap_kernel<<<(vm_hi - vm_lo + 3) / 4, 128>>>(b->d, vm_lo, vm_hi, d_bytes, n_bytes, per_vm, 1);
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaStreamSynchronize(0));  // the staging buffer is reused by the next populate_* call
  return ZKB_OK;
}

int32_t zkb_set_register(ZkbBatch* b, uint32_t vm_lo, uint32_t vm_hi, uint32_t reg, const uint8_t* value_be, uint8_t is_pointer, uint32_t per_vm) {
  if (!range_ok(b, vm_lo, vm_hi) || reg >= ZK_REGISTERS_COUNT || !value_be) return ZKB_ERR_INVALID_ARGUMENT;
  if (download_hot(b) != ZKB_OK) return ZKB_ERR_CUDA;
  for (uint32_t v = vm_lo; v < vm_hi; v++) {
    VmHot& h = b->h_hot[v];
    be32_to_limbs(per_vm ? value_be + (size_t)(v - vm_lo) * 32 : value_be, h.regs[reg + 1]);
    h.x[X_PTRMASK] = (h.x[X_PTRMASK] & ~(1u << (reg + 1))) | (is_pointer ? 1u << (reg + 1) : 0u);
  }
  b->hot_dirty = true;
  return ZKB_OK;
}

int32_t zkb_set_local_field(ZkbBatch* b, uint32_t vm_lo, uint32_t vm_hi, uint32_t field, uint32_t value) {
  if (!range_ok(b, vm_lo, vm_hi)) return ZKB_ERR_INVALID_ARGUMENT;
  if (download_hot(b) != ZKB_OK) return ZKB_ERR_CUDA;
  for (uint32_t v = vm_lo; v < vm_hi; v++) {
    VmHot& h = b->h_hot[v];
    switch (field) {
      case ZKB_FIELD_MEMORY_PAGE_COUNTER: h.live[L_PAGE_COUNTER - 40] = value; break;
      case ZKB_FIELD_ERGS_PER_PUBDATA: h.live[L_EPP - 40] = value; break;
      case ZKB_FIELD_TX_NUMBER: h.live[L_TX_PSP - 40] = (h.live[L_TX_PSP - 40] & 0xFFFF0000u) | (value & 0xFFFFu); break;
      case ZKB_FIELD_TIMESTAMP: h.x[X_TIMESTAMP] = value; break;
      default: return ZKB_ERR_INVALID_ARGUMENT;
    }
  }
  b->hot_dirty = true;
  return ZKB_OK;
}

int32_t zkb_run(ZkbBatch* b, uint32_t max_cycles_per_vm, void* cuda_stream) {
  if (!b) return ZKB_ERR_INVALID_ARGUMENT;
  CUDA_OK(cudaSetDevice(b->cfg.device));
  int32_t rc = upload(b);
  if (rc != ZKB_OK) return rc;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  b->last_stream = st;
  if (st != nullptr) {  // populate_* / reset work runs on the default stream: order it before the launch on `st`
    CUDA_OK(cudaEventRecord(b->ev_setup, 0));
    CUDA_OK(cudaStreamWaitEvent(st, b->ev_setup, 0));
  }
  // Regroup the schedule by bootloader code: a warp runs four VMs and a lockstep CTA 96 side by side, at full speed only
  // while they execute the same program.  A batch whose neighbours run different programs (a block of unrelated
  // transactions) is therefore scheduled in code order: slot i runs VM order[i] (stable, so equal programs keep their
  // relative order).  State and streams stay indexed by VM: results are unchanged, only the assignment of VMs to octets.
  if (b->order_dirty) {
    b->order_dirty = false;
    const uint32_t n = b->cfg.n_vms;
    uint32_t runs = 0;   // maximal runs of equal code in VM order; grouped already <=> one run per distinct code
    std::vector<uint8_t> seen(b->codes.size() + 1, 0);
    bool grouped = true;
    for (uint32_t v = 0; v < n; v++)
      if (v == 0 || b->boot_code[v] != b->boot_code[v - 1]) {
        runs++;
        uint8_t& s = seen[(size_t)(b->boot_code[v] + 1)];
        if (s) grouped = false;
        s = 1;
      }
    if (b->regroup && !grouped) {
      std::vector<uint32_t> order(n);
      for (uint32_t v = 0; v < n; v++) order[v] = v;
      std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return b->boot_code[x] < b->boot_code[y]; });
      if (!b->d_order) {
        cudaError_t e = dalloc(b, &b->d_order, n, false);
        if (e != cudaSuccess) return set_err(ZKB_ERR_OUT_OF_MEMORY, std::string("schedule order: ") + cudaGetErrorString(e));
      }
      CUDA_OK(cudaMemcpyAsync(b->d_order, order.data(), (size_t)n * 4, cudaMemcpyHostToDevice, st));
      CUDA_OK(cudaStreamSynchronize(st));   // `order` is a pageable temporary
      b->d.order = b->d_order;
    } else {
      b->d.order = nullptr;
    }
    (void)runs;
  }
  CUDA_OK(cudaMemsetAsync(b->d.queue, 0, 8, st));
  CUDA_OK(cudaEventRecord(b->ev0, st));
  // Lockstep schedule: a CTA pulls `chunk` VMs per turn.  With chunk = VMs per CTA a batch of 65 536 VMs is 4.61 waves of
  // 148 x 96 VMs and the fifth wave leaves 57 SMs idle; chunk = ceil(n_vms / (waves x CTAs)) spreads the same number of
  // waves evenly (89 VMs per turn: 4.98 waves).  ZKB_CHUNK: experiment override (0 / unset = balanced).
  {
    uint32_t chunk = ZKB_VMS_PER_CTA;
    const uint64_t per_wave = (uint64_t)b->grid * ZKB_VMS_PER_CTA;
    const uint64_t waves = std::max<uint64_t>(1, (b->cfg.n_vms + per_wave - 1) / per_wave);
    if (b->balance_waves) chunk = (uint32_t)((b->cfg.n_vms + waves * b->grid - 1) / (waves * b->grid));
    if (const char* env = getenv("ZKB_CHUNK")) {
      const int v = atoi(env);
      if (v > 0) chunk = (uint32_t)v;
    }
    b->d.chunk = std::min<uint32_t>(std::max<uint32_t>(chunk, 1u), ZKB_VMS_PER_CTA);
  }
  int grid = std::min<int>(b->grid, (int)((b->cfg.n_vms + b->d.chunk - 1) / b->d.chunk));
  // FAST kernel (no out-of-line precompile routine anywhere in it), then the FULL kernel for the VMs the fast one parked
  // with a pending ecrecover / long keccak256 (none on most batches: its CTAs then find nothing to do and exit)
  if (b->lockstep) {
    zkb_run_kernel<true, false><<<grid, ZKB_WARPS_PER_CTA * 32, ZKB_RUN_SMEM_BYTES, st>>>(b->d, max_cycles_per_vm);
    zkb_run_kernel<true, true><<<grid, ZKB_WARPS_PER_CTA * 32, ZKB_RUN_SMEM_BYTES, st>>>(b->d, max_cycles_per_vm);
  } else {
    zkb_run_kernel<false, false><<<grid, ZKB_WARPS_PER_CTA * 32, ZKB_RUN_SMEM_BYTES, st>>>(b->d, max_cycles_per_vm);
    zkb_run_kernel<false, true><<<grid, ZKB_WARPS_PER_CTA * 32, ZKB_RUN_SMEM_BYTES, st>>>(b->d, max_cycles_per_vm);
  }
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaEventRecord(b->ev1, st));
  b->n_launches = 2;
  b->hot_stale = true;
  b->launched = true;
  b->offsets_valid = false;
  b->flat_valid = false;
  b->consume_valid = false;
  return ZKB_OK;
}

int32_t zkb_sync(ZkbBatch* b) {
  if (!b) return ZKB_ERR_INVALID_ARGUMENT;
  CUDA_OK(cudaSetDevice(b->cfg.device));
  return wait_last_run(b);
}

int32_t zkb_last_run_ms(ZkbBatch* b, float* ms, uint32_t* n_kernel_launches) {
  if (!b) return ZKB_ERR_INVALID_ARGUMENT;
  CUDA_OK(cudaEventSynchronize(b->ev1));
  float t = 0;
  CUDA_OK(cudaEventElapsedTime(&t, b->ev0, b->ev1));
  if (ms) *ms = t;
  if (n_kernel_launches) *n_kernel_launches = b->n_launches;
  return ZKB_OK;
}

int32_t zkb_vm_status(ZkbBatch* b, uint32_t vm_lo, uint32_t vm_hi, ZkbVmStatus* out) {
  if (!range_ok(b, vm_lo, vm_hi) || !out) return ZKB_ERR_INVALID_ARGUMENT;
  const uint32_t* c = nullptr;
  int32_t rc = summary(b, &c);
  if (rc != ZKB_OK) return rc;
  for (uint32_t v = vm_lo; v < vm_hi; v++) {
    out[v - vm_lo].code = c[(size_t)v * 8 + 6];
    out[v - vm_lo].cycles = c[(size_t)v * 8 + 7];
  }
  return ZKB_OK;
}

int32_t zkb_read_local_state(ZkbBatch* b, uint32_t vm, ZkbLocalState* out) {
  if (!b || vm >= b->cfg.n_vms || !out) return ZKB_ERR_INVALID_ARGUMENT;
  int32_t rc = download_hot(b);
  if (rc != ZKB_OK) return rc;
  const VmHot& h = b->h_hot[vm];
  memset(out, 0, sizeof(*out));
  memcpy(out->previous_code_word, h.prev_word, 32);
  out->previous_code_memory_page = h.x[X_PREV_CODE_PAGE];
  for (int i = 0; i < 15; i++) memcpy(out->registers[i], h.regs[i + 1], 32);
  out->register_is_pointer = (uint16_t)(h.x[X_PTRMASK] >> 1);
  out->flags = (uint8_t)h.x[X_FLAGS];
  out->pending_exception = (uint8_t)h.x[X_PENDING];
  out->timestamp = h.x[X_TIMESTAMP];
  out->monotonic_cycle_counter = h.x[X_CYCLE];
  out->spent_pubdata_counter = h.live[L_SPENT_PUBDATA - 40];
  out->memory_page_counter = h.live[L_PAGE_COUNTER - 40];
  out->absolute_execution_step = 0;
  out->current_ergs_per_pubdata_byte = h.live[L_EPP - 40];
  out->tx_number_in_block = (uint16_t)(h.live[L_TX_PSP - 40] & 0xFFFFu);
  out->previous_super_pc = (uint16_t)(h.live[L_TX_PSP - 40] >> 16);
  memcpy(out->context_u128_register, &h.live[L_CTX - 40], 16);
  out->callstack_depth = h.live[L_DEPTH - 40];
  ZkbFrame& f = out->current_frame;
  const uint32_t* F = h.F;
  memcpy(f.this_address, &F[F_THIS], 20);
  memcpy(f.msg_sender, &F[F_SENDER], 20);
  memcpy(f.code_address, &F[F_CODE_ADDR], 20);
  f.base_memory_page = F[F_BASE_PAGE];
  f.code_page = F[F_CODE_PAGE];
  f.sp = (uint16_t)(F[F_SP_PC] & 0xFFFFu);
  f.pc = (uint16_t)(F[F_SP_PC] >> 16);
  f.exception_handler_location = (uint16_t)(F[F_EH_SHARDS] & 0xFFFFu);
  f.ergs_remaining = F[F_ERGS];
  f.this_shard_id = (uint8_t)(F[F_EH_SHARDS] >> 16);
  f.caller_shard_id = (uint8_t)(F[F_EH_SHARDS] >> 24);
  f.code_shard_id = (uint8_t)(F[F_MISC] & 0xFFu);
  f.is_static = (uint8_t)((F[F_MISC] >> 8) & 1u);
  f.is_local_frame = (uint8_t)((F[F_MISC] >> 16) & 1u);
  memcpy(f.context_u128_value, &F[F_CTX], 16);
  f.heap_bound = F[F_HEAP_BOUND];
  f.aux_heap_bound = F[F_AUX_BOUND];
  return ZKB_OK;
}

int32_t zkb_stream_counts(ZkbBatch* b, uint32_t kind, uint32_t vm_lo, uint32_t vm_hi, uint32_t* counts_out) {
  if (!range_ok(b, vm_lo, vm_hi) || kind >= ZKB_N_STREAMS || !counts_out) return ZKB_ERR_INVALID_ARGUMENT;
  const uint32_t* c = nullptr;
  int32_t rc = summary(b, &c);
  if (rc != ZKB_OK) return rc;
  for (uint32_t v = vm_lo; v < vm_hi; v++) counts_out[v - vm_lo] = c[(size_t)v * 8 + kind];
  return ZKB_OK;
}

int32_t zkb_totals(ZkbBatch* b, uint64_t* total_cycles, uint64_t stream_bytes[ZKB_N_STREAMS]) {
  if (!b) return ZKB_ERR_INVALID_ARGUMENT;
  const uint32_t* c = nullptr;
  int32_t rc = summary(b, &c);
  if (rc != ZKB_OK) return rc;
  uint64_t cyc = 0, sb[ZKB_N_STREAMS] = {0, 0, 0, 0, 0, 0};
  for (size_t v = 0; v < b->cfg.n_vms; v++) {
    cyc += c[v * 8 + 7];
    for (int k = 0; k < ZKB_N_STREAMS; k++) sb[k] += (uint64_t)c[v * 8 + k] * REC_BYTES[k];
  }
  if (total_cycles) *total_cycles = cyc;
  if (stream_bytes) memcpy(stream_bytes, sb, sizeof(sb));
  return ZKB_OK;
}

int32_t zkb_stream_device_view(ZkbBatch* b, uint32_t kind, void** dptr, uint64_t* stride_bytes) {
  if (!b || kind >= ZKB_N_STREAMS) return ZKB_ERR_INVALID_ARGUMENT;
  if (dptr) *dptr = b->d.streams[kind];
  if (stride_bytes) *stride_bytes = (uint64_t)b->cfg.cap_records[kind] * REC_BYTES[kind];
  return ZKB_OK;
}

int32_t zkb_read_stream(ZkbBatch* b, uint32_t vm, uint32_t kind, void* dst, uint64_t max_bytes, uint64_t* n_bytes) {
  if (!b || vm >= b->cfg.n_vms || kind >= ZKB_N_STREAMS) return ZKB_ERR_INVALID_ARGUMENT;
  const uint32_t* c = nullptr;
  int32_t rc = summary(b, &c);
  if (rc != ZKB_OK) return rc;
  uint64_t n = b->cfg.witness_mode ? (uint64_t)c[(size_t)vm * 8 + kind] * REC_BYTES[kind] : 0;
  if (n_bytes) *n_bytes = n;
  uint64_t take = std::min(n, max_bytes);
  if (dst && take) {
    CUDA_OK(cudaSetDevice(b->cfg.device));
    CUDA_OK(cudaMemcpy(dst, b->d.streams[kind] + (size_t)vm * b->cfg.cap_records[kind] * REC_BYTES[kind], take, cudaMemcpyDeviceToHost));
  }
  return ZKB_OK;
}

// byte offsets of every VM in every packed stream, computed once per finished run from the host-visible summary
static void refresh_offsets(ZkbBatch* b) {
  if (b->offsets_valid) return;
  uint32_t n = b->cfg.n_vms;
  // ONE pass over the per-VM summary (8 words per VM, the six counts side by side) for all six prefix sums
  uint64_t run[ZKB_N_STREAMS];
  for (uint32_t kind = 0; kind < ZKB_N_STREAMS; kind++) b->h_offsets[kind][0] = run[kind] = 0;
  for (uint32_t v = 0; v < n; v++) {
    const uint32_t* c = b->h_counts + (size_t)v * 8;
    for (uint32_t kind = 0; kind < ZKB_N_STREAMS; kind++) {
      run[kind] += (uint64_t)c[kind] * REC_BYTES[kind];
      b->h_offsets[kind][v + 1] = run[kind];
    }
  }
  b->offsets_valid = true;
}

static int32_t ensure_pack_capacity(ZkbBatch* b) {
  // one buffer per stream kind, laid out back to back, sized for the current counts (grow-only)
  refresh_offsets(b);
  uint64_t need = 0;
  for (uint32_t k = 0; k < ZKB_N_STREAMS; k++) need += (b->h_offsets[k][b->cfg.n_vms] + 255) / 256 * 256;
  if (need > b->pack_capacity) {
    CUDA_OK(cudaDeviceSynchronize());  // a previous async fetch may still read the old buffer
    if (b->d_pack) CUDA_OK(cudaFree(b->d_pack));
    b->d_pack = nullptr;
    b->pack_capacity = 0;
    uint64_t cap = need + need / 16;
    CUDA_OK(cudaMalloc(&b->d_pack, cap));
    b->pack_capacity = cap;
  }
  return ZKB_OK;
}

static uint8_t* pack_region(ZkbBatch* b, uint32_t kind) {
  uint64_t at = 0;
  for (uint32_t k = 0; k < kind; k++) at += (b->h_offsets[k][b->cfg.n_vms] + 255) / 256 * 256;
  return b->d_pack + at;
}

static int32_t pack_async(ZkbBatch* b, uint32_t kind, cudaStream_t st, uint8_t** dptr, uint64_t* n_bytes) {
  CUDA_OK(cudaSetDevice(b->cfg.device));
  const uint32_t* c = nullptr;
  int32_t rc = summary(b, &c);  // waits for THIS batch's run only; no device->host copy
  if (rc != ZKB_OK) return rc;
  rc = ensure_pack_capacity(b);
  if (rc != ZKB_OK) return rc;
  uint32_t n = b->cfg.n_vms;
  uint64_t total = b->h_offsets[kind][n];
  uint64_t* d_off = b->d_offsets + (size_t)kind * (n + 1);
  uint8_t* dst = pack_region(b, kind);
  CUDA_OK(cudaMemcpyAsync(d_off, b->h_offsets[kind], (n + 1) * 8, cudaMemcpyHostToDevice, st));
  if (total) {
    zkb_pack_kernel<<<n, 128, 0, st>>>(b->d.streams[kind], (uint64_t)b->cfg.cap_records[kind] * REC_BYTES[kind], d_off, dst, n);
    CUDA_OK(cudaGetLastError());
  }
  *dptr = dst;
  *n_bytes = total;
  return ZKB_OK;
}

int32_t zkb_pack_stream_device(ZkbBatch* b, uint32_t kind, void** dptr, uint64_t* n_bytes, void* cuda_stream) {
  if (!b || kind >= ZKB_N_STREAMS || !b->cfg.witness_mode) return ZKB_ERR_INVALID_ARGUMENT;
  uint8_t* p = nullptr;
  uint64_t total = 0;
  int32_t rc = pack_async(b, kind, (cudaStream_t)cuda_stream, &p, &total);
  if (rc != ZKB_OK) return rc;
  CUDA_OK(cudaStreamSynchronize((cudaStream_t)cuda_stream));
  if (dptr) *dptr = p;
  if (n_bytes) *n_bytes = total;
  return ZKB_OK;
}

int32_t zkb_pack_stream_device_async(ZkbBatch* b, uint32_t kind, void** dptr, uint64_t* n_bytes, void* cuda_stream) {
  if (!b || kind >= ZKB_N_STREAMS || !b->cfg.witness_mode) return ZKB_ERR_INVALID_ARGUMENT;
  uint8_t* p = nullptr;
  uint64_t total = 0;
  int32_t rc = pack_async(b, kind, (cudaStream_t)cuda_stream, &p, &total);
  if (rc != ZKB_OK) return rc;
  if (dptr) *dptr = p;
  if (n_bytes) *n_bytes = total;
  return ZKB_OK;
}

int32_t zkb_fetch_stream_packed_async(ZkbBatch* b, uint32_t kind, void* host_dst, uint64_t host_capacity, uint64_t* offsets_out,
                                      void* cuda_stream) {
  if (!b || kind >= ZKB_N_STREAMS || !b->cfg.witness_mode) return ZKB_ERR_INVALID_ARGUMENT;
  uint8_t* p = nullptr;
  uint64_t total = 0;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  int32_t rc = pack_async(b, kind, st, &p, &total);
  if (rc != ZKB_OK) return rc;
  if (offsets_out) memcpy(offsets_out, b->h_offsets[kind], ((size_t)b->cfg.n_vms + 1) * 8);
  if (total > host_capacity) return set_err(ZKB_ERR_INVALID_ARGUMENT, "fetch_stream_packed: host buffer too small");
  if (total && host_dst) {
    CUDA_OK(cudaMemcpyAsync(host_dst, p, total, cudaMemcpyDeviceToHost, st));
    b->d2h_bytes += total;
  }
  return ZKB_OK;
}

int32_t zkb_fetch_stream_packed(ZkbBatch* b, uint32_t kind, void* host_dst, uint64_t host_capacity, uint64_t* offsets_out) {
  int32_t rc = zkb_fetch_stream_packed_async(b, kind, host_dst, host_capacity, offsets_out, nullptr);
  if (rc != ZKB_OK) return rc;
  CUDA_OK(cudaStreamSynchronize(nullptr));
  return ZKB_OK;
}

// ---- transport encoding (include/zkb_codec.h) ------------------------------------------------------------------
// encoding pass (payloads into a staging area at worst-case offsets, true sizes) + scan on the encoder's own stream, then
// (after the host has read the totals from mapped memory and sized the blob) the compaction into the blob.  Returns the
// device blob; it stays valid until the next encode / destroy.
static int32_t encode_async(ZkbBatch* b, cudaStream_t user_stream, uint8_t** dptr, uint64_t* n_bytes, uint32_t kinds_mask = 63u) {
  CUDA_OK(cudaSetDevice(b->cfg.device));
  if (kinds_mask & 3u) kinds_mask |= 3u;   // cycle rows and memory queries are coded jointly (format v2): both or neither
  const uint32_t* c = nullptr;
  int32_t rc = summary(b, &c);  // waits for THIS batch's run only
  if (rc != ZKB_OK) return rc;
  refresh_offsets(b);           // canonical (packed) byte offsets of every VM, in pinned host memory
  const size_t n = b->cfg.n_vms;
  if (!b->d_enc_sizes) {
    cudaError_t e = dalloc(b, &b->d_enc_sizes, n * ZKB_N_STREAMS, false);
    if (e == cudaSuccess) e = dalloc(b, &b->d_enc_offsets, (n + 1) * ZKB_N_STREAMS, false);
    if (e == cudaSuccess) e = dalloc(b, &b->d_enc_canon, (n + 1) * ZKB_N_STREAMS, false);
    void* hp = nullptr;
    void* dp = nullptr;
    if (e == cudaSuccess) e = cudaHostAlloc(&hp, 64, cudaHostAllocMapped);
    if (e == cudaSuccess) e = cudaHostGetDevicePointer(&dp, hp, 0);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&b->ev_enc, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&b->ev_enc_done, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&b->ev_blob_read, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&b->enc_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) return set_err(ZKB_ERR_OUT_OF_MEMORY, std::string("encoder buffers: ") + cudaGetErrorString(e));
    b->h_enc_totals = (uint64_t*)hp;
    b->d_enc_totals = (uint64_t*)dp;
  }
  EncArgs a{};
  a.sizes = b->d_enc_sizes;
  a.offsets = b->d_enc_offsets;
  a.totals = b->d_enc_totals;
  a.canon = b->d_enc_canon;
  a.kinds_mask = kinds_mask & 63u;
  // staging area: every stream's region holds (records of the batch) x (longest encoding of one record)
  static const uint32_t WORST[ZKB_N_STREAMS] = {8u + ZKB_ROW_TX_WORDS * 4u, 52u, 132u, 52u, 132u, 8u};
  uint64_t stage_need = 0;
  for (int k = 0; k < ZKB_N_STREAMS; k++) {
    a.stage_base[k] = stage_need;
    stage_need += (b->h_offsets[k][n] / REC_BYTES[k] * WORST[k] + 255) / 256 * 256;
  }
  // The run has finished (summary() waited for it), so the passes need no dependency on the caller's stream -- and they
  // must not sit on it: in a pipelined host loop that stream still carries the previous sub-batch's D2H copy.
  cudaStream_t st = b->enc_stream;
  if (stage_need > b->enc_stage_capacity) {
    CUDA_OK(cudaStreamSynchronize(st));   // (the previous encode of this batch is long finished; its blob copy does not read staging)
    if (b->d_enc_stage) CUDA_OK(cudaFree(b->d_enc_stage));
    b->d_enc_stage = nullptr;
    b->enc_stage_capacity = 0;
    const uint64_t cap = stage_need + stage_need / 8 + 4096;
    cudaError_t e = cudaMalloc(&b->d_enc_stage, cap);
    if (e != cudaSuccess) return set_err(ZKB_ERR_OUT_OF_MEMORY, std::string("encoder staging cudaMalloc: ") + cudaGetErrorString(e));
    b->enc_stage_capacity = cap;
  }
  a.stage = b->d_enc_stage;
  for (int k = 0; k < ZKB_N_STREAMS; k++)
    CUDA_OK(cudaMemcpyAsync(b->d_enc_canon + (size_t)k * (n + 1), b->h_offsets[k], (n + 1) * 8, cudaMemcpyHostToDevice, st));
  int n_sm = 148;
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, b->cfg.device);
  const int grid = (int)std::max<size_t>(1, std::min<size_t>((n + 7) / 8, (size_t)n_sm * 8));
  zkb_encode_kernel<<<grid, 256, 0, st>>>(b->d, a);
  zkb_encode_scan_kernel<<<ZKB_N_STREAMS, 1024, 0, st>>>(b->d, a);
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaEventRecord(b->ev_enc, st));
  CUDA_OK(cudaEventSynchronize(b->ev_enc));
  ZkbEncodedHeader h{};
  h.magic = ZKB_CODEC_MAGIC;
  h.version = ZKB_CODEC_VERSION;
  h.n_vms = (uint32_t)n;
  h.reserved0 = (kinds_mask & 63u) == 63u ? 0u : (kinds_mask & 63u);   // 0 = every stream is in the blob
  h.counts_offset = sizeof(ZkbEncodedHeader);
  h.offsets_offset = h.counts_offset + n * 32;
  uint64_t at = h.offsets_offset + (n + 1) * ZKB_N_STREAMS * 8;
  for (int k = 0; k < ZKB_N_STREAMS; k++) {
    at = (at + 15) / 16 * 16;
    h.payload_offset[k] = a.payload_offset[k] = at;
    h.payload_bytes[k] = ((volatile uint64_t*)b->h_enc_totals)[k];
    at += h.payload_bytes[k];
  }
  h.total_bytes = (at + 15) / 16 * 16;
  for (size_t v = 0; v < n; v++)
    for (int k = 0; k < ZKB_N_STREAMS; k++)
      if ((kinds_mask >> k) & 1u) h.raw_bytes += (uint64_t)c[v * 8 + k] * REC_BYTES[k];
  if (h.total_bytes > b->enc_capacity) {
    CUDA_OK(cudaDeviceSynchronize());  // an earlier async fetch may still read the old blob
    if (b->d_enc) CUDA_OK(cudaFree(b->d_enc));
    b->d_enc = nullptr;
    b->enc_capacity = 0;
    const uint64_t cap = h.total_bytes + h.total_bytes / 8 + 4096;
    cudaError_t e = cudaMalloc(&b->d_enc, cap);
    if (e != cudaSuccess) return set_err(ZKB_ERR_OUT_OF_MEMORY, std::string("encoder blob cudaMalloc: ") + cudaGetErrorString(e));
    b->enc_capacity = cap;
  }
  a.blob = b->d_enc;
  a.counts_offset = h.counts_offset;
  a.offsets_offset = h.offsets_offset;
  if (b->blob_read_pending) CUDA_OK(cudaStreamWaitEvent(st, b->ev_blob_read, 0));  // the previous blob is still being copied out
  zkb_encode_header_kernel<<<1, 32, 0, st>>>(h, b->d_enc);
  zkb_encode_compact_kernel<<<(int)std::max<size_t>(1, std::min<size_t>(n, (size_t)n_sm * 16)), 128, 0, st>>>(b->d, a);
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaEventRecord(b->ev_enc_done, st));
  CUDA_OK(cudaStreamWaitEvent(user_stream, b->ev_enc_done, 0));   // whatever the caller queues next sees the finished blob
  *dptr = b->d_enc;
  *n_bytes = h.total_bytes;
  return ZKB_OK;
}

int32_t zkb_encode_streams_device(ZkbBatch* b, void** dptr, uint64_t* n_bytes, void* cuda_stream) {
  if (!b || !b->cfg.witness_mode) return ZKB_ERR_INVALID_ARGUMENT;
  uint8_t* p = nullptr;
  uint64_t total = 0;
  int32_t rc = encode_async(b, (cudaStream_t)cuda_stream, &p, &total);
  if (rc != ZKB_OK) return rc;
  if (dptr) *dptr = p;
  if (n_bytes) *n_bytes = total;
  return ZKB_OK;
}

int32_t zkb_fetch_encoded_async(ZkbBatch* b, void* host_dst, uint64_t host_capacity, uint64_t* n_bytes, void* cuda_stream) {
  if (!b || !b->cfg.witness_mode) return ZKB_ERR_INVALID_ARGUMENT;
  uint8_t* p = nullptr;
  uint64_t total = 0;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  int32_t rc = encode_async(b, st, &p, &total);
  if (rc != ZKB_OK) return rc;
  if (n_bytes) *n_bytes = total;
  if (!host_dst) return ZKB_OK;   // size query
  if (total > host_capacity) return set_err(ZKB_ERR_INVALID_ARGUMENT, "fetch_encoded: host buffer too small");
  CUDA_OK(cudaMemcpyAsync(host_dst, p, total, cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaEventRecord(b->ev_blob_read, st));
  b->blob_read_pending = true;
  b->d2h_bytes += total;
  return ZKB_OK;
}

int32_t zkb_fetch_encoded_kinds_async(ZkbBatch* b, uint32_t kinds_mask, void* host_dst, uint64_t host_capacity, uint64_t* n_bytes, void* cuda_stream) {
  if (!b || !b->cfg.witness_mode || !(kinds_mask & 63u)) return ZKB_ERR_INVALID_ARGUMENT;
  uint8_t* p = nullptr;
  uint64_t total = 0;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  int32_t rc = encode_async(b, st, &p, &total, kinds_mask);
  if (rc != ZKB_OK) return rc;
  if (n_bytes) *n_bytes = total;
  if (!host_dst) return ZKB_OK;
  if (total > host_capacity) return set_err(ZKB_ERR_INVALID_ARGUMENT, "fetch_encoded: host buffer too small");
  CUDA_OK(cudaMemcpyAsync(host_dst, p, total, cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaEventRecord(b->ev_blob_read, st));
  b->blob_read_pending = true;
  b->d2h_bytes += total;
  return ZKB_OK;
}

int32_t zkb_fetch_encoded(ZkbBatch* b, void* host_dst, uint64_t host_capacity, uint64_t* n_bytes) {
  int32_t rc = zkb_fetch_encoded_async(b, host_dst, host_capacity, n_bytes, nullptr);
  if (rc != ZKB_OK) return rc;
  CUDA_OK(cudaStreamSynchronize(nullptr));
  return ZKB_OK;
}

// ---- device-side consumer (consume.cuh) --------------------------------------------------------
int32_t zkb_consume(ZkbBatch* b, uint32_t cycles_per_snapshot, void* cuda_stream) {
  if (!b || !b->cfg.witness_mode || cycles_per_snapshot == 0) return ZKB_ERR_INVALID_ARGUMENT;
  CUDA_OK(cudaSetDevice(b->cfg.device));
  int32_t rc = upload(b);
  if (rc != ZKB_OK) return rc;
  if (wait_last_run(b) != ZKB_OK) return ZKB_ERR_CUDA;
  const size_t n = b->cfg.n_vms;
  const uint32_t max_snaps = b->cfg.cap_records[ZKB_STREAM_ROWS] / cycles_per_snapshot + 2;
  ConsumeOut& o = b->consume;
  if (!o.snaps || o.max_snaps < max_snaps) {
    cudaError_t e = dalloc(b, &o.snaps, n * max_snaps * ZKB_SNAP_WORDS, false);
    if (e == cudaSuccess && !o.n_snaps) e = dalloc(b, &o.n_snaps, n, true);
    if (e == cudaSuccess && !o.finals) e = dalloc(b, &o.finals, n * 24, true);
    if (e == cudaSuccess && !o.fstack) e = dalloc(b, &o.fstack, n * (b->cfg.max_depth + 1), true);
    if (e != cudaSuccess) return set_err(ZKB_ERR_OUT_OF_MEMORY, std::string("zkb_consume cudaMalloc: ") + cudaGetErrorString(e));
    o.max_snaps = max_snaps;
  }
  o.period = cycles_per_snapshot;
  o.hot_init = b->d_hot_init;
  cudaStream_t st = (cudaStream_t)cuda_stream;
  int n_sm = 148;
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, b->cfg.device);
  zkb_consume_state_kernel<<<(unsigned)std::max<size_t>(1, std::min<size_t>((n + 7) / 8, (size_t)n_sm * 8)), 256, 0, st>>>(b->d, o);
  zkb_consume_hash_kernel<<<(unsigned)((n * 3 + 127) / 128), 128, 0, st>>>(b->d, o);
  CUDA_OK(cudaGetLastError());
  b->h_n_snaps.resize(n);
  CUDA_OK(cudaMemcpyAsync(b->h_n_snaps.data(), o.n_snaps, n * 4, cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaStreamSynchronize(st));
  b->consume_valid = true;
  return ZKB_OK;
}

int32_t zkb_snapshot_counts(ZkbBatch* b, uint32_t vm_lo, uint32_t vm_hi, uint32_t* counts_out) {
  if (!range_ok(b, vm_lo, vm_hi) || !counts_out || !b->consume_valid) return set_err(ZKB_ERR_INVALID_ARGUMENT, "zkb_snapshot_counts: call zkb_consume first");
  for (uint32_t v = vm_lo; v < vm_hi; v++) counts_out[v - vm_lo] = b->h_n_snaps[v];
  return ZKB_OK;
}

int32_t zkb_read_snapshots(ZkbBatch* b, uint32_t vm, void* dst, uint64_t max_bytes, uint64_t* n_bytes) {
  if (!b || vm >= b->cfg.n_vms || !b->consume_valid) return set_err(ZKB_ERR_INVALID_ARGUMENT, "zkb_read_snapshots: call zkb_consume first");
  const uint64_t n = (uint64_t)b->h_n_snaps[vm] * sizeof(ZkbSnapshot);
  if (n_bytes) *n_bytes = n;
  const uint64_t take = std::min(n, max_bytes);
  if (dst && take) {
    CUDA_OK(cudaSetDevice(b->cfg.device));
    CUDA_OK(cudaMemcpy(dst, b->consume.snaps + (size_t)vm * b->consume.max_snaps * ZKB_SNAP_WORDS, take, cudaMemcpyDeviceToHost));
  }
  return ZKB_OK;
}

int32_t zkb_read_queue_digests(ZkbBatch* b, uint32_t vm_lo, uint32_t vm_hi, uint8_t* digests_out) {
  if (!range_ok(b, vm_lo, vm_hi) || !digests_out || !b->consume_valid) return set_err(ZKB_ERR_INVALID_ARGUMENT, "zkb_read_queue_digests: call zkb_consume first");
  if (vm_lo == vm_hi) return ZKB_OK;
  CUDA_OK(cudaSetDevice(b->cfg.device));
  std::vector<uint32_t> w((size_t)(vm_hi - vm_lo) * 24);
  CUDA_OK(cudaMemcpy(w.data(), b->consume.finals + (size_t)vm_lo * 24, w.size() * 4, cudaMemcpyDeviceToHost));
  for (size_t i = 0; i < w.size(); i++) {   // sha256 digests are big-endian words
    digests_out[4 * i + 0] = (uint8_t)(w[i] >> 24);
    digests_out[4 * i + 1] = (uint8_t)(w[i] >> 16);
    digests_out[4 * i + 2] = (uint8_t)(w[i] >> 8);
    digests_out[4 * i + 3] = (uint8_t)w[i];
  }
  return ZKB_OK;
}

// all VMs' snapshots packed VM-major (+ the queue digests) into pinned host memory, asynchronously on cuda_stream:
// [n_vms + 1] u64 snapshot offsets (in snapshots), the snapshots, then n_vms x 3 x 32 bytes of final queue digests (as words)
__global__ void zkb_pack_snapshots_kernel(const uint32_t* __restrict__ snaps, uint32_t max_snaps, const uint64_t* __restrict__ offsets, uint32_t n_vms,
                                          uint32_t* __restrict__ out) {
  const uint32_t vm = blockIdx.x;
  if (vm >= n_vms) return;
  const uint64_t lo = offsets[vm], n = (offsets[vm + 1] - lo) * ZKB_SNAP_WORDS;
  const uint32_t* src = snaps + (size_t)vm * max_snaps * ZKB_SNAP_WORDS;
  uint32_t* dst = out + lo * ZKB_SNAP_WORDS;
  for (uint64_t i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
}

int32_t zkb_fetch_consumed_async(ZkbBatch* b, void* host_dst, uint64_t host_capacity, uint64_t* n_bytes, void* cuda_stream) {
  if (!b || !b->consume_valid) return set_err(ZKB_ERR_INVALID_ARGUMENT, "zkb_fetch_consumed_async: call zkb_consume first");
  CUDA_OK(cudaSetDevice(b->cfg.device));
  const size_t n = b->cfg.n_vms;
  std::vector<uint64_t> off(n + 1, 0);
  for (size_t v = 0; v < n; v++) off[v + 1] = off[v] + b->h_n_snaps[v];
  const uint64_t table = (n + 1) * 8, body = off[n] * sizeof(ZkbSnapshot), tail = n * 96;
  const uint64_t total = table + body + tail;
  if (n_bytes) *n_bytes = total;
  if (!host_dst) return ZKB_OK;
  if (total > host_capacity) return set_err(ZKB_ERR_INVALID_ARGUMENT, "zkb_fetch_consumed_async: host buffer too small");
  if (total > b->snap_pack_capacity) {
    CUDA_OK(cudaDeviceSynchronize());
    if (b->d_snap_pack) CUDA_OK(cudaFree(b->d_snap_pack));
    b->d_snap_pack = nullptr;
    CUDA_OK(cudaMalloc(&b->d_snap_pack, total + total / 8));
    b->snap_pack_capacity = total + total / 8;
  }
  cudaStream_t st = (cudaStream_t)cuda_stream;
  memcpy(host_dst, off.data(), table);   // the table is host knowledge already
  CUDA_OK(cudaMemcpyAsync(b->d_snap_pack, host_dst, table, cudaMemcpyHostToDevice, st));
  zkb_pack_snapshots_kernel<<<(unsigned)n, 128, 0, st>>>(b->consume.snaps, b->consume.max_snaps, (const uint64_t*)b->d_snap_pack, (uint32_t)n,
                                                        (uint32_t*)(b->d_snap_pack + table));
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaMemcpyAsync(b->d_snap_pack + table + body, b->consume.finals, tail, cudaMemcpyDeviceToDevice, st));
  CUDA_OK(cudaMemcpyAsync((uint8_t*)host_dst + table, b->d_snap_pack + table, body + tail, cudaMemcpyDeviceToHost, st));
  b->d2h_bytes += body + tail;
  return ZKB_OK;
}

int32_t zkb_flatten_logs(ZkbBatch* b, void* cuda_stream) {
  if (!b || !b->cfg.witness_mode) return ZKB_ERR_INVALID_ARGUMENT;
  CUDA_OK(cudaSetDevice(b->cfg.device));
  int32_t rc = upload(b);
  if (rc != ZKB_OK) return rc;
  if (wait_last_run(b) != ZKB_OK) return ZKB_ERR_CUDA;
  size_t n = b->cfg.n_vms;
  if (!b->flat_allocated) {
    FlatOut& f = b->flat;
    f.cap_net = std::max(1u, b->cfg.cap_records[ZKB_STREAM_LOG]);
    f.cap_hist = 2 * f.cap_net;  // every write / event can be followed by at most one rollback query
    cudaError_t e = cudaSuccess;
#define ALLOC(ptr, count) if (e == cudaSuccess) e = dalloc(b, &(ptr), (count), false);
    ALLOC(f.hist[0], n * f.cap_hist * 32);
    ALLOC(f.hist[1], n * f.cap_hist * 32);
    ALLOC(f.net[0], n * f.cap_net * 32);
    ALLOC(f.net[1], n * f.cap_net * 32);
    ALLOC(f.rb[0], n * f.cap_net);
    ALLOC(f.rb[1], n * f.cap_net);
    ALLOC(f.counts, n * 4);
    ALLOC(f.status, n);
#undef ALLOC
    if (e != cudaSuccess) return set_err(ZKB_ERR_OUT_OF_MEMORY, std::string("flatten cudaMalloc: ") + cudaGetErrorString(e));
    b->flat_allocated = true;
  }
  cudaStream_t st = (cudaStream_t)cuda_stream;
  int grid = std::max(1, std::min<int>((int)((n + 3) / 4), 148 * 16));
  zkb_flatten_kernel<<<grid, 128, 0, st>>>(b->d, b->flat);
  CUDA_OK(cudaGetLastError());
  b->h_flat_counts.resize(n * 4);
  b->h_flat_status.resize(n);
  CUDA_OK(cudaMemcpyAsync(b->h_flat_counts.data(), b->flat.counts, n * 16, cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaMemcpyAsync(b->h_flat_status.data(), b->flat.status, n * 4, cudaMemcpyDeviceToHost, st));
  CUDA_OK(cudaStreamSynchronize(st));
  b->flat_valid = true;
  return ZKB_OK;
}

// K11 driver: stable radix sort of n LogQueryRec by (group, slot hash) + gather + boundary flags (logsort.cuh)
static int32_t logsort_device(int32_t device, const uint32_t* d_recs, uint64_t n, const uint32_t* d_group_of, uint32_t* d_out, uint8_t* d_boundary,
                              uint64_t* n_groups_out, uint32_t key_bits, cudaStream_t st) {
  CUDA_OK(cudaSetDevice(device));
  if (n_groups_out) *n_groups_out = 0;
  if (n == 0) return ZKB_OK;
  if (n >= (1ull << 32)) return set_err(ZKB_ERR_INVALID_ARGUMENT, "zkb_sort_log_queries: more than 2^32 records");
  const uint32_t n_tiles = (uint32_t)((n + ZKB_SORT_TILE - 1) / ZKB_SORT_TILE);
  uint64_t *keys[2] = {nullptr, nullptr};
  uint32_t *idx[2] = {nullptr, nullptr}, *hist = nullptr;
  unsigned long long* d_ng = nullptr;
  cudaError_t e = cudaMalloc(&keys[0], n * 8);
  if (e == cudaSuccess) e = cudaMalloc(&keys[1], n * 8);
  if (e == cudaSuccess) e = cudaMalloc(&idx[0], n * 4);
  if (e == cudaSuccess) e = cudaMalloc(&idx[1], n * 4);
  if (e == cudaSuccess) e = cudaMalloc(&hist, (size_t)256 * n_tiles * 4);
  if (e == cudaSuccess) e = cudaMalloc(&d_ng, 8);
  auto release = [&]() {
    cudaFree(keys[0]); cudaFree(keys[1]); cudaFree(idx[0]); cudaFree(idx[1]); cudaFree(hist); cudaFree(d_ng);
  };
  if (e != cudaSuccess) {
    release();
    return set_err(ZKB_ERR_OUT_OF_MEMORY, std::string("zkb_sort_log_queries cudaMalloc: ") + cudaGetErrorString(e));
  }
  zkb_logsort_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_recs, n, d_group_of, keys[0], idx[0]);
  int cur = 0;
  for (uint32_t shift = 0; shift < key_bits; shift += 8) {
    zkb_logsort_hist_kernel<<<n_tiles, ZKB_SORT_THREADS, 0, st>>>(keys[cur], n, shift, hist, n_tiles);
    zkb_logsort_scan_kernel<<<1, 1024, 0, st>>>(hist, 256 * n_tiles, n, nullptr);
    zkb_logsort_scatter_kernel<<<n_tiles, ZKB_SORT_THREADS, 0, st>>>(keys[cur], idx[cur], n, shift, hist, n_tiles, keys[cur ^ 1], idx[cur ^ 1]);
    cur ^= 1;
  }
  cudaMemsetAsync(d_ng, 0, 8, st);
  zkb_logsort_gather_kernel<<<(unsigned)((n * 32 + 255) / 256), 256, 0, st>>>(d_recs, keys[cur], idx[cur], n, d_out, d_boundary, d_ng);
  unsigned long long ng = 0;
  e = cudaMemcpyAsync(&ng, d_ng, 8, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e == cudaSuccess) e = cudaGetLastError();
  release();
  if (e != cudaSuccess) return set_err(ZKB_ERR_CUDA, std::string("zkb_sort_log_queries: ") + cudaGetErrorString(e));
  if (n_groups_out) *n_groups_out = ng;
  return ZKB_OK;
}

int32_t zkb_sort_log_queries(int32_t device, const void* d_recs, uint64_t n_records, const uint32_t* d_group_of, void* d_sorted_out,
                             uint8_t* d_boundary_out, uint64_t* n_groups_out, void* cuda_stream) {
  if (n_records && (!d_recs || !d_sorted_out || !d_boundary_out)) return ZKB_ERR_INVALID_ARGUMENT;
  return logsort_device(device, (const uint32_t*)d_recs, n_records, d_group_of, (uint32_t*)d_sorted_out, d_boundary_out, n_groups_out,
                        d_group_of ? 64u : 48u, (cudaStream_t)cuda_stream);
}

// per-VM expansion of "record i belongs to VM v" for the flattened storage histories (VM-major packed)
__global__ void zkb_fill_groups_kernel(const uint64_t* __restrict__ offsets, uint32_t n_vms, uint32_t* __restrict__ group_of) {
  const uint32_t vm = blockIdx.x;
  if (vm >= n_vms) return;
  for (uint64_t i = offsets[vm] + threadIdx.x; i < offsets[vm + 1]; i += blockDim.x) group_of[i] = vm;
}
__global__ void zkb_pack_hist_kernel(const uint32_t* __restrict__ hist, uint32_t cap_hist, const uint64_t* __restrict__ offsets, uint32_t n_vms,
                                     uint32_t* __restrict__ out) {
  const uint32_t vm = blockIdx.x;
  if (vm >= n_vms) return;
  const uint64_t lo = offsets[vm], n = offsets[vm + 1] - lo;
  const uint4* src = reinterpret_cast<const uint4*>(hist + (size_t)vm * cap_hist * 32);
  uint4* dst = reinterpret_cast<uint4*>(out + lo * 32);
  for (uint64_t i = threadIdx.x; i < n * 8; i += blockDim.x) dst[i] = src[i];
}

int32_t zkb_net_storage_history(ZkbBatch* b, void* host_sorted_out, uint64_t host_capacity, uint8_t* host_boundary_out, uint64_t* offsets_out,
                                uint64_t* n_slots_out, void* cuda_stream) {
  if (!b || !offsets_out) return ZKB_ERR_INVALID_ARGUMENT;
  if (!b->flat_valid) {
    int32_t rc = zkb_flatten_logs(b, cuda_stream);
    if (rc != ZKB_OK) return rc;
  }
  CUDA_OK(cudaSetDevice(b->cfg.device));
  cudaStream_t st = (cudaStream_t)cuda_stream;
  const size_t n = b->cfg.n_vms;
  offsets_out[0] = 0;
  for (size_t v = 0; v < n; v++) offsets_out[v + 1] = offsets_out[v] + b->h_flat_counts[v * 4 + 0];   // records, not bytes
  const uint64_t total = offsets_out[n];
  if (n_slots_out) *n_slots_out = 0;
  if (!host_sorted_out || total == 0) return ZKB_OK;   // size query
  if (total * ZKB_LOG_BYTES > host_capacity || !host_boundary_out) return set_err(ZKB_ERR_INVALID_ARGUMENT, "zkb_net_storage_history: host buffers too small");
  if (n >= (1u << 20)) return set_err(ZKB_ERR_INVALID_ARGUMENT, "zkb_net_storage_history: more than 2^20 VMs per batch");
  uint64_t* d_off = nullptr;
  uint32_t *d_packed = nullptr, *d_sorted = nullptr, *d_group = nullptr;
  uint8_t* d_bound = nullptr;
  cudaError_t e = cudaMalloc(&d_off, (n + 1) * 8);
  if (e == cudaSuccess) e = cudaMalloc(&d_packed, total * ZKB_LOG_BYTES);
  if (e == cudaSuccess) e = cudaMalloc(&d_sorted, total * ZKB_LOG_BYTES);
  if (e == cudaSuccess) e = cudaMalloc(&d_group, total * 4);
  if (e == cudaSuccess) e = cudaMalloc(&d_bound, total);
  auto release = [&]() {
    cudaFree(d_off); cudaFree(d_packed); cudaFree(d_sorted); cudaFree(d_group); cudaFree(d_bound);
  };
  if (e != cudaSuccess) {
    release();
    return set_err(ZKB_ERR_OUT_OF_MEMORY, std::string("zkb_net_storage_history cudaMalloc: ") + cudaGetErrorString(e));
  }
  e = cudaMemcpyAsync(d_off, offsets_out, (n + 1) * 8, cudaMemcpyHostToDevice, st);
  zkb_pack_hist_kernel<<<(unsigned)n, 128, 0, st>>>(b->flat.hist[0], b->flat.cap_hist, d_off, (uint32_t)n, d_packed);
  zkb_fill_groups_kernel<<<(unsigned)n, 64, 0, st>>>(d_off, (uint32_t)n, d_group);
  uint64_t n_slots = 0;
  int32_t rc = logsort_device(b->cfg.device, d_packed, total, d_group, d_sorted, d_bound, &n_slots, 64u, st);
  if (rc == ZKB_OK) {
    e = cudaMemcpyAsync(host_sorted_out, d_sorted, total * ZKB_LOG_BYTES, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(host_boundary_out, d_bound, total, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) rc = set_err(ZKB_ERR_CUDA, cudaGetErrorString(e));
    b->d2h_bytes += total * (ZKB_LOG_BYTES + 1);
  }
  release();
  if (n_slots_out) *n_slots_out = n_slots;
  return rc;
}

int32_t zkb_flat_counts(ZkbBatch* b, uint32_t kind, uint32_t vm_lo, uint32_t vm_hi, uint32_t* counts_out, uint32_t* status_out) {
  if (!range_ok(b, vm_lo, vm_hi) || kind >= 4 || !b->flat_valid) return set_err(ZKB_ERR_INVALID_ARGUMENT, "zkb_flat_counts: call zkb_flatten_logs first");
  for (uint32_t v = vm_lo; v < vm_hi; v++) {
    if (counts_out) counts_out[v - vm_lo] = b->h_flat_counts[(size_t)v * 4 + kind];
    if (status_out) status_out[v - vm_lo] = b->h_flat_status[v];
  }
  return ZKB_OK;
}

int32_t zkb_read_flat(ZkbBatch* b, uint32_t vm, uint32_t kind, void* dst, uint64_t max_bytes, uint64_t* n_bytes) {
  if (!b || vm >= b->cfg.n_vms || kind >= 4 || !b->flat_valid) return set_err(ZKB_ERR_INVALID_ARGUMENT, "zkb_read_flat: call zkb_flatten_logs first");
  uint64_t n = (uint64_t)b->h_flat_counts[(size_t)vm * 4 + kind] * ZKB_LOG_BYTES;
  if (n_bytes) *n_bytes = n;
  uint64_t take = std::min(n, max_bytes);
  if (dst && take) {
    CUDA_OK(cudaSetDevice(b->cfg.device));
    const FlatOut& f = b->flat;
    const uint32_t* src = kind < 2 ? f.hist[kind] + (size_t)vm * f.cap_hist * 32 : f.net[kind - 2] + (size_t)vm * f.cap_net * 32;
    CUDA_OK(cudaMemcpy(dst, src, take, cudaMemcpyDeviceToHost));
  }
  return ZKB_OK;
}

int32_t zkb_read_storage(ZkbBatch* b, uint32_t vm, uint8_t shard_id, const uint8_t address[20], const uint8_t key_be[32], uint8_t value_be_out[32]) {
  if (!b || vm >= b->cfg.n_vms || !address || !key_be || !value_be_out) return ZKB_ERR_INVALID_ARGUMENT;
  CUDA_OK(cudaSetDevice(b->cfg.device));
  uint32_t slots = b->cfg.storage_slots;
  std::vector<uint32_t> tags(slots), keys((size_t)slots * 8), addrs((size_t)slots * 8), vals((size_t)slots * 8);
  CUDA_OK(cudaMemcpy(tags.data(), b->d.st_tags + (size_t)vm * slots, slots * 4, cudaMemcpyDeviceToHost));
  CUDA_OK(cudaMemcpy(keys.data(), b->d.st_keys + (size_t)vm * slots * 8, (size_t)slots * 32, cudaMemcpyDeviceToHost));
  CUDA_OK(cudaMemcpy(addrs.data(), b->d.st_addr + (size_t)vm * slots * 8, (size_t)slots * 32, cudaMemcpyDeviceToHost));
  CUDA_OK(cudaMemcpy(vals.data(), b->d.st_vals + (size_t)vm * slots * 8, (size_t)slots * 32, cudaMemcpyDeviceToHost));
  uint32_t key[8], aw[5];
  be32_to_limbs(key_be, key);
  memcpy(aw, address, 20);
  memset(value_be_out, 0, 32);
  for (uint32_t s = 0; s < slots; s++) {
    if (!tags[s]) continue;
    if (memcmp(&keys[(size_t)s * 8], key, 32) == 0 && memcmp(&addrs[(size_t)s * 8], aw, 20) == 0 && addrs[(size_t)s * 8 + 5] == shard_id) {
      limbs_to_be32(&vals[(size_t)s * 8], value_be_out);
      break;
    }
  }
  return ZKB_OK;
}

int32_t zkb_read_heap(ZkbBatch* b, uint32_t vm, uint32_t byte_offset, uint32_t n_bytes, uint8_t* out) {
  if (!b || vm >= b->cfg.n_vms || (!out && n_bytes)) return ZKB_ERR_INVALID_ARGUMENT;
  CUDA_OK(cudaSetDevice(b->cfg.device));
  int32_t rc = download_hot(b);
  if (rc != ZKB_OK) return rc;
  uint32_t level = b->h_hot[vm].x[X_FAR_DEPTH];
  uint32_t lv[4];
  CUDA_OK(cudaMemcpy(lv, b->d.lvl + ((size_t)vm * (b->cfg.max_far_depth + 1) + level) * 4, 16, cudaMemcpyDeviceToHost));
  memset(out, 0, n_bytes);
  if (lv[0] == ZKB_NO_SLAB) return ZKB_OK;
  uint32_t hw = b->cfg.heap_bytes / 32;
  std::vector<uint32_t> slab((size_t)hw * 8);
  CUDA_OK(cudaMemcpy(slab.data(), b->d.heap_mem + ((size_t)vm * b->cfg.n_heap_slabs + lv[0]) * hw * 8, (size_t)hw * 32, cudaMemcpyDeviceToHost));
  for (uint32_t i = 0; i < n_bytes; i++) {
    uint32_t a = byte_offset + i, w = a / 32, k = a % 32;
    if (w >= hw) break;
    uint32_t limb = slab[(size_t)w * 8 + (31 - k) / 4];
    out[i] = (uint8_t)(limb >> (8 * ((31 - k) % 4)));
  }
  return ZKB_OK;
}

int32_t zkb_hash_bytecodes(int32_t device, const uint8_t* words_be, const uint64_t* offsets_words, uint32_t n, uint8_t marker,
                           uint8_t* hashes_be_out) {
  if (!offsets_words || (n && !hashes_be_out)) return ZKB_ERR_INVALID_ARGUMENT;
  if (n == 0) return ZKB_OK;
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) return set_err(ZKB_ERR_NO_DEVICE, "no CUDA device: this library has no CPU fallback");
  if (device < 0 || device >= n_dev) return ZKB_ERR_INVALID_ARGUMENT;
  const uint64_t total_words = offsets_words[n];
  if (total_words && !words_be) return ZKB_ERR_INVALID_ARGUMENT;
  for (uint32_t i = 0; i < n; i++) {
    if (offsets_words[i + 1] < offsets_words[i]) return ZKB_ERR_INVALID_ARGUMENT;
    if (offsets_words[i + 1] - offsets_words[i] > 0xFFFFu)
      return set_err(ZKB_ERR_INVALID_ARGUMENT, "bytecode longer than 65535 words (MAX_CODE_PAGE_SIZE_IN_WORDS)");
  }
  CUDA_OK(cudaSetDevice(device));
  uint8_t *d_words = nullptr, *d_hashes = nullptr;
  uint64_t* d_off = nullptr;
  cudaError_t e = cudaMalloc(&d_words, std::max<uint64_t>(total_words * 32, 32));
  if (e == cudaSuccess) e = cudaMalloc(&d_off, (size_t)(n + 1) * 8);
  if (e == cudaSuccess) e = cudaMalloc(&d_hashes, (size_t)n * 32);
  if (e == cudaSuccess && total_words) e = cudaMemcpy(d_words, words_be, total_words * 32, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d_off, offsets_words, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    zkb_hash_bytecodes_kernel<<<(n + 127) / 128, 128>>>(d_words, d_off, n, marker, d_hashes);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpy(hashes_be_out, d_hashes, (size_t)n * 32, cudaMemcpyDeviceToHost);
  cudaFree(d_words);
  cudaFree(d_off);
  cudaFree(d_hashes);
  if (e != cudaSuccess) return set_err(ZKB_ERR_CUDA, cudaGetErrorString(e));
  return ZKB_OK;
}

// K14 (alubench.cuh): ops per second of one integer micro-benchmark at full occupancy.  For the pipe peaks an "op" is one
// thread-level instruction (mad.lo.u32 / add.u32 / lop3.b32); for the U256 entries one 256-bit operation of one octet.
int32_t zkb_alu_microbench(int32_t device, uint32_t op, uint32_t iters, double* ops_per_second, float* kernel_ms) {
  if (op >= ALUB_N || iters == 0 || !ops_per_second) return ZKB_ERR_INVALID_ARGUMENT;
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) return set_err(ZKB_ERR_NO_DEVICE, "no CUDA device: this library has no CPU fallback");
  if (device < 0 || device >= n_dev) return ZKB_ERR_INVALID_ARGUMENT;
  CUDA_OK(cudaSetDevice(device));
  int n_sm = 148;
  CUDA_OK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device));
  const int grid = n_sm * 8;   // 8 CTAs x 256 threads = 64 warps per SM
  uint32_t* sink = nullptr;
  const uint32_t seed = 0x5eed0001u;
  CUDA_OK(cudaMalloc(&sink, 1024 * 4));
  cudaEvent_t e0, e1;
  CUDA_OK(cudaEventCreate(&e0));
  CUDA_OK(cudaEventCreate(&e1));
  cudaError_t e = cudaSuccess;
  float best = 0.f;
  for (int rep = 0; rep < 4 && e == cudaSuccess; rep++) {   // first pass = warm-up, then best of three
    cudaEventRecord(e0, 0);
    switch (op) {
      case ALUB_IMAD: zkb_alubench_kernel<ALUB_IMAD><<<grid, 256>>>(iters, seed, sink); break;
      case ALUB_IADD3: zkb_alubench_kernel<ALUB_IADD3><<<grid, 256>>>(iters, seed, sink); break;
      case ALUB_LOP3: zkb_alubench_kernel<ALUB_LOP3><<<grid, 256>>>(iters, seed, sink); break;
      case ALUB_U256_ADD: zkb_alubench_kernel<ALUB_U256_ADD><<<grid, 256>>>(iters, seed, sink); break;
      case ALUB_U256_SUB: zkb_alubench_kernel<ALUB_U256_SUB><<<grid, 256>>>(iters, seed, sink); break;
      case ALUB_U256_MUL: zkb_alubench_kernel<ALUB_U256_MUL><<<grid, 256>>>(iters, seed, sink); break;
      case ALUB_U256_DIV: zkb_alubench_kernel<ALUB_U256_DIV><<<grid, 256>>>(iters, seed, sink); break;
      default: zkb_alubench_kernel<ALUB_U256_SHL><<<grid, 256>>>(iters, seed, sink); break;
    }
    e = cudaGetLastError();
    cudaEventRecord(e1, 0);
    if (e == cudaSuccess) e = cudaEventSynchronize(e1);
    float ms = 0.f;
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && (best == 0.f || ms < best)) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  if (e != cudaSuccess) return set_err(ZKB_ERR_CUDA, cudaGetErrorString(e));
  const double threads = (double)grid * 256.0;
  const double ops = op <= ALUB_LOP3 ? threads * (double)iters * ALUB_CHAINS * ALUB_UNROLL : threads / 8.0 * (double)iters;
  *ops_per_second = ops / ((double)best * 1e-3);
  if (kernel_ms) *kernel_ms = best;
  return ZKB_OK;
}

int32_t zkb_ingest_bytecodes(ZkbBatch* b, const uint8_t* words_be, const uint64_t* offsets_words, uint32_t n, uint8_t* hashes_be_out) {
  if (!b || !offsets_words || (n && !hashes_be_out)) return ZKB_ERR_INVALID_ARGUMENT;
  int32_t rc = zkb_hash_bytecodes(b->cfg.device, words_be, offsets_words, n, (uint8_t)ZK_CODE_AT_REST_MARKER, hashes_be_out);
  if (rc != ZKB_OK) return rc;
  for (uint32_t i = 0; i < n; i++) {
    rc = zkb_load_bytecode(b, hashes_be_out + 32 * (size_t)i, words_be + 32 * offsets_words[i], (uint32_t)(offsets_words[i + 1] - offsets_words[i]));
    if (rc != ZKB_OK) return rc;
  }
  return ZKB_OK;
}

int32_t zkb_peer_push_async(int32_t device, int32_t peer_device, const void* src, void* dst_peer, uint64_t n_bytes, uint32_t n_ctas,
                            void* cuda_stream) {
  if ((!src || !dst_peer) && n_bytes) return ZKB_ERR_INVALID_ARGUMENT;
  if (n_bytes == 0) return ZKB_OK;
  if (((uintptr_t)src | (uintptr_t)dst_peer) & 7u) return set_err(ZKB_ERR_INVALID_ARGUMENT, "zkb_peer_push_async: pointers must be 8-byte aligned");
  const bool wide = ((((uintptr_t)src | (uintptr_t)dst_peer) & 15u) == 0);
  CUDA_OK(cudaSetDevice(device));
  if (peer_device != device) {
    cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e != cudaSuccess) cudaGetLastError();   // already enabled (also by the lazy IPC mapping): fine
  }
  const uint32_t vec = wide ? 16u : 8u;
  const uint64_t n16 = n_bytes / vec;
  const uint32_t n_tail = (uint32_t)(n_bytes % vec);
  const uint8_t* ts = (const uint8_t*)src + n16 * vec;
  uint8_t* td = (uint8_t*)dst_peer + n16 * vec;
  if (wide)
    zkb_peer_push_kernel<uint4><<<std::max(1u, n_ctas), ZKB_PUSH_THREADS, 0, (cudaStream_t)cuda_stream>>>((const uint4*)src, (uint4*)dst_peer, n16, ts, td, n_tail);
  else
    zkb_peer_push_kernel<uint2><<<std::max(1u, n_ctas), ZKB_PUSH_THREADS, 0, (cudaStream_t)cuda_stream>>>((const uint2*)src, (uint2*)dst_peer, n16, ts, td, n_tail);
  CUDA_OK(cudaGetLastError());
  return ZKB_OK;
}

int32_t zkb_peer_sink_create(int32_t device, uint64_t n_bytes, void** dptr_out, uint8_t ipc_handle_out[64]) {
  if (!dptr_out || !ipc_handle_out || n_bytes == 0) return ZKB_ERR_INVALID_ARGUMENT;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
  CUDA_OK(cudaSetDevice(device));
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, n_bytes);
  if (e != cudaSuccess) return set_err(ZKB_ERR_OUT_OF_MEMORY, std::string("peer sink cudaMalloc: ") + cudaGetErrorString(e));
  cudaIpcMemHandle_t h;
  e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    return set_err(ZKB_ERR_CUDA, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e));
  }
  memcpy(ipc_handle_out, &h, 64);
  *dptr_out = p;
  return ZKB_OK;
}

int32_t zkb_peer_sink_open(int32_t device, const uint8_t ipc_handle[64], void** dptr_out) {
  if (!dptr_out || !ipc_handle) return ZKB_ERR_INVALID_ARGUMENT;
  CUDA_OK(cudaSetDevice(device));   // opened with the READER's device current: lazy peer access maps it for this device's kernels
  cudaIpcMemHandle_t h;
  memcpy(&h, ipc_handle, 64);
  void* p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) return set_err(ZKB_ERR_CUDA, std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
  *dptr_out = p;
  return ZKB_OK;
}

int32_t zkb_peer_sink_close(int32_t device, void* dptr, uint32_t owner) {
  if (!dptr) return ZKB_OK;
  CUDA_OK(cudaSetDevice(device));
  CUDA_OK(cudaDeviceSynchronize());
  if (owner) CUDA_OK(cudaFree(dptr));
  else CUDA_OK(cudaIpcCloseMemHandle(dptr));
  return ZKB_OK;
}

int32_t zkb_transfer_stats(ZkbBatch* b, uint64_t* h2d_bytes, uint64_t* d2h_bytes, uint32_t reset) {
  if (!b) return ZKB_ERR_INVALID_ARGUMENT;
  if (h2d_bytes) *h2d_bytes = b->h2d_bytes;
  if (d2h_bytes) *d2h_bytes = b->d2h_bytes;
  if (reset) b->h2d_bytes = b->d2h_bytes = 0;
  return ZKB_OK;
}

int32_t zkb_snapshot(ZkbBatch* b) {
  if (!b) return ZKB_ERR_INVALID_ARGUMENT;
  CUDA_OK(cudaSetDevice(b->cfg.device));
  int32_t rc = upload(b);
  if (rc != ZKB_OK) return rc;
  CUDA_OK(cudaDeviceSynchronize());
  const ZkbConfig& c = b->cfg;
  const DevBatch& d = b->d;
  size_t n = c.n_vms, levels = c.max_far_depth + 1;
  if (b->snap.empty()) {
    auto add = [&](void* live, size_t bytes, bool sparse = false) { b->snap.push_back(ZkbBatch::Region{live, nullptr, bytes, sparse}); };
    add(d.hot, n * sizeof(VmHot));
    add(d.callstack, n * c.max_depth * 128);
    add(d.stack_mem, n * levels * c.stack_words * 32, true);   // regions 2..6: SnapView of zkb_restore (order matters)
    add(d.stack_ptr, n * levels * c.stack_words, true);
    add(d.heap_mem, n * c.n_heap_slabs * (size_t)c.heap_bytes, true);
    add(d.lvl, n * levels * 16);
    add(d.slab_hwm, n * c.n_heap_slabs * 4);
    add(d.pt, n * ZKB_PT_ENTRIES * 8);
    add(d.dec, n * ZKB_DEC_ENTRIES * 8);
    add(d.defer, n * ZKB_DEFER_WORDS * 4);
    add(d.st_tags, n * c.storage_slots * 4);
    add(d.st_keys, n * c.storage_slots * 32);
    add(d.st_addr, n * c.storage_slots * 32);
    add(d.st_vals, n * c.storage_slots * 32);
    for (auto& r : b->snap) {
      cudaError_t e = cudaMalloc(&r.saved, std::max<size_t>(r.bytes, 16));
      if (e != cudaSuccess) {
        for (auto& q : b->snap)
          if (q.saved) cudaFree(q.saved);
        b->snap.clear();
        return set_err(ZKB_ERR_OUT_OF_MEMORY, std::string("snapshot cudaMalloc: ") + cudaGetErrorString(e));
      }
    }
  }
  rc = download_hot(b);
  if (rc != ZKB_OK) return rc;
  for (auto& r : b->snap) CUDA_OK(cudaMemcpy(r.saved, r.live, r.bytes, cudaMemcpyDeviceToDevice));
  b->snap_counts.assign(b->h_counts, b->h_counts + (size_t)b->cfg.n_vms * 8);
  b->has_snapshot = true;
  return ZKB_OK;
}

int32_t zkb_restore(ZkbBatch* b, void* cuda_stream) {
  if (!b || !b->has_snapshot) return set_err(ZKB_ERR_INVALID_ARGUMENT, "zkb_restore without zkb_snapshot");
  CUDA_OK(cudaSetDevice(b->cfg.device));
  cudaStream_t st = (cudaStream_t)cuda_stream;
  if (wait_last_run(b) != ZKB_OK) return ZKB_ERR_CUDA;  // a running launch still owns the state and the host summary
  memcpy(b->h_counts, b->snap_counts.data(), b->snap_counts.size() * sizeof(uint32_t));
  b->offsets_valid = false;
  // stack pages and heap slabs first, by their high-water marks (which the dense copies below then restore as well)
  static const bool dense_only = getenv("ZKB_RESTORE_DENSE") != nullptr;   // experiment / cross-check knob
  if (!dense_only) {
    SnapView sv{(const uint32_t*)b->snap[2].saved, (const uint8_t*)b->snap[3].saved, (const uint32_t*)b->snap[4].saved,
                (const uint32_t*)b->snap[5].saved, (const uint32_t*)b->snap[6].saved};
    int n_sm = 148;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, b->cfg.device);
    const int grid = (int)std::min<uint32_t>((b->cfg.n_vms + 3) / 4, (uint32_t)n_sm * 16);
    zkb_restore_sparse_kernel<<<grid, 128, 0, st>>>(b->d, sv);
    CUDA_OK(cudaGetLastError());
  }
  for (auto& r : b->snap)
    if (dense_only || !r.sparse) CUDA_OK(cudaMemcpyAsync(r.live, r.saved, r.bytes, cudaMemcpyDeviceToDevice, st));
  CUDA_OK(cudaEventRecord(b->ev0, st));
  CUDA_OK(cudaEventRecord(b->ev1, st));
  b->launched = true;
  b->last_stream = st;
  b->hot_dirty = false;
  b->hot_stale = true;
  return ZKB_OK;
}

}  // extern "C"

#include "comm.cuh"   // multi-GPU exchange entry points (NCCL resolved at run time)
