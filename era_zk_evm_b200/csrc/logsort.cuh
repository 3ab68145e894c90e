// K11: grouping of log queries by storage slot -- the second half of SURVEY §8 row f-2.
//
// Reference: InMemoryStorage::flatten_and_net_history().1 (/root/reference/src/testing/storage.rs:50-73) files every
// query of the forward history under its slot (shard_id, address, key), keeping history order inside a slot:
// `HashMap<(u8, Address, U256), Vec<LogQuery>>`.  On the device that is ONE stable sort of the records by a 64-bit key
//     key = group << 44 | hash44(shard, address, key)           (group: caller-chosen, e.g. the VM index, or 0)
// followed by a gather of the 128-byte records and a boundary flag per record (1 = first query of its slot; equal
// hashes of different slots are told apart by comparing the full identity, so a hash collision costs a spurious
// neighbourhood, never a wrong group).  Stable LSD radix sort, 8 bits per pass, over (key, index) pairs: each pass is
// histogram -> exclusive scan -> stable scatter; passes whose digit is constant over the input are skipped.
// HBM-bound: per pass 12 B read + 12 B written per record, plus one 128 B read + 128 B write for the gather.
#pragma once
#include <stdint.h>

#include "../../include/zkb_records.h"

namespace zkb {

#define ZKB_SORT_THREADS 256
#define ZKB_SORT_CHUNKS 16
#define ZKB_SORT_TILE (ZKB_SORT_THREADS * ZKB_SORT_CHUNKS)

// 64-bit identity hash of a slot (shard, address, key); the host restatement lives in tests/ and oracle/
__host__ __device__ __forceinline__ uint64_t slot_hash64(uint32_t shard, const uint32_t* addr_words, const uint32_t* key_limbs) {
  uint64_t h = 0x9E3779B97F4A7C15ull ^ shard;
#pragma unroll
  for (int i = 0; i < 5; i++) {
    h = (h ^ addr_words[i]) * 0xFF51AFD7ED558CCDull;
    h ^= h >> 32;
  }
#pragma unroll
  for (int i = 0; i < 8; i++) {
    h = (h ^ key_limbs[i]) * 0xFF51AFD7ED558CCDull;
    h ^= h >> 32;
  }
  return h;
}

// sort keys of n LogQueryRec records; group_of (optional): per-record group id (20 bits used), else 0
__global__ void __launch_bounds__(256) zkb_logsort_keys_kernel(const uint32_t* __restrict__ recs, uint64_t n, const uint32_t* __restrict__ group_of,
                                                               uint64_t* __restrict__ keys, uint32_t* __restrict__ idx) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint4* r = reinterpret_cast<const uint4*>(recs + i * 32);
  const uint4 a = __ldg(r), b = __ldg(r + 1), k0 = __ldg(r + 2), k1 = __ldg(r + 3);
  const uint32_t addr[5] = {a.z, a.w, b.x, b.y, b.z};
  const uint32_t key[8] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
  const uint64_t h = slot_hash64(a.y >> 24, addr, key);
  const uint64_t g = group_of ? (uint64_t)(group_of[i] & 0xFFFFFu) : 0ull;
  keys[i] = g << 44 | (h >> 20);
  idx[i] = (uint32_t)i;
}

// pass 1 of a radix pass: digit histogram per tile -> hist[digit][tile]
__global__ void __launch_bounds__(ZKB_SORT_THREADS) zkb_logsort_hist_kernel(const uint64_t* __restrict__ keys, uint64_t n, uint32_t shift,
                                                                            uint32_t* __restrict__ hist, uint32_t n_tiles) {
  __shared__ uint32_t s_cnt[256];
  s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const uint64_t base = (uint64_t)blockIdx.x * ZKB_SORT_TILE;
#pragma unroll 4
  for (int c = 0; c < ZKB_SORT_CHUNKS; c++) {
    const uint64_t i = base + (uint64_t)c * ZKB_SORT_THREADS + threadIdx.x;
    if (i < n) atomicAdd(&s_cnt[(uint32_t)(keys[i] >> shift) & 255u], 1u);
  }
  __syncthreads();
  hist[(size_t)threadIdx.x * n_tiles + blockIdx.x] = s_cnt[threadIdx.x];
}

// exclusive scan of hist[256 * n_tiles] (digit-major) in place, one block; also reports whether one digit holds everything
__global__ void __launch_bounds__(1024) zkb_logsort_scan_kernel(uint32_t* __restrict__ hist, uint32_t n_items, uint64_t n, uint32_t* __restrict__ trivial_out) {
  __shared__ uint32_t s_sum[1024];
  const uint32_t t = threadIdx.x, per = (n_items + 1023) / 1024;
  const uint32_t lo = min(n_items, t * per), hi = min(n_items, lo + per);
  uint32_t local = 0;
  for (uint32_t i = lo; i < hi; i++) local += hist[i];
  s_sum[t] = local;
  __syncthreads();
  for (uint32_t o = 1; o < 1024; o <<= 1) {
    uint32_t v = t >= o ? s_sum[t - o] : 0;
    __syncthreads();
    s_sum[t] += v;
    __syncthreads();
  }
  uint32_t run = s_sum[t] - local;
  for (uint32_t i = lo; i < hi; i++) {
    const uint32_t c = hist[i];
    hist[i] = run;
    run += c;
  }
  (void)n;
  (void)trivial_out;
}

// pass 3: stable scatter.  A tile is processed chunk by chunk (256 consecutive items, one per thread); inside a chunk
// an item's rank among the items of the same digit is (same digit in earlier warps) + (same digit in lower lanes).
__global__ void __launch_bounds__(ZKB_SORT_THREADS) zkb_logsort_scatter_kernel(const uint64_t* __restrict__ keys_in, const uint32_t* __restrict__ idx_in, uint64_t n,
                                                                               uint32_t shift, const uint32_t* __restrict__ hist, uint32_t n_tiles,
                                                                               uint64_t* __restrict__ keys_out, uint32_t* __restrict__ idx_out) {
  __shared__ uint32_t s_base[256];          // next output slot of each digit for this tile
  __shared__ uint32_t s_warp[8][256];       // per-warp digit counts of the current chunk
  const uint32_t t = threadIdx.x, lane = t & 31u, warp = t >> 5;
  s_base[t] = hist[(size_t)t * n_tiles + blockIdx.x];
  const uint64_t base = (uint64_t)blockIdx.x * ZKB_SORT_TILE;
  for (int c = 0; c < ZKB_SORT_CHUNKS; c++) {
    for (int w = 0; w < 8; w++) s_warp[w][t] = 0;
    __syncthreads();
    const uint64_t i = base + (uint64_t)c * ZKB_SORT_THREADS + t;
    const bool valid = i < n;
    uint64_t k = 0;
    uint32_t v = 0, d = 0xFFFFFFFFu;
    if (valid) {
      k = keys_in[i];
      v = idx_in[i];
      d = (uint32_t)(k >> shift) & 255u;
    }
    const uint32_t peers = __match_any_sync(0xffffffffu, d);
    const uint32_t rank_in_warp = __popc(peers & ((1u << lane) - 1u));
    if (valid && rank_in_warp == 0) s_warp[warp][d] = __popc(peers);
    __syncthreads();
    if (valid) {
      uint32_t before = 0;
      for (uint32_t w = 0; w < warp; w++) before += s_warp[w][d];
      const uint32_t pos = s_base[d] + before + rank_in_warp;
      keys_out[pos] = k;
      idx_out[pos] = v;
    }
    __syncthreads();
    uint32_t tot = 0;
    for (int w = 0; w < 8; w++) tot += s_warp[w][t];
    s_base[t] += tot;
    __syncthreads();
  }
}

// gather: out[i] = recs[idx[i]], boundary[i] = 1 when record i opens a new (group, slot)
__global__ void __launch_bounds__(256) zkb_logsort_gather_kernel(const uint32_t* __restrict__ recs, const uint64_t* __restrict__ keys, const uint32_t* __restrict__ idx,
                                                                 uint64_t n, uint32_t* __restrict__ out, uint8_t* __restrict__ boundary,
                                                                 unsigned long long* __restrict__ n_groups) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= n) return;
  const uint32_t w = __ldg(recs + (size_t)idx[i] * 32 + lane);
  out[i * 32 + lane] = w;
  // identity words of a slot: shard (byte 3 of word 1), address words 2..6, key words 8..15
  bool differs = true;
  if (i > 0 && keys[i] == keys[i - 1]) {
    const uint32_t p = __ldg(recs + (size_t)idx[i - 1] * 32 + lane);
    const bool ident = (lane == 1) || (lane >= 2 && lane <= 6) || (lane >= 8 && lane <= 15);
    const uint32_t mask = lane == 1 ? 0xFF000000u : 0xFFFFFFFFu;
    differs = __any_sync(0xffffffffu, ident && ((w ^ p) & mask) != 0);
  }
  if (lane == 0) {
    boundary[i] = differs ? 1 : 0;
    if (differs) atomicAdd(n_groups, 1ull);
  }
}

}  // namespace zkb
