// Octet-cooperative keccak-f[1600]: the 5 x 5 x 64 state of one VM lives in 5 of the 8 lanes of its octet -- octet lane
// x (0..4) holds column x, a[y] = A[x][y] (five u64 = ten registers); lanes 5..7 carry zeros and only take part in the
// octet's shuffles.  Per round: theta's column parity is lane-local (the column IS the lane), D needs two 64-bit
// shuffles; rho + pi scatter the rotated lanes through a 25 x u64 shared-memory scratch of the VM, chi reads its three
// inputs per output back from the same scratch (conflict-free), iota on lane 0.  ~90 warp instructions per round for
// the four VMs of a warp.
// Replaces the keccak256 round function of the external `DefaultPrecompilesProcessor`
// (zk_evm_abstractions@v1.4.1, called from /root/reference/src/vm_state/helpers.rs:211-213; pinned by the live
// tests at src/testing/tests/precompiles/keccak256.rs:144-196).
#pragma once
#include <stdint.h>
#include "u256.cuh"

namespace zkb {

__constant__ uint64_t c_keccak_rc[24] = {
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL, 0x000000000000808bULL,
    0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008aULL, 0x0000000000000088ULL,
    0x0000000080008009ULL, 0x000000008000000aULL, 0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL,
    0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
    0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};

// rho offsets r[x + 5y]
__constant__ uint8_t c_keccak_rot[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};

__device__ __forceinline__ uint64_t oshfl64(uint64_t v, int src) {
  uint32_t lo = oshfl((uint32_t)v, src);
  uint32_t hi = oshfl((uint32_t)(v >> 32), src);
  return ((uint64_t)hi << 32) | lo;
}

// variable rotate: swap the halves for n >= 32, then two funnel shifts by n mod 32 (a shift of 0 returns the high operand)
__device__ __forceinline__ uint64_t rotl64(uint64_t x, uint32_t n) {
#ifdef ZKB_OLD_ROTL
  n &= 63u;
  return n ? (x << n) | (x >> (64u - n)) : x;
#endif
  const uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
  const bool sw = (n & 32u) != 0;
  const uint32_t a = sw ? hi : lo, b = sw ? lo : hi;   // (b : a) = x rotated by 0 or 32
  const uint32_t m = n & 31u;
  return ((uint64_t)__funnelshift_l(a, b, m) << 32) | __funnelshift_l(b, a, m);
}

struct KeccakState {
  uint64_t a[5];  // column `octet lane` of the state (zeros in lanes 5..7)
};

struct KeccakLanes {
  int xm1, xp1, xp2;  // octet lanes holding columns x-1 / x+1 / x+2 (own lane for lanes 5..7)
  uint32_t rot;       // rho offsets of this column, 6 bits per y
  uint32_t dst;       // scratch slot X + 5Y of pi's destination B[X = y][Y = 2x + 3y], 5 bits per y (31 = idle lane: slot 25+)
};

__device__ __forceinline__ KeccakLanes keccak_lanes(uint32_t lane) {
  KeccakLanes k;
  if (lane < 5) {
    int x = (int)lane;
    k.xm1 = (x + 4) % 5;
    k.xp1 = (x + 1) % 5;
    k.xp2 = (x + 2) % 5;
    k.rot = 0;
    k.dst = 0;
#pragma unroll
    for (int y = 0; y < 5; y++) {
      k.rot |= (uint32_t)c_keccak_rot[x + 5 * y] << (6 * y);
      k.dst |= (uint32_t)(y + 5 * ((2 * x + 3 * y) % 5)) << (5 * y);
    }
  } else {
    k.xm1 = k.xp1 = k.xp2 = (int)lane;
    k.rot = 0;
    k.dst = 0;
  }
  return k;
}

// ks: the VM's 25 x u64 shared-memory scratch.  All 8 lanes of the octet must call.
__device__ __forceinline__ void keccak_f1600(KeccakState& s, const KeccakLanes& k, uint64_t* ks, uint32_t lane) {
  const bool col = lane < 5;
#pragma unroll 1
  for (int round = 0; round < 24; round++) {
    // theta
    uint64_t* kb = ks;
    uint64_t c = s.a[0] ^ s.a[1] ^ s.a[2] ^ s.a[3] ^ s.a[4];
    uint64_t d = oshfl64(c, k.xm1) ^ rotl64(oshfl64(c, k.xp1), 1);
    // rho + pi: B[y][2x + 3y] = rotl(A[x][y] ^ D[x], r[x][y])
    if (col) {
#pragma unroll
      for (int y = 0; y < 5; y++) kb[(k.dst >> (5 * y)) & 31u] = rotl64(s.a[y] ^ d, (k.rot >> (6 * y)) & 63u);
    }
    osync();
    // chi: A'[X][Y] = B[X][Y] ^ (~B[X+1][Y] & B[X+2][Y])
    if (col) {
#pragma unroll
      for (int y = 0; y < 5; y++) {
        uint64_t b0 = kb[lane + 5 * y], b1 = kb[k.xp1 + 5 * y], b2 = kb[k.xp2 + 5 * y];
        s.a[y] = b0 ^ (~b1 & b2);
      }
      if (lane == 0) s.a[0] ^= c_keccak_rc[round];  // iota
    }
    osync();
  }
}

}  // namespace zkb

// ---------------------------------------------------------------------------------------------------------------------
// Thread-per-state keccak256 sponge over a VM heap slab: the DEFERRED half of the keccak256 precompile.
//
// The interpreter cycle that executes the precompile only moves data (it emits the memory-read witness, reserves the
// output word and stashes a descriptor); the permutations run after the cycle, batched over the VMs of the CTA with
// ONE THREAD PER STATE: 25 x u64 in registers, every rotation amount a compile-time constant (2 funnel shifts), chi as
// one LOP3 per 32-bit half -- ~190 instructions per round and state against ~720 lane-instructions (90 warp
// instructions for four VMs on 32 lanes) of the octet-cooperative layout above, which stays in use only for the
// ecrecover address hash.
namespace zkb {

// 64-bit rotate by a compile-time amount as TWO funnel shifts (SHF.L.W); written as (x << n) | (x >> (64 - n)) nvcc emits
// plain shifts + merges, ~4 instructions per rotation (29 rotations per round)
__device__ __forceinline__ uint64_t kc_rotl(uint64_t x, const int n) {
  const uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
  if (n == 0) return x;
  if (n == 32) return ((uint64_t)lo << 32) | hi;
  if (n < 32) return ((uint64_t)__funnelshift_l(lo, hi, n) << 32) | __funnelshift_l(hi, lo, n);
  return ((uint64_t)__funnelshift_l(hi, lo, n - 32) << 32) | __funnelshift_l(lo, hi, n - 32);
}

__device__ __forceinline__ void keccak_f1600_regs(uint64_t (&st)[25]) {
#pragma unroll 1
  for (int round = 0; round < 24; round++) {
    uint64_t bc[5];
#pragma unroll
    for (int x = 0; x < 5; x++) bc[x] = st[x] ^ st[x + 5] ^ st[x + 10] ^ st[x + 15] ^ st[x + 20];
#pragma unroll
    for (int x = 0; x < 5; x++) {
      const uint64_t d = bc[(x + 4) % 5] ^ kc_rotl(bc[(x + 1) % 5], 1);
#pragma unroll
      for (int y = 0; y < 25; y += 5) st[x + y] ^= d;
    }
    // rho + pi along the 24-cycle of the lane permutation
    constexpr int piln[24] = {10, 7, 11, 17, 18, 3, 5, 16, 8, 21, 24, 4, 15, 23, 19, 13, 12, 2, 20, 14, 22, 9, 6, 1};
    constexpr int rotc[24] = {1, 3, 6, 10, 15, 21, 28, 36, 45, 55, 2, 14, 27, 41, 56, 8, 25, 43, 62, 18, 39, 61, 20, 44};
    uint64_t t = st[1];
#pragma unroll
    for (int i = 0; i < 24; i++) {
      const uint64_t keep = st[piln[i]];
      st[piln[i]] = kc_rotl(t, rotc[i]);
      t = keep;
    }
    // chi
#pragma unroll
    for (int y = 0; y < 25; y += 5) {
#pragma unroll
      for (int x = 0; x < 5; x++) bc[x] = st[y + x];
#pragma unroll
      for (int x = 0; x < 5; x++) st[y + x] = bc[x] ^ (~bc[(x + 1) % 5] & bc[(x + 2) % 5]);
    }
    st[0] ^= c_keccak_rc[round];
  }
}

__device__ __forceinline__ uint64_t kc_bswap64(uint64_t v) {
  return ((uint64_t)__byte_perm((uint32_t)v, 0, 0x0123) << 32) | __byte_perm((uint32_t)(v >> 32), 0, 0x0123);
}

// 8 bytes of the big-endian byte stream of a heap slab, starting at stream unit q (8-byte aligned position 8 q), as a
// little-endian u64 (stream byte 8 q is the least significant).  A U256 word is stored as 8 little-endian u32 limbs =
// the 32 stream bytes reversed, so the unit is one aligned 8-byte load, byte-swapped.  Beyond the slab: zero.
__device__ __forceinline__ uint64_t kc_stream_unit(const uint8_t* slab, uint64_t q, uint32_t heap_words) {
  const uint64_t w = q >> 2;
  if (slab == nullptr || w >= heap_words) return 0ull;
  return kc_bswap64(*reinterpret_cast<const uint64_t*>(slab + w * 32 + 24 - (q & 3u) * 8));
}

// keccak256 of stream bytes [in_off, in_off + in_len) of `slab`; digest as four little-endian u64 (the first 32 bytes)
__device__ __forceinline__ void keccak256_slab(const uint8_t* slab, uint32_t heap_words, uint32_t in_off, uint32_t in_len, uint64_t digest[4]) {
  uint64_t st[25];
#pragma unroll
  for (int i = 0; i < 25; i++) st[i] = 0ull;
  const uint64_t end = (uint64_t)in_off + in_len;
  const uint32_t n_blocks = in_len / 136 + 1;
  const uint32_t k8 = (in_off & 7u) * 8;
#pragma unroll 1
  for (uint32_t blk = 0; blk < n_blocks; blk++) {
    const uint64_t a0 = (uint64_t)in_off + (uint64_t)blk * 136;
    const uint32_t nb = (uint32_t)min((uint64_t)136, end - a0);  // valid bytes in this block
    uint64_t q = a0 >> 3;
    uint64_t cur = nb ? kc_stream_unit(slab, q, heap_words) : 0ull;
#pragma unroll
    for (int i = 0; i < 17; i++) {
      const uint32_t b0 = 8u * (uint32_t)i;
      uint64_t v = 0ull;
      if (b0 < nb) {
        v = cur;
        if (k8) {  // the rate word straddles two aligned units
          const uint64_t nxt = kc_stream_unit(slab, q + 1, heap_words);
          v = (cur >> k8) | (nxt << (64 - k8));
          cur = nxt;
        } else {
          cur = kc_stream_unit(slab, q + 1, heap_words);
        }
        q++;
        if (b0 + 8 > nb) v &= (1ull << (8 * (nb - b0))) - 1ull;
      }
      if (nb < 136) {  // final block: pad10*1 with the keccak domain byte 0x01
        if ((uint32_t)i == nb / 8) v ^= 1ull << (8 * (nb % 8));
        if (i == 16) v ^= 0x80ull << 56;
      }
      st[i] ^= v;
    }
    keccak_f1600_regs(st);
  }
#pragma unroll
  for (int i = 0; i < 4; i++) digest[i] = st[i];
}

}  // namespace zkb
