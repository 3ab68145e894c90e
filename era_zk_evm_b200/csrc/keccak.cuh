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

__device__ __forceinline__ uint64_t rotl64(uint64_t x, uint32_t n) {
  n &= 63u;
  return n ? (x << n) | (x >> (64u - n)) : x;
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
