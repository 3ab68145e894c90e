// Warp-cooperative keccak-f[1600]: lane i (0..24) holds state lane A[x + 5y] with i = x + 5y as one u64.
// One round = 9 64-bit shuffles (theta 4+2, rho/pi 1, chi 2) + ~20 ALU ops; 24 rounds per permutation.
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

__device__ __forceinline__ uint64_t shfl64(uint64_t v, int src) {
  uint32_t lo = __shfl_sync(ZK_FULL, (uint32_t)v, src);
  uint32_t hi = __shfl_sync(ZK_FULL, (uint32_t)(v >> 32), src);
  return ((uint64_t)hi << 32) | lo;
}

__device__ __forceinline__ uint64_t rotl64(uint64_t x, uint32_t n) {
  n &= 63u;
  return n ? (x << n) | (x >> (64u - n)) : x;
}

struct KeccakLanes {
  int c5, c10, c15, c20;  // same column, other rows
  int xm1, xp1, xp2;      // same row, x-1 / x+1 / x+2
  int pi_src;             // lane whose value lands here after rho+pi
  uint32_t pi_rot;        // rotation applied to that value
};

__device__ __forceinline__ KeccakLanes keccak_lanes(uint32_t lane) {
  KeccakLanes k;
  if (lane < 25) {
    int x = lane % 5, y = lane / 5;
    k.c5 = (lane + 5) % 25;
    k.c10 = (lane + 10) % 25;
    k.c15 = (lane + 15) % 25;
    k.c20 = (lane + 20) % 25;
    k.xm1 = y * 5 + (x + 4) % 5;
    k.xp1 = y * 5 + (x + 1) % 5;
    k.xp2 = y * 5 + (x + 2) % 5;
    // B[y'][2x'+3y'] = rot(A[x'][y']): this lane is (X, Y) with X = y', Y = (2x' + 3y') % 5  =>  y' = X, x' = (Y - 3X) / 2 mod 5
    int yp = x;
    int xp = ((y - 3 * x) % 5 + 5) % 5;
    xp = (xp * 3) % 5;  // multiply by 2^-1 = 3 (mod 5)
    k.pi_src = xp + 5 * yp;
    k.pi_rot = c_keccak_rot[k.pi_src];
  } else {
    k.c5 = k.c10 = k.c15 = k.c20 = k.xm1 = k.xp1 = k.xp2 = k.pi_src = (int)lane;
    k.pi_rot = 0;
  }
  return k;
}

__device__ __forceinline__ uint64_t keccak_f1600(uint64_t a, const KeccakLanes& k, uint32_t lane) {
#pragma unroll 1
  for (int round = 0; round < 24; round++) {
    // column parity in three exchanges instead of four: rows {y, y+1}, then {y .. y+3}, then row y+4
    uint64_t t2 = a ^ shfl64(a, k.c5);
    t2 ^= shfl64(t2, k.c10);
    uint64_t c = t2 ^ shfl64(a, k.c20);
    uint64_t d = shfl64(c, k.xm1) ^ rotl64(shfl64(c, k.xp1), 1);
    a ^= d;
    uint64_t b = rotl64(shfl64(a, k.pi_src), k.pi_rot);
    a = b ^ (~shfl64(b, k.xp1) & shfl64(b, k.xp2));
    if (lane == 0) a ^= c_keccak_rc[round];
  }
  return a;
}

}  // namespace zkb
