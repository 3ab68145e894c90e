// Warp-cooperative 256-bit integer arithmetic for sm_100a.
//
// Representation: a U256 is ONE 32-bit register per lane; lane l (0..7) holds limb l (little-endian,
// limb 0 = least significant 32 bits), lanes 8..31 hold 0.  All 32 lanes of the warp call every function
// (one VM per warp, warp-uniform control flow).  Carries are resolved with two __ballot_sync votes and one
// integer add (generate/propagate trick) instead of an 8-step ripple, so ADD/SUB cost ~10 warp instructions.
//
// Semantics replaced (reference, ethereum_types::U256 as used in /root/reference/src/opcodes/execution/):
//   u_add  -> overflowing_add  add.rs:35        u_sub -> overflowing_sub  sub.rs:35
//   u_mul  -> full_mul         mul.rs:35-39     u_divmod -> div_mod       div.rs:50
//   u_shl/u_shr (n >= 256 -> 0) shift.rs:49-58, uma.rs:299-361
#pragma once
#include <stdint.h>

#define ZK_FULL 0xffffffffu

namespace zkb {

typedef uint32_t u256l;  // the per-lane limb of a warp-distributed U256

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ uint32_t bcast(uint32_t v, int src) { return __shfl_sync(ZK_FULL, v, src); }

__device__ __forceinline__ bool u_is_zero(u256l v) { return (__ballot_sync(ZK_FULL, v != 0) & 0xFFu) == 0; }
__device__ __forceinline__ bool u_eq(u256l a, u256l b) { return (__ballot_sync(ZK_FULL, a != b) & 0xFFu) == 0; }

// carry-in vector from generate/propagate votes: bit i = carry into limb i, bit n = carry out
__device__ __forceinline__ uint32_t carry_chain(uint32_t G, uint32_t P) {
  uint32_t Gs = G << 1;
  return Gs | ((P + Gs) ^ P ^ Gs);
}

// a + b over `nl` limbs (8 or 16); returns the sum limb, sets carry-out
__device__ __forceinline__ u256l u_add_n(u256l a, u256l b, uint32_t lane, uint32_t nl, bool& carry_out) {
  uint32_t mask = (1u << nl) - 1u;
  uint32_t s = a + b;
  uint32_t G = __ballot_sync(ZK_FULL, s < a) & mask;
  uint32_t P = __ballot_sync(ZK_FULL, s == 0xFFFFFFFFu) & mask;
  uint32_t K = carry_chain(G, P);
  carry_out = (K >> nl) & 1u;
  return lane < nl ? s + ((K >> lane) & 1u) : 0u;
}

__device__ __forceinline__ u256l u_add(u256l a, u256l b, uint32_t lane, bool& of) { return u_add_n(a, b, lane, 8, of); }

__device__ __forceinline__ u256l u_sub(u256l a, u256l b, uint32_t lane, bool& borrow_out) {
  uint32_t d = a - b;
  uint32_t G = __ballot_sync(ZK_FULL, a < b) & 0xFFu;
  uint32_t P = __ballot_sync(ZK_FULL, a == b) & 0xFFu;
  uint32_t K = carry_chain(G, P);
  borrow_out = (K >> 8) & 1u;
  return lane < 8 ? d - ((K >> lane) & 1u) : 0u;
}

// unsigned compare: -1, 0, +1 (the most significant differing limb decides)
__device__ __forceinline__ int u_cmp(u256l a, u256l b) {
  uint32_t gt = __ballot_sync(ZK_FULL, a > b) & 0xFFu;
  uint32_t lt = __ballot_sync(ZK_FULL, a < b) & 0xFFu;
  return gt > lt ? 1 : (gt < lt ? -1 : 0);
}

// 256 x 256 -> 512: lane k (0..15) accumulates column k of the schoolbook product in 96 bits, then two
// 16-limb carry-resolved additions fold the columns.  lo/hi come back in lanes 0..7.
__device__ __forceinline__ void u_mul(u256l a, u256l b, uint32_t lane, u256l& lo, u256l& hi) {
  uint64_t acc = 0;
  uint32_t acc_top = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    uint32_t ai = __shfl_sync(ZK_FULL, a, i);
    int j = (int)lane - i;
    uint32_t bj = __shfl_sync(ZK_FULL, b, j & 7);
    bj = (j >= 0 && j < 8) ? bj : 0u;
    uint64_t prod = (uint64_t)ai * (uint64_t)bj;
    acc += prod;
    acc_top += (acc < prod) ? 1u : 0u;
  }
  uint32_t c0 = (uint32_t)acc, c1 = (uint32_t)(acc >> 32), c2 = acc_top;
  uint32_t y = __shfl_up_sync(ZK_FULL, c1, 1);
  y = lane >= 1 ? y : 0u;
  uint32_t z = __shfl_up_sync(ZK_FULL, c2, 2);
  z = lane >= 2 ? z : 0u;
  bool dummy;
  uint32_t r = u_add_n(lane < 16 ? c0 : 0u, lane < 16 ? y : 0u, lane, 16, dummy);
  r = u_add_n(r, lane < 16 ? z : 0u, lane, 16, dummy);
  uint32_t up = __shfl_down_sync(ZK_FULL, r, 8);
  lo = lane < 8 ? r : 0u;
  hi = lane < 8 ? up : 0u;
}

// logical shifts by n bits; n >= 256 yields 0 (ethereum_types semantics)
__device__ __forceinline__ u256l u_shl(u256l v, uint32_t n, uint32_t lane) {
  uint32_t k = n >> 5, b = n & 31u;
  int s0 = (int)lane - (int)k, s1 = s0 - 1;
  uint32_t x0 = __shfl_sync(ZK_FULL, v, s0 & 31);
  uint32_t x1 = __shfl_sync(ZK_FULL, v, s1 & 31);
  x0 = (s0 >= 0 && s0 < 8) ? x0 : 0u;
  x1 = (s1 >= 0 && s1 < 8) ? x1 : 0u;
  uint32_t r = __funnelshift_l(x1, x0, b);
  return (lane < 8 && n < 256) ? r : 0u;
}

__device__ __forceinline__ u256l u_shr(u256l v, uint32_t n, uint32_t lane) {
  uint32_t k = n >> 5, b = n & 31u;
  uint32_t s0 = lane + k, s1 = s0 + 1;
  uint32_t x0 = __shfl_sync(ZK_FULL, v, s0 & 31);
  uint32_t x1 = __shfl_sync(ZK_FULL, v, s1 & 31);
  x0 = s0 < 8 ? x0 : 0u;
  x1 = s1 < 8 ? x1 : 0u;
  uint32_t r = __funnelshift_r(x0, x1, b);
  return (lane < 8 && n < 256) ? r : 0u;
}

// number of significant bits (0 for zero)
__device__ __forceinline__ uint32_t u_bits(u256l v) {
  uint32_t nz = __ballot_sync(ZK_FULL, v != 0) & 0xFFu;
  if (nz == 0) return 0;
  int top = 31 - __clz(nz);
  uint32_t tv = __shfl_sync(ZK_FULL, v, top);
  return (uint32_t)top * 32u + (32u - (uint32_t)__clz(tv));
}

// div_mod for b != 0: restoring shift-subtract over the significant bit range only (bits(a) - bits(b) + 1 steps).
__device__ __forceinline__ void u_divmod(u256l a, u256l b, uint32_t lane, u256l& q, u256l& r) {
  uint32_t na = u_bits(a), nb = u_bits(b);
  q = 0;
  if (na < nb) {
    r = a;
    return;
  }
  uint32_t sh = na - nb;
  u256l d = u_shl(b, sh, lane);  // aligned divisor (fits: bits(d) == na <= 256)
  u256l rem = a;
  for (int i = (int)sh; i >= 0; i--) {
    uint32_t gt = __ballot_sync(ZK_FULL, rem > d) & 0xFFu;
    uint32_t lt = __ballot_sync(ZK_FULL, rem < d) & 0xFFu;
    if (gt >= lt) {  // rem >= d (warp-uniform branch)
      bool bo;
      rem = u_sub(rem, d, lane, bo);
      if (lane == (uint32_t)(i >> 5)) q |= 1u << (i & 31);
    }
    // d >>= 1
    uint32_t up = __shfl_down_sync(ZK_FULL, d, 1);
    up = lane < 7 ? up : 0u;
    d = lane < 8 ? __funnelshift_r(d, up, 1) : 0u;
  }
  r = rem;
}

__device__ __forceinline__ uint32_t bswap32(uint32_t x) { return __byte_perm(x, 0, 0x0123); }

}  // namespace zkb
