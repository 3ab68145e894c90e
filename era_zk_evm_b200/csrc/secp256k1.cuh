// Octet-cooperative secp256k1 public-key recovery for the ecrecover precompile (external `DefaultPrecompilesProcessor`,
// zk_evm_abstractions@v1.4.1, selected by address 0x01 from /root/reference/src/vm_state/helpers.rs:211-213;
// known-answer vectors: src/testing/tests/precompiles/ecrecover.rs:127-143).
//
// Every field / scalar element is an octet-distributed U256 (u256.cuh: limb l in octet lane l), so a point is three
// registers per lane and the whole recovery keeps its state in registers; the four VMs of a warp recover their keys
// side by side.  A modular multiplication is one 256x256->512 octet multiply (u_mul) followed by pseudo-Mersenne folding: 2^256 = C (mod M) with C = 2^256 - M, i.e.
// value = lo + hi * C repeated until the high half is zero (C is 33 bits for the field prime, 129 bits for the group
// order: at most 2 resp. 3 folds).  Q = u1 G + u2 R runs as one interleaved double-and-add over both scalars.
// The helpers are __noinline__ on purpose: the recovery is ~6 000 modular multiplications, and inlining them would
// add tens of KB of SASS to an interpreter whose instruction-cache footprint is already its first-order cost.
#pragma once
#include <stdint.h>
#include "keccak.cuh"
#include "u256.cuh"

namespace zkb {
namespace secp {

__constant__ uint32_t c_p[8] = {0xFFFFFC2Fu, 0xFFFFFFFEu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
__constant__ uint32_t c_n[8] = {0xD0364141u, 0xBFD25E8Cu, 0xAF48A03Bu, 0xBAAEDCE6u, 0xFFFFFFFEu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
__constant__ uint32_t c_cp[8] = {0x000003D1u, 0x00000001u, 0, 0, 0, 0, 0, 0};                                  // 2^256 - p
__constant__ uint32_t c_cn[8] = {0x2FC9BEBFu, 0x402DA173u, 0x50B75FC4u, 0x45512319u, 0x00000001u, 0, 0, 0};    // 2^256 - n
__constant__ uint32_t c_gx[8] = {0x16F81798u, 0x59F2815Bu, 0x2DCE28D9u, 0x029BFCDBu, 0xCE870B07u, 0x55A06295u, 0xF9DCBBACu, 0x79BE667Eu};
__constant__ uint32_t c_gy[8] = {0xFB10D4B8u, 0x9C47D08Fu, 0xA6855419u, 0xFD17B448u, 0x0E1108A8u, 0x5DA4FBFCu, 0x26A3C465u, 0x483ADA77u};

struct Mod {
  u256l m, c;  // modulus and 2^256 - modulus
};
__device__ __forceinline__ u256l ld8(const uint32_t* t, uint32_t lane) { return t[lane]; }
__device__ __forceinline__ Mod mod_p(uint32_t lane) { return Mod{ld8(c_p, lane), ld8(c_cp, lane)}; }
__device__ __forceinline__ Mod mod_n(uint32_t lane) { return Mod{ld8(c_n, lane), ld8(c_cn, lane)}; }
__device__ __forceinline__ u256l small(uint32_t v, uint32_t lane) { return lane == 0 ? v : 0u; }

__device__ __forceinline__ u256l mod_add(u256l a, u256l b, Mod M, uint32_t lane) {
  bool of, bo;
  u256l s = u_add(a, b, lane, of);
  if (of || u_cmp(s, M.m) >= 0) s = u_sub(s, M.m, lane, bo);
  return s;
}
__device__ __forceinline__ u256l mod_sub(u256l a, u256l b, Mod M, uint32_t lane) {
  bool bo, of;
  u256l d = u_sub(a, b, lane, bo);
  if (bo) d = u_add(d, M.m, lane, of);
  return d;
}
// (a * b) mod M.m
__device__ __noinline__ u256l mod_mul(u256l a, u256l b, u256l m, u256l c, uint32_t lane) {
  u256l lo, hi;
  u_mul(a, b, lane, lo, hi);
  while (!u_is_zero(hi)) {  // value = lo + hi * 2^256 = lo + hi * c (mod m); octet-uniform loop, <= 3 rounds
    u256l plo, phi;
    u_mul(hi, c, lane, plo, phi);
    bool of, of2;
    lo = u_add(lo, plo, lane, of);
    hi = u_add(phi, small(of ? 1u : 0u, lane), lane, of2);
  }
  bool bo;
  while (u_cmp(lo, m) >= 0) lo = u_sub(lo, m, lane, bo);
  return lo;
}
__device__ __forceinline__ u256l mmul(u256l a, u256l b, Mod M, uint32_t lane) { return mod_mul(a, b, M.m, M.c, lane); }

// base^e mod M, left-to-right square-and-multiply over the 256 exponent bits (exponents here are public constants)
__device__ __noinline__ u256l mod_pow(u256l base, u256l e, u256l m, u256l c, uint32_t lane) {
  u256l r = small(1u, lane);
#pragma unroll 1
  for (int bit = 255; bit >= 0; bit--) {
    r = mod_mul(r, r, m, c, lane);
    uint32_t limb = oshfl(e, bit >> 5);
    if ((limb >> (bit & 31)) & 1u) r = mod_mul(r, base, m, c, lane);
  }
  return r;
}

struct JPoint {
  u256l x, y, z;  // z == 0 <=> point at infinity
};

// dbl-2009-l (a = 0)
__device__ __noinline__ JPoint jdouble(JPoint p, uint32_t lane) {
  const Mod P = mod_p(lane);
  if (u_is_zero(p.z) || u_is_zero(p.y)) return JPoint{0u, small(1u, lane), 0u};
  u256l a = mmul(p.x, p.x, P, lane), b = mmul(p.y, p.y, P, lane), c = mmul(b, b, P, lane);
  u256l xb = mod_add(p.x, b, P, lane);
  u256l d = mod_sub(mod_sub(mmul(xb, xb, P, lane), a, P, lane), c, P, lane);
  d = mod_add(d, d, P, lane);
  u256l e = mod_add(mod_add(a, a, P, lane), a, P, lane), f = mmul(e, e, P, lane);
  u256l x3 = mod_sub(f, mod_add(d, d, P, lane), P, lane);
  u256l c8 = mod_add(c, c, P, lane);
  c8 = mod_add(c8, c8, P, lane);
  c8 = mod_add(c8, c8, P, lane);
  u256l y3 = mod_sub(mmul(e, mod_sub(d, x3, P, lane), P, lane), c8, P, lane);
  u256l yz = mmul(p.y, p.z, P, lane);
  return JPoint{x3, y3, mod_add(yz, yz, P, lane)};
}

// mixed addition, q = (qx, qy) affine and not infinity
__device__ __noinline__ JPoint jadd_affine(JPoint p, u256l qx, u256l qy, uint32_t lane) {
  const Mod P = mod_p(lane);
  if (u_is_zero(p.z)) return JPoint{qx, qy, small(1u, lane)};
  u256l zz = mmul(p.z, p.z, P, lane);
  u256l u2 = mmul(qx, zz, P, lane), s2 = mmul(qy, mmul(p.z, zz, P, lane), P, lane);
  u256l h = mod_sub(u2, p.x, P, lane), r = mod_sub(s2, p.y, P, lane);
  if (u_is_zero(h)) {
    if (u_is_zero(r)) return jdouble(p, lane);
    return JPoint{0u, small(1u, lane), 0u};
  }
  u256l hh = mmul(h, h, P, lane), hhh = mmul(h, hh, P, lane), v = mmul(p.x, hh, P, lane);
  u256l x3 = mod_sub(mod_sub(mmul(r, r, P, lane), hhh, P, lane), mod_add(v, v, P, lane), P, lane);
  u256l y3 = mod_sub(mmul(r, mod_sub(v, x3, P, lane), P, lane), mmul(p.y, hhh, P, lane), P, lane);
  return JPoint{x3, y3, mmul(p.z, h, P, lane)};
}

// SEC 1 v2 §4.1.6 with x = r (no r + n candidate: the precompile's recovery id is one bit).  All 8 lanes of the octet must call; ks = the VM's keccak scratch.
// Returns false when nothing can be recovered; otherwise `address` = keccak256(Qx || Qy)[12..] right-aligned in a word.
__device__ __noinline__ bool ecrecover_octet(u256l hash, u256l r, u256l s, uint32_t v_odd, uint32_t lane, uint64_t* ks, u256l& address) {
  address = 0u;
  const Mod P = mod_p(lane), N = mod_n(lane);
  if (u_is_zero(r) || u_is_zero(s) || u_cmp(r, N.m) >= 0 || u_cmp(s, N.m) >= 0) return false;
  // y^2 = x^3 + 7; p = 3 (mod 4) => y = rhs^((p + 1) / 4)
  u256l x = r;
  u256l rhs = mod_add(mmul(mmul(x, x, P, lane), x, P, lane), small(7u, lane), P, lane);
  // (p + 1) / 4 = 0x3FFFFFFF FFFFFFFF ... FFFFFFFF BFFFFF0C
  u256l e_sqrt = lane == 0 ? 0xBFFFFF0Cu : lane == 7 ? 0x3FFFFFFFu : 0xFFFFFFFFu;
  u256l y = mod_pow(rhs, e_sqrt, P.m, P.c, lane);
  if (!u_eq(mmul(y, y, P, lane), rhs)) return false;
  bool bo;
  if ((oshfl(y, 0) & 1u) != (v_odd & 1u)) y = u_sub(P.m, y, lane, bo);
  u256l z = hash;
  while (u_cmp(z, N.m) >= 0) z = u_sub(z, N.m, lane, bo);
  // r^-1 = r^(n - 2) (mod n)
  u256l e_inv = u_sub(N.m, small(2u, lane), lane, bo);
  u256l rinv = mod_pow(r, e_inv, N.m, N.c, lane);
  u256l u1 = mmul(mod_sub(0u, z, N, lane), rinv, N, lane), u2 = mmul(s, rinv, N, lane);
  const u256l gx = ld8(c_gx, lane), gy = ld8(c_gy, lane);
  JPoint acc{0u, small(1u, lane), 0u};
#pragma unroll 1
  for (int bit = 255; bit >= 0; bit--) {
    acc = jdouble(acc, lane);
    uint32_t b1 = (oshfl(u1, bit >> 5) >> (bit & 31)) & 1u;
    uint32_t b2 = (oshfl(u2, bit >> 5) >> (bit & 31)) & 1u;
    if (b1) acc = jadd_affine(acc, gx, gy, lane);
    if (b2) acc = jadd_affine(acc, x, y, lane);
  }
  if (u_is_zero(acc.z)) return false;
  u256l e_pinv = u_sub(P.m, small(2u, lane), lane, bo);
  u256l zi = mod_pow(acc.z, e_pinv, P.m, P.c, lane);
  u256l zi2 = mmul(zi, zi, P, lane);
  u256l qx = mmul(acc.x, zi2, P, lane), qy = mmul(acc.y, mmul(zi2, zi, P, lane), P, lane);
  // keccak256 of the 64-byte big-endian (Qx || Qy): one rate block of eight u64 words; word i = bytes [8 i, 8 i + 8)
  // little-endian goes to column x = i % 5, row y = i / 5 (keccak.cuh layout: octet lane x holds column x)
  KeccakState st;
#pragma unroll
  for (int y = 0; y < 5; y++) st.a[y] = 0;
  {
    // word i of the message: i < 4 from Qx (limbs 7 - 2i, 6 - 2i), else from Qy (limbs 7 - 2(i - 4), 6 - 2(i - 4))
    const int i0 = (int)lane, i1 = (int)lane + 5;  // rows 0 and 1 of this column
    uint32_t xl = oshfl(qx, (7 - 2 * i0) & 7), xh = oshfl(qx, (6 - 2 * i0) & 7);
    uint32_t yl0 = oshfl(qy, (7 - 2 * (i0 - 4)) & 7), yh0 = oshfl(qy, (6 - 2 * (i0 - 4)) & 7);
    uint32_t yl1 = oshfl(qy, (7 - 2 * (i1 - 4)) & 7), yh1 = oshfl(qy, (6 - 2 * (i1 - 4)) & 7);
    if (lane < 4) st.a[0] = (uint64_t)bswap32(xh) << 32 | bswap32(xl);
    else if (lane == 4) st.a[0] = (uint64_t)bswap32(yh0) << 32 | bswap32(yl0);
    if (lane < 3) st.a[1] = (uint64_t)bswap32(yh1) << 32 | bswap32(yl1);
    if (lane == 3) st.a[1] = 0x01ull;            // word 8: pad10*1 with the keccak domain byte
    if (lane == 1) st.a[3] = 0x80ull << 56;      // word 16: last byte of the rate
  }
  keccak_f1600(st, keccak_lanes(lane), ks, lane);
  // digest = words 0..3 = row 0 of columns 0..3, read as one big-endian 256-bit word
  int t = 7 - (int)lane;
  uint32_t lo = oshfl((uint32_t)st.a[0], (t >> 1) & 7), hi = oshfl((uint32_t)(st.a[0] >> 32), (t >> 1) & 7);
  u256l digest = bswap32((t & 1) ? hi : lo);
  address = lane < 5 ? digest : 0u;  // the low 20 bytes
  return true;
}

}  // namespace secp
}  // namespace zkb
