// Octet-cooperative 256-bit integer arithmetic for sm_100a.
//
// Representation: a U256 is ONE 32-bit register per lane of an OCTET = 8 consecutive lanes of a warp; octet lane l
// (0..7) holds limb l (little-endian, limb 0 = least significant 32 bits).  A warp carries FOUR independent U256
// contexts (four VMs, one per octet); every cross-lane primitive below names only the calling octet in its member mask
// (the cooperative-groups tile<8> pattern), so the four octets may execute the same instruction together (converged:
// one issue slot serves four VMs) or different code (diverged: each octet syncs only with itself).
// Carries are resolved with two octet votes and one integer add (generate/propagate trick) instead of an 8-step ripple,
// so ADD/SUB cost ~10 warp instructions for four VMs.
//
// Semantics replaced (reference, ethereum_types::U256 as used in /root/reference/src/opcodes/execution/):
//   u_add  -> overflowing_add  add.rs:35        u_sub -> overflowing_sub  sub.rs:35
//   u_mul  -> full_mul         mul.rs:35-39     u_divmod -> div_mod       div.rs:50
//   u_shl/u_shr (n >= 256 -> 0) shift.rs:49-58, uma.rs:299-361
#pragma once
#include <stdint.h>

#define ZK_FULL 0xffffffffu
#define ZK_OCT 8u             // lanes per VM
#define ZK_VMS_PER_WARP 4u

namespace zkb {

typedef uint32_t u256l;  // the per-lane limb of an octet-distributed U256

__device__ __forceinline__ uint32_t oct_lane() { return threadIdx.x & 7u; }
__device__ __forceinline__ uint32_t oct_shift() { return threadIdx.x & 24u; }
__device__ __forceinline__ uint32_t oct_mask() { return 0xFFu << oct_shift(); }
__device__ __forceinline__ uint32_t oct_index() { return (threadIdx.x >> 3) & 3u; }  // octet within the warp
// value of octet lane `src` (taken modulo 8)
__device__ __forceinline__ uint32_t oshfl(uint32_t v, int src) { return __shfl_sync(oct_mask(), v, src, 8); }
__device__ __forceinline__ uint32_t oshfl_xor(uint32_t v, int m) { return __shfl_xor_sync(oct_mask(), v, m, 8); }
// 8-bit vote of the octet (bit l = octet lane l)
__device__ __forceinline__ uint32_t oballot(bool p) { return __ballot_sync(oct_mask(), p) >> oct_shift(); }
__device__ __forceinline__ bool oall(bool p) { return __all_sync(oct_mask(), p); }
__device__ __forceinline__ void osync() { __syncwarp(oct_mask()); }

__device__ __forceinline__ bool u_is_zero(u256l v) { return oballot(v != 0) == 0; }
__device__ __forceinline__ bool u_eq(u256l a, u256l b) { return oballot(a != b) == 0; }

// carry-in vector from generate/propagate votes: bit i = carry into limb i (bit 0 = cin), bit 8 = carry out
__device__ __forceinline__ uint32_t carry_chain(uint32_t G, uint32_t P, uint32_t cin = 0u) {
  uint32_t Gs = (G << 1) | cin;
  return Gs | ((P + Gs) ^ P ^ Gs);
}

// a + b + cin over the octet's 8 limbs; returns the sum limb, sets carry-out
__device__ __forceinline__ u256l u_addc(u256l a, u256l b, uint32_t lane, uint32_t cin, bool& carry_out) {
  uint32_t s = a + b;
  uint32_t G = oballot(s < a);
  uint32_t P = oballot(s == 0xFFFFFFFFu);
  uint32_t K = carry_chain(G, P, cin);
  carry_out = (K >> 8) & 1u;
  return s + ((K >> lane) & 1u);
}

__device__ __forceinline__ u256l u_add(u256l a, u256l b, uint32_t lane, bool& of) { return u_addc(a, b, lane, 0u, of); }

__device__ __forceinline__ u256l u_sub(u256l a, u256l b, uint32_t lane, bool& borrow_out) {
  uint32_t d = a - b;
  uint32_t G = oballot(a < b);
  uint32_t P = oballot(a == b);
  uint32_t K = carry_chain(G, P);
  borrow_out = (K >> 8) & 1u;
  return d - ((K >> lane) & 1u);
}

// unsigned compare: -1, 0, +1 (the most significant differing limb decides)
__device__ __forceinline__ int u_cmp(u256l a, u256l b) {
  uint32_t gt = oballot(a > b);
  uint32_t lt = oballot(a < b);
  return gt > lt ? 1 : (gt < lt ? -1 : 0);
}

// 256 x 256 -> 512 schoolbook product on 8 lanes: octet lane l accumulates column l (products a_i b_j with
// i + j = l, i <= l) and column l + 8 (i + j = l + 8, i > l) -- both use the same b limb (j = (l - i) mod 8), so one
// pair of shuffles feeds both columns.  The 96-bit column sums are then folded with four carry-resolved 8-limb adds.
__device__ __forceinline__ void u_mul(u256l a, u256l b, uint32_t lane, u256l& lo, u256l& hi) {
  uint64_t accL = 0, accH = 0;
  uint32_t topL = 0, topH = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    uint32_t ai = oshfl(a, i);
    uint32_t bj = oshfl(b, ((int)lane - i) & 7);
    uint64_t prod = (uint64_t)ai * (uint64_t)bj;
    bool low = (uint32_t)i <= lane;
    uint64_t pl = low ? prod : 0ull, ph = low ? 0ull : prod;
    accL += pl;
    topL += (accL < pl) ? 1u : 0u;
    accH += ph;
    topH += (accH < ph) ? 1u : 0u;
  }
  // column k contributes c0 at limb k, c1 at limb k + 1, c2 at limb k + 2
  const uint32_t c0L = (uint32_t)accL, c1L = (uint32_t)(accL >> 32), c0H = (uint32_t)accH, c1H = (uint32_t)(accH >> 32);
  const uint32_t t1L = oshfl(c1L, ((int)lane - 1) & 7), t1H = oshfl(c1H, ((int)lane - 1) & 7);
  const uint32_t t2L = oshfl(topL, ((int)lane - 2) & 7), t2H = oshfl(topH, ((int)lane - 2) & 7);
  const uint32_t yL = lane >= 1 ? t1L : 0u, yH = lane >= 1 ? t1H : t1L;
  const uint32_t zL = lane >= 2 ? t2L : 0u, zH = lane >= 2 ? t2H : t2L;
  bool c, dummy;
  uint32_t rL = u_addc(c0L, yL, lane, 0u, c);
  uint32_t rH = u_addc(c0H, yH, lane, c ? 1u : 0u, dummy);
  rL = u_addc(rL, zL, lane, 0u, c);
  rH = u_addc(rH, zH, lane, c ? 1u : 0u, dummy);
  lo = rL;
  hi = rH;
}

// logical shifts by n bits; n >= 256 yields 0 (ethereum_types semantics)
__device__ __forceinline__ u256l u_shl(u256l v, uint32_t n, uint32_t lane) {
  uint32_t k = n >> 5, b = n & 31u;
  int s0 = (int)lane - (int)k, s1 = s0 - 1;
  uint32_t x0 = oshfl(v, s0 & 7);
  uint32_t x1 = oshfl(v, s1 & 7);
  x0 = (s0 >= 0 && s0 < 8) ? x0 : 0u;
  x1 = (s1 >= 0 && s1 < 8) ? x1 : 0u;
  uint32_t r = __funnelshift_l(x1, x0, b);
  return n < 256 ? r : 0u;
}

__device__ __forceinline__ u256l u_shr(u256l v, uint32_t n, uint32_t lane) {
  uint32_t k = n >> 5, b = n & 31u;
  uint32_t s0 = lane + k, s1 = s0 + 1;
  uint32_t x0 = oshfl(v, s0 & 7);
  uint32_t x1 = oshfl(v, s1 & 7);
  x0 = s0 < 8 ? x0 : 0u;
  x1 = s1 < 8 ? x1 : 0u;
  uint32_t r = __funnelshift_r(x0, x1, b);
  return n < 256 ? r : 0u;
}

// number of significant bits (0 for zero)
__device__ __forceinline__ uint32_t u_bits(u256l v) {
  uint32_t nz = oballot(v != 0);
  if (nz == 0) return 0;
  int top = 31 - __clz(nz);
  uint32_t tv = oshfl(v, top);
  return (uint32_t)top * 32u + (32u - (uint32_t)__clz(tv));
}

// div_mod for b != 0: Knuth's algorithm D on 32-bit limbs, octet-distributed (round 1 used a bit-serial shift-subtract:
// up to 256 steps of two votes + a subtraction; this is at most 8 steps).  Normalise so that the divisor's top limb has
// its high bit set; per quotient limb j (from the top): estimate qhat from the two leading limbs of the running
// remainder, subtract qhat * (divisor << 32 j) with one carry-resolved add (assembling the product) and one
// borrow-resolved subtract, add the divisor back at most twice (the estimate from one divisor limb overshoots by <= 2).
// The running remainder has 9 limbs: 8 in the lanes plus `top` (octet-uniform).
__device__ __forceinline__ void u_divmod(u256l a, u256l b, uint32_t lane, u256l& q, u256l& r) {
  const uint32_t nza = oballot(a != 0), nzb = oballot(b != 0);
  q = 0;
  const int n = 32 - __clz(nzb);                  // significant limbs of b (>= 1: b != 0)
  const int la = nza ? 32 - __clz(nza) : 0;       // significant limbs of a
  if (la < n) {  // fewer limbs: a < b
    r = a;
    return;
  }
  const uint32_t s = __clz(oshfl(b, n - 1));      // normalisation shift (0..31)
  // v = b << s (fits: the top limb's leading zeros are shifted out), u = a << s (9 limbs: lanes + top)
  const uint32_t b_dn = oshfl(b, ((int)lane + 7) & 7), a_dn = oshfl(a, ((int)lane + 7) & 7);
  const uint32_t v = __funnelshift_l(lane ? b_dn : 0u, b, s);
  uint32_t u = __funnelshift_l(lane ? a_dn : 0u, a, s);
  uint32_t top = s ? (oshfl(a, 7) >> (32 - s)) : 0u;
  const uint32_t vtop = oshfl(v, n - 1);
  for (int j = la - n; j >= 0; j--) {
    // leading two limbs of the remainder window: u[j + n] (= top when j + n == 8), u[j + n - 1]
    const uint32_t uh = (j + n == 8) ? top : oshfl(u, (j + n) & 7), ul = oshfl(u, (j + n - 1) & 7);
    const uint64_t num = (uint64_t)uh << 32 | ul;
    uint32_t qhat = uh >= vtop ? 0xFFFFFFFFu : (uint32_t)(num / vtop);
    // P = qhat * (v << 32 j): lane l multiplies limb v[l - j]; product limb i = lo_i + hi_(i-1)
    const uint32_t vs = oshfl(v, ((int)lane - j) & 7);
    const uint32_t vj = (int)lane >= j ? vs : 0u;
    const uint64_t p = (uint64_t)qhat * vj;
    const uint32_t p_lo = (uint32_t)p, p_hi = (uint32_t)(p >> 32);
    const uint32_t hi_dn = oshfl(p_hi, ((int)lane + 7) & 7);
    bool c;
    const uint32_t P = u_addc(p_lo, lane ? hi_dn : 0u, lane, 0u, c);
    const uint32_t P8 = oshfl(p_hi, 7) + (c ? 1u : 0u);     // limb 8 of the product (cannot overflow: qhat * v < 2^288)
    bool bo;
    uint32_t d = u_sub(u, P, lane, bo);
    const uint64_t t = (uint64_t)top - P8 - (bo ? 1u : 0u);
    uint32_t dtop = (uint32_t)t;
    bool neg = (t >> 63) != 0;
    while (neg) {  // qhat was too large (at most twice): add the shifted divisor back
      qhat--;
      bool cy;
      d = u_addc(d, vj, lane, 0u, cy);
      const uint64_t t2 = (uint64_t)dtop + (cy ? 1u : 0u);
      dtop = (uint32_t)t2;
      neg = (t2 >> 32) == 0;   // still negative until the add carries out of limb 8
    }
    u = d;
    top = dtop;
    if ((int)lane == j) q = qhat;
  }
  // remainder = u >> s (top is zero by now)
  const uint32_t u_up = oshfl(u, (lane + 1) & 7);
  r = __funnelshift_r(u, lane < 7 ? u_up : 0u, s);
}

__device__ __forceinline__ uint32_t bswap32(uint32_t x) { return __byte_perm(x, 0, 0x0123); }

}  // namespace zkb
