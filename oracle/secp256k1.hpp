// TEST INFRASTRUCTURE — part of the CPU oracle. secp256k1 public-key recovery for the ecrecover precompile.
//
// The reference delegates to the external `DefaultPrecompilesProcessor` (zk_evm_abstractions@v1.4.1, absent from
// /root/reference), which calls k256's recover-from-prehash.  Restated here from the published algorithm (SEC 1 v2
// §4.1.6): R = (x = r, y with parity v), Q = r^-1 (s R - z G), address = keccak256(Qx || Qy)[12..].  Known-answer
// vectors: src/testing/tests/precompiles/ecrecover.rs:127-143 (tests/golden/hash_vectors.json).
// Scalar code, textbook formulas (affine <-> Jacobian, plain double-and-add): written for clarity, independent of the
// warp-cooperative device implementation (era_zk_evm_b200/csrc/secp256k1.cuh).
#pragma once
#include "u256.hpp"

namespace orc_secp {

static const U256 P = {{0xFFFFFFFEFFFFFC2Full, 0xFFFFFFFFFFFFFFFFull, 0xFFFFFFFFFFFFFFFFull, 0xFFFFFFFFFFFFFFFFull}};
static const U256 N = {{0xBFD25E8CD0364141ull, 0xBAAEDCE6AF48A03Bull, 0xFFFFFFFFFFFFFFFEull, 0xFFFFFFFFFFFFFFFFull}};
static const U256 GX = {{0x59F2815B16F81798ull, 0x029BFCDB2DCE28D9ull, 0x55A06295CE870B07ull, 0x79BE667EF9DCBBACull}};
static const U256 GY = {{0x9C47D08FFB10D4B8ull, 0xFD17B448A6855419ull, 0x5DA4FBFC0E1108A8ull, 0x483ADA7726A3C465ull}};

inline U256 mod_add(const U256& a, const U256& b, const U256& m) {
  bool of;
  U256 r = u256_add(a, b, &of);
  if (of || u256_cmp(r, m) >= 0) {
    bool bo;
    r = u256_sub(r, m, &bo);
  }
  return r;
}
inline U256 mod_sub(const U256& a, const U256& b, const U256& m) {
  bool bo;
  U256 r = u256_sub(a, b, &bo);
  if (bo) {
    bool of;
    r = u256_add(r, m, &of);
  }
  return r;
}
// (a * b) mod m by binary long division of the 512-bit product (slow and obviously correct)
inline U256 mod_mul(const U256& a, const U256& b, const U256& m) {
  uint64_t t[8];
  u256_full_mul(a, b, t);
  U256 r = U256::zero();
  for (int bit = 511; bit >= 0; bit--) {
    bool top = r.w[3] >> 63;
    // r = 2 r + bit
    for (int i = 3; i > 0; i--) r.w[i] = (r.w[i] << 1) | (r.w[i - 1] >> 63);
    r.w[0] = (r.w[0] << 1) | ((t[bit / 64] >> (bit % 64)) & 1);
    if (top || u256_cmp(r, m) >= 0) {
      bool bo;
      r = u256_sub(r, m, &bo);
    }
  }
  return r;
}
inline U256 mod_pow(U256 base, const U256& e, const U256& m) {
  U256 r = U256::from_u64(1);
  for (int bit = 0; bit < 256; bit++) {
    if ((e.w[bit / 64] >> (bit % 64)) & 1) r = mod_mul(r, base, m);
    base = mod_mul(base, base, m);
  }
  return r;
}
inline U256 mod_inv(const U256& a, const U256& m) {  // m prime
  bool bo;
  return mod_pow(a, u256_sub(m, U256::from_u64(2), &bo), m);
}

struct Point {
  U256 x, y;
  bool inf;
};

inline Point add(const Point& a, const Point& b) {  // affine chord-and-tangent
  if (a.inf) return b;
  if (b.inf) return a;
  U256 lambda;
  if (a.x == b.x) {
    if (!(a.y == b.y) || a.y.is_zero()) return Point{U256::zero(), U256::zero(), true};
    U256 xx = mod_mul(a.x, a.x, P);
    U256 num = mod_add(mod_add(xx, xx, P), xx, P);
    lambda = mod_mul(num, mod_inv(mod_add(a.y, a.y, P), P), P);
  } else {
    lambda = mod_mul(mod_sub(b.y, a.y, P), mod_inv(mod_sub(b.x, a.x, P), P), P);
  }
  U256 x3 = mod_sub(mod_sub(mod_mul(lambda, lambda, P), a.x, P), b.x, P);
  U256 y3 = mod_sub(mod_mul(lambda, mod_sub(a.x, x3, P), P), a.y, P);
  return Point{x3, y3, false};
}

// Jacobian double-and-add to keep the oracle usable in tests (affine adds would need an inversion per step)
struct JPoint {
  U256 x, y, z;  // z == 0 <=> infinity
};
inline JPoint jdouble(const JPoint& p) {
  if (p.z.is_zero() || p.y.is_zero()) return JPoint{U256::zero(), U256::from_u64(1), U256::zero()};
  U256 a = mod_mul(p.x, p.x, P), b = mod_mul(p.y, p.y, P), c = mod_mul(b, b, P);
  U256 xb = mod_add(p.x, b, P);
  U256 d = mod_sub(mod_sub(mod_mul(xb, xb, P), a, P), c, P);
  d = mod_add(d, d, P);
  U256 e = mod_add(mod_add(a, a, P), a, P), f = mod_mul(e, e, P);
  U256 x3 = mod_sub(f, mod_add(d, d, P), P);
  U256 c8 = mod_add(c, c, P);
  c8 = mod_add(c8, c8, P);
  c8 = mod_add(c8, c8, P);
  U256 y3 = mod_sub(mod_mul(e, mod_sub(d, x3, P), P), c8, P);
  U256 yz = mod_mul(p.y, p.z, P);
  return JPoint{x3, y3, mod_add(yz, yz, P)};
}
inline JPoint jadd_affine(const JPoint& p, const Point& q) {  // q affine, not infinity
  if (p.z.is_zero()) return JPoint{q.x, q.y, U256::from_u64(1)};
  U256 zz = mod_mul(p.z, p.z, P);
  U256 u2 = mod_mul(q.x, zz, P), s2 = mod_mul(q.y, mod_mul(p.z, zz, P), P);
  U256 h = mod_sub(u2, p.x, P), r = mod_sub(s2, p.y, P);
  if (h.is_zero()) {
    if (r.is_zero()) return jdouble(p);
    return JPoint{U256::zero(), U256::from_u64(1), U256::zero()};
  }
  U256 hh = mod_mul(h, h, P), hhh = mod_mul(h, hh, P), v = mod_mul(p.x, hh, P);
  U256 x3 = mod_sub(mod_sub(mod_mul(r, r, P), hhh, P), mod_add(v, v, P), P);
  U256 y3 = mod_sub(mod_mul(r, mod_sub(v, x3, P), P), mod_mul(p.y, hhh, P), P);
  return JPoint{x3, y3, mod_mul(p.z, h, P)};
}
inline Point to_affine(const JPoint& p) {
  if (p.z.is_zero()) return Point{U256::zero(), U256::zero(), true};
  U256 zi = mod_inv(p.z, P), zi2 = mod_mul(zi, zi, P);
  return Point{mod_mul(p.x, zi2, P), mod_mul(p.y, mod_mul(zi2, zi, P), P), false};
}
inline Point scalar_mul(const U256& k, const Point& q) {
  JPoint acc{U256::zero(), U256::from_u64(1), U256::zero()};
  for (int bit = 255; bit >= 0; bit--) {
    acc = jdouble(acc);
    if ((k.w[bit / 64] >> (bit % 64)) & 1) acc = jadd_affine(acc, q);
  }
  return to_affine(acc);
}

// returns false when no key can be recovered (r or s out of [1, n-1], x^3 + 7 not a square, Q = infinity)
inline bool recover(const U256& hash, const U256& r, const U256& s, bool v_odd, U256* qx, U256* qy) {
  if (r.is_zero() || s.is_zero() || u256_cmp(r, N) >= 0 || u256_cmp(s, N) >= 0) return false;
  const U256& x = r;  // r < n < p
  U256 rhs = mod_add(mod_mul(mod_mul(x, x, P), x, P), U256::from_u64(7), P);
  // p = 3 mod 4: sqrt = rhs^((p+1)/4)
  U256 e = P;
  bool of;
  e = u256_add(e, U256::from_u64(1), &of);  // p + 1 does not overflow 256 bits
  for (int i = 0; i < 2; i++) {             // >> 2
    for (int k = 0; k < 3; k++) e.w[k] = (e.w[k] >> 1) | (e.w[k + 1] << 63);
    e.w[3] >>= 1;
  }
  U256 y = mod_pow(rhs, e, P);
  if (!(mod_mul(y, y, P) == rhs)) return false;
  if ((y.w[0] & 1) != (v_odd ? 1u : 0u)) {
    bool bo;
    y = u256_sub(P, y, &bo);
  }
  U256 z = hash;
  while (u256_cmp(z, N) >= 0) {
    bool bo;
    z = u256_sub(z, N, &bo);
  }
  U256 rinv = mod_inv(r, N);
  U256 u1 = mod_mul(mod_sub(U256::zero(), z, N), rinv, N), u2 = mod_mul(s, rinv, N);
  Point q = add(scalar_mul(u1, Point{GX, GY, false}), scalar_mul(u2, Point{x, y, false}));
  if (q.inf) return false;
  *qx = q.x;
  *qy = q.y;
  return true;
}

}  // namespace orc_secp
