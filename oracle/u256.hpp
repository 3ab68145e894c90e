// TEST INFRASTRUCTURE — part of the CPU oracle (see oracle/zkvm_oracle.cpp header). Not linked into the product.
//
// 256-bit unsigned integer with the semantics of ethereum_types::U256 as used by the reference
// (4 x u64 little-endian limbs; /root/reference/src/opcodes/execution/{add,sub,mul,div,shift}.rs).
#pragma once
#include <cstdint>
#include <cstring>

struct U256 {
  uint64_t w[4];
  static U256 zero() { return U256{{0, 0, 0, 0}}; }
  static U256 from_u64(uint64_t v) { return U256{{v, 0, 0, 0}}; }
  static U256 from_u128(uint64_t lo, uint64_t hi) { return U256{{lo, hi, 0, 0}}; }
  bool is_zero() const { return (w[0] | w[1] | w[2] | w[3]) == 0; }
  uint32_t low_u32() const { return (uint32_t)w[0]; }
  uint64_t low_u64() const { return w[0]; }
  bool operator==(const U256& o) const { return w[0] == o.w[0] && w[1] == o.w[1] && w[2] == o.w[2] && w[3] == o.w[3]; }
  bool operator!=(const U256& o) const { return !(*this == o); }
  // utils.rs:36-48 (U256::from_big_endian / to_big_endian)
  static U256 from_be(const uint8_t* b) {
    U256 r;
    for (int i = 0; i < 4; i++) {
      uint64_t v = 0;
      for (int j = 0; j < 8; j++) v = (v << 8) | b[(3 - i) * 8 + j];
      r.w[i] = v;
    }
    return r;
  }
  void to_be(uint8_t* b) const {
    for (int i = 0; i < 4; i++)
      for (int j = 0; j < 8; j++) b[(3 - i) * 8 + j] = (uint8_t)(w[i] >> (56 - 8 * j));
  }
  void to_limbs32(uint32_t* out) const {
    for (int i = 0; i < 4; i++) {
      out[2 * i] = (uint32_t)w[i];
      out[2 * i + 1] = (uint32_t)(w[i] >> 32);
    }
  }
  static U256 from_limbs32(const uint32_t* in) {
    U256 r;
    for (int i = 0; i < 4; i++) r.w[i] = (uint64_t)in[2 * i] | ((uint64_t)in[2 * i + 1] << 32);
    return r;
  }
};

inline int u256_cmp(const U256& a, const U256& b) {
  for (int i = 3; i >= 0; i--) {
    if (a.w[i] < b.w[i]) return -1;
    if (a.w[i] > b.w[i]) return 1;
  }
  return 0;
}

// add.rs:35 overflowing_add
inline U256 u256_add(const U256& a, const U256& b, bool* of) {
  U256 r;
  unsigned __int128 c = 0;
  for (int i = 0; i < 4; i++) {
    c += (unsigned __int128)a.w[i] + b.w[i];
    r.w[i] = (uint64_t)c;
    c >>= 64;
  }
  *of = c != 0;
  return r;
}

// sub.rs:35 overflowing_sub
inline U256 u256_sub(const U256& a, const U256& b, bool* of) {
  U256 r;
  uint64_t borrow = 0;
  for (int i = 0; i < 4; i++) {
    unsigned __int128 d = (unsigned __int128)a.w[i] - b.w[i] - borrow;
    r.w[i] = (uint64_t)d;
    borrow = (uint64_t)(d >> 64) & 1;
  }
  *of = borrow != 0;
  return r;
}

// mul.rs:35 full_mul -> 512 bits as 8 x u64
inline void u256_full_mul(const U256& a, const U256& b, uint64_t out[8]) {
  for (int i = 0; i < 8; i++) out[i] = 0;
  for (int i = 0; i < 4; i++) {
    unsigned __int128 carry = 0;
    for (int j = 0; j < 4; j++) {
      unsigned __int128 t = (unsigned __int128)a.w[i] * b.w[j] + out[i + j] + carry;
      out[i + j] = (uint64_t)t;
      carry = t >> 64;
    }
    out[i + 4] = (uint64_t)carry;
  }
}

// shift.rs:49-58: U256 << n, n >= 256 gives zero
inline U256 u256_shl(const U256& a, uint32_t n) {
  if (n >= 256) return U256::zero();
  U256 r = U256::zero();
  uint32_t ws = n / 64, bs = n % 64;
  for (int i = 3; i >= (int)ws; i--) {
    uint64_t v = a.w[i - ws] << bs;
    if (bs && i - (int)ws - 1 >= 0) v |= a.w[i - ws - 1] >> (64 - bs);
    r.w[i] = v;
  }
  return r;
}

inline U256 u256_shr(const U256& a, uint32_t n) {
  if (n >= 256) return U256::zero();
  U256 r = U256::zero();
  uint32_t ws = n / 64, bs = n % 64;
  for (int i = 0; i + ws < 4; i++) {
    uint64_t v = a.w[i + ws] >> bs;
    if (bs && i + ws + 1 < 4) v |= a.w[i + ws + 1] << (64 - bs);
    r.w[i] = v;
  }
  return r;
}

inline U256 u256_or(const U256& a, const U256& b) { return U256{{a.w[0] | b.w[0], a.w[1] | b.w[1], a.w[2] | b.w[2], a.w[3] | b.w[3]}}; }
inline U256 u256_and(const U256& a, const U256& b) { return U256{{a.w[0] & b.w[0], a.w[1] & b.w[1], a.w[2] & b.w[2], a.w[3] & b.w[3]}}; }
inline U256 u256_xor(const U256& a, const U256& b) { return U256{{a.w[0] ^ b.w[0], a.w[1] ^ b.w[1], a.w[2] ^ b.w[2], a.w[3] ^ b.w[3]}}; }

// div.rs:50 div_mod — Knuth algorithm D on 32-bit digits (b != 0)
inline void u256_div_mod(const U256& a, const U256& b, U256* q, U256* r) {
  uint32_t u[9], v[8], qd[8];
  a.to_limbs32(u);
  b.to_limbs32(v);
  int m = 8, n = 8;
  while (n > 0 && v[n - 1] == 0) n--;
  while (m > 0 && u[m - 1] == 0) m--;
  for (int i = 0; i < 8; i++) qd[i] = 0;
  if (m < n) {
    *q = U256::zero();
    *r = a;
    return;
  }
  if (n == 1) {
    uint64_t rem = 0;
    for (int i = m - 1; i >= 0; i--) {
      uint64_t cur = (rem << 32) | u[i];
      qd[i] = (uint32_t)(cur / v[0]);
      rem = cur % v[0];
    }
    *q = U256::from_limbs32(qd);
    *r = U256::from_u64(rem);
    return;
  }
  int s = __builtin_clz(v[n - 1]);
  uint32_t vn[8], un[9];
  for (int i = n - 1; i > 0; i--) vn[i] = (v[i] << s) | (s ? (uint32_t)((uint64_t)v[i - 1] >> (32 - s)) : 0);
  vn[0] = v[0] << s;
  un[m] = s ? (uint32_t)((uint64_t)u[m - 1] >> (32 - s)) : 0;
  for (int i = m - 1; i > 0; i--) un[i] = (u[i] << s) | (s ? (uint32_t)((uint64_t)u[i - 1] >> (32 - s)) : 0);
  un[0] = u[0] << s;
  for (int j = m - n; j >= 0; j--) {
    uint64_t num = ((uint64_t)un[j + n] << 32) | un[j + n - 1];
    uint64_t qhat = num / vn[n - 1];
    uint64_t rhat = num % vn[n - 1];
    while (qhat >= (1ull << 32) || qhat * vn[n - 2] > ((rhat << 32) | un[j + n - 2])) {
      qhat--;
      rhat += vn[n - 1];
      if (rhat >= (1ull << 32)) break;
    }
    int64_t borrow = 0;
    int64_t t;
    for (int i = 0; i < n; i++) {
      uint64_t p = qhat * vn[i];
      t = (int64_t)un[i + j] - borrow - (int64_t)(p & 0xFFFFFFFFull);
      un[i + j] = (uint32_t)t;
      borrow = (int64_t)(p >> 32) - (t >> 32);
    }
    t = (int64_t)un[j + n] - borrow;
    un[j + n] = (uint32_t)t;
    qd[j] = (uint32_t)qhat;
    if (t < 0) {
      qd[j]--;
      uint64_t carry = 0;
      for (int i = 0; i < n; i++) {
        uint64_t sum = (uint64_t)un[i + j] + vn[i] + carry;
        un[i + j] = (uint32_t)sum;
        carry = sum >> 32;
      }
      un[j + n] += (uint32_t)carry;
    }
  }
  uint32_t rd[8];
  for (int i = 0; i < 8; i++) rd[i] = 0;
  for (int i = 0; i < n; i++) rd[i] = (un[i] >> s) | (s && i + 1 <= m ? (uint32_t)((uint64_t)un[i + 1] << (32 - s)) : 0);
  *q = U256::from_limbs32(qd);
  *r = U256::from_limbs32(rd);
}
