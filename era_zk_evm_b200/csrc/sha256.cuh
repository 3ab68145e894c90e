// sha256 compression function for the sha256 precompile (external `DefaultPrecompilesProcessor`,
// zk_evm_abstractions@v1.4.1, selected by address 0x02 from /root/reference/src/vm_state/helpers.rs:211-213;
// hash vectors: src/testing/tests/precompiles/sha256.rs:119-136).
//
// One octet (8 lanes) = one VM: the 8 working variables are octet-uniform registers, the rolling 16-word message
// schedule lives in the VM's shared-memory scratch (broadcast reads, octet lane 0 writes) so the precompile adds no
// registers to the interpreter's hot loop; the four VMs of a warp run their compressions in the same issue slots.
// 64 rounds per 64-byte block, one block = two VM heap words.
#pragma once
#include <stdint.h>
#include "u256.cuh"

namespace zkb {

__constant__ uint32_t c_sha256_k[64] = {
    0x428a2f98u, 0x71374491u, 0xb5c0fbcfu, 0xe9b5dba5u, 0x3956c25bu, 0x59f111f1u, 0x923f82a4u, 0xab1c5ed5u, 0xd807aa98u, 0x12835b01u,
    0x243185beu, 0x550c7dc3u, 0x72be5d74u, 0x80deb1feu, 0x9bdc06a7u, 0xc19bf174u, 0xe49b69c1u, 0xefbe4786u, 0x0fc19dc6u, 0x240ca1ccu,
    0x2de92c6fu, 0x4a7484aau, 0x5cb0a9dcu, 0x76f988dau, 0x983e5152u, 0xa831c66du, 0xb00327c8u, 0xbf597fc7u, 0xc6e00bf3u, 0xd5a79147u,
    0x06ca6351u, 0x14292967u, 0x27b70a85u, 0x2e1b2138u, 0x4d2c6dfcu, 0x53380d13u, 0x650a7354u, 0x766a0abbu, 0x81c2c92eu, 0x92722c85u,
    0xa2bfe8a1u, 0xa81a664bu, 0xc24b8b70u, 0xc76c51a3u, 0xd192e819u, 0xd6990624u, 0xf40e3585u, 0x106aa070u, 0x19a4c116u, 0x1e376c08u,
    0x2748774cu, 0x34b0bcb5u, 0x391c0cb3u, 0x4ed8aa4au, 0x5b9cca4fu, 0x682e6ff3u, 0x748f82eeu, 0x78a5636fu, 0x84c87814u, 0x8cc70208u,
    0x90befffau, 0xa4506cebu, 0xbef9a3f7u, 0xc67178f2u};

__constant__ uint32_t c_sha256_iv[8] = {0x6a09e667u, 0xbb67ae85u, 0x3c6ef372u, 0xa54ff53au, 0x510e527fu, 0x9b05688cu, 0x1f83d9abu, 0x5be0cd19u};

__device__ __forceinline__ uint32_t rotr32(uint32_t x, uint32_t n) { return __funnelshift_r(x, x, n); }

// hw: shared-memory scratch of the VM; hw[0..7] = chaining value (updated in place), hw[8..23] = the 16 message
// words of the block (big-endian), clobbered.  All 8 lanes of the octet must call.
__device__ __noinline__ void sha256_compress_smem(uint32_t* hw, uint32_t lane) {
  uint32_t* w = hw + 8;
  osync();
  uint32_t a = hw[0], b = hw[1], c = hw[2], d = hw[3], e = hw[4], f = hw[5], g = hw[6], h = hw[7];
#pragma unroll 1
  for (int t = 0; t < 64; t++) {
    uint32_t wt;
    if (t < 16) {
      wt = w[t];
    } else {
      uint32_t w15 = w[(t - 15) & 15], w2 = w[(t - 2) & 15], w16 = w[t & 15], w7 = w[(t - 7) & 15];
      uint32_t s0 = rotr32(w15, 7) ^ rotr32(w15, 18) ^ (w15 >> 3);
      uint32_t s1 = rotr32(w2, 17) ^ rotr32(w2, 19) ^ (w2 >> 10);
      wt = w16 + s0 + w7 + s1;
      osync();
      if (lane == 0) w[t & 15] = wt;
      osync();
    }
    uint32_t S1 = rotr32(e, 6) ^ rotr32(e, 11) ^ rotr32(e, 25);
    uint32_t ch = (e & f) ^ (~e & g);
    uint32_t t1 = h + S1 + ch + c_sha256_k[t] + wt;
    uint32_t S0 = rotr32(a, 2) ^ rotr32(a, 13) ^ rotr32(a, 22);
    uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
    uint32_t t2 = S0 + mj;
    h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
  }
  osync();
  if (lane == 0) {
    hw[0] += a; hw[1] += b; hw[2] += c; hw[3] += d; hw[4] += e; hw[5] += f; hw[6] += g; hw[7] += h;
  }
  osync();
}

}  // namespace zkb
