// K14: integer-pipe micro-benchmarks (BASELINE north_star: "integer-pipe utilisation for the U256 ALU against sm_100a
// peak").  Two kinds of kernels, launched by zkb_alu_microbench (zkb.cu) at full occupancy (8 CTAs x 256 threads per SM):
//   * the MEASURED integer peaks of this part: a chain-parallel loop of mad.lo.u32 (SASS IMAD, the fma pipe) and one of
//     add.u32 / lop3.b32 (SASS IADD3 / LOP3, the alu pipe), eight independent chains per thread;
//   * the octet-distributed U256 primitives of u256.cuh exactly as the interpreter uses them (u_add / u_sub / u_mul /
//     u_divmod / u_shl): one U256 operation per octet and iteration, operands fed back so nothing is hoisted.
// The ratio (U256 ops/s x the limb operations one op needs) / (measured peak) is the integer-pipe utilisation of the ALU
// in isolation; ncu's sm__inst_executed_pipe_{alu,fma} on `bench.py --workload alu_loop` gives it inside the interpreter.
#pragma once
#include <stdint.h>

#include "u256.cuh"

namespace zkb {

enum { ALUB_IMAD = 0, ALUB_IADD3 = 1, ALUB_LOP3 = 2, ALUB_U256_ADD = 3, ALUB_U256_SUB = 4, ALUB_U256_MUL = 5, ALUB_U256_DIV = 6, ALUB_U256_SHL = 7, ALUB_N = 8 };
#define ALUB_CHAINS 8
#define ALUB_UNROLL 16

template <int OP>
__global__ void __launch_bounds__(256) zkb_alubench_kernel(uint32_t iters, uint32_t seed, uint32_t* sink) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (OP <= ALUB_LOP3) {
    uint32_t x[ALUB_CHAINS];
#pragma unroll
    for (int k = 0; k < ALUB_CHAINS; k++) x[k] = tid * 2654435761u + seed + k;
    const uint32_t a = seed | 1u, b = seed * 3u + 7u;
    for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
      for (int u = 0; u < ALUB_UNROLL; u++) {
#pragma unroll
        for (int k = 0; k < ALUB_CHAINS; k++) {
          if (OP == ALUB_IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[k]) : "r"(a), "r"(b));
          else if (OP == ALUB_IADD3) asm volatile("add.u32 %0, %0, %1;\n\tadd.u32 %0, %0, %2;" : "+r"(x[k]) : "r"(a), "r"(b));   // = ONE three-input IADD3
          else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[k]) : "r"(a), "r"(b));
        }
      }
    }
    uint32_t r = 0;
#pragma unroll
    for (int k = 0; k < ALUB_CHAINS; k++) r ^= x[k];
    if (r == 0x12345678u) sink[tid & 1023u] = r;   // (never true for all threads: keeps the chains alive)
  } else {
    const uint32_t lane = oct_lane();
    // a: 256-bit, b: 128-bit significant (the shape of token amounts / moduli); both differ per octet
    u256l a = (tid >> 3) * 0x9E3779B9u + lane * 0x85EBCA6Bu + seed;
    u256l b = lane < 4 ? ((tid >> 3) * 0xC2B2AE35u + lane * 0x27D4EB2Fu + (seed | 1u)) : 0u;
    uint32_t acc = 0;
    for (uint32_t it = 0; it < iters; it++) {
      if (OP == ALUB_U256_ADD) {
        bool of;
        a = u_add(a, b, lane, of);
        acc += of ? 1u : 0u;
      } else if (OP == ALUB_U256_SUB) {
        bool of;
        a = u_sub(a, b, lane, of);
        acc += of ? 1u : 0u;
      } else if (OP == ALUB_U256_MUL) {
        u256l lo, hi;
        u_mul(a, b, lane, lo, hi);
        a = lo ^ hi ^ 0x5bd1e995u;
      } else if (OP == ALUB_U256_DIV) {
        u256l q, r;
        u_divmod(a, b, lane, q, r);
        a = (q ^ r) + 0x9E3779B9u * (lane + 1u);   // back to a full-width dividend
      } else {
        a = u_shl(a, (acc & 63u) + 1u, lane) | (lane == 0 ? 1u : 0u);
        acc += a;
      }
    }
    if (oballot((a ^ acc) == 0x9abcdef0u) == 0xFFu) sink[tid & 1023u] = a;   // (practically never: keeps the chain alive)
  }
}

}  // namespace zkb
