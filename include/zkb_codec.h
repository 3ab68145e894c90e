/* zkb_codec.h — lossless TRANSPORT encoding of the six witness streams (host <-> device wire format).
 *
 * Why: the canonical records (zkb_records.h) are what the VmWitnessTracer callbacks carry
 * (/root/reference/src/witness_trace/mod.rs:11-72) -- 338 bytes per VM cycle on the ERC-20 workload, and a PCIe 5 x16
 * link moves ~55 GB/s, so the end-to-end rate of the batch API was the PCIe rate, not the kernel's.  Most of those bytes
 * are predictable from the previous record of the same VM (cycle / timestamp / pc counters, the unchanged frame tail
 * of a cycle row) or zero (short U256 operands, dst1, reserved words).  The encoder (zkb_encode_kernel, csrc/codec.cuh)
 * XORs every 32-bit word of a record with a prediction and sends a presence bitmap + the non-zero residual words; the
 * decoder below inverts it exactly.  Nothing is dropped: decode(encode(streams)) == streams byte for byte
 * (tests/test_codec.py, host side against the CPU oracle; -m gpu: the CUDA encoder's blob == the oracle encoder's).
 *
 * Blob layout (little-endian):
 *   ZkbEncodedHeader                                        128 bytes
 *   uint32 counts[n_vms][8]        per VM: records in each of the six streams, ZkbVmCode, cycles
 *   uint64 offsets[6][n_vms + 1]   byte offset of VM v's records inside stream k's payload
 *   payload of stream 0 .. 5       (each starts 16-byte aligned)
 * Record encodings (u32 words):
 *   ROWS      mask_lo, mask_hi, residual words with mask bit set, ascending word index     (64 words per record)
 *   MEM       mask (12 bits), residuals     LOG / FRAME   mask (32 bits), residuals     DECOMMIT  mask (12 bits), residuals
 *   REFUND    the two raw words
 * Predictions (prev = the previous record of the same VM and stream, all-zero before the first):
 *   ROWS   w0 cycle: prev + 1;  w1 timestamp: prev + TIME_DELTA_PER_CYCLE;  w2-3 raw opcode: 0;
 *          w4 variant | resolved << 16 | err << 24: (w2 & 0x7FF) | 1 << 16  (the unmasked, condition-true case);
 *          w5 pc_before | pc_after << 16: p | (p + 1) << 16 with p = prev.pc_after;  w6 sp | flags | bits: prev;
 *          w7 ergs_after: prev - OPCODES_PRICES[w2 & 0x7FF];  w8-39 operands: 0;  w43 per-cycle record counts: 0;
 *          every other word (the frame tail): prev
 *   MEM    w0 timestamp, w1 page, w3 type/flags: prev;  w2 index: prev + 1;  value: 0
 *   LOG    w0-7 (timestamp, tx, aux, shard, address, flags): prev;  key / read / written: 0
 *   DECOMMIT  w0-3: prev;  hash: 0            FRAME  every word: prev
 */
#ifndef ZKB_CODEC_H
#define ZKB_CODEC_H
#include <stdint.h>
#include <string.h>

#include "zkb_records.h"

#ifdef __cplusplus
extern "C" {
#endif

#define ZKB_CODEC_MAGIC 0x31424B5Au /* "ZKB1" */
#define ZKB_CODEC_VERSION 1u

typedef struct ZkbEncodedHeader {
  uint32_t magic, version, n_vms;
  uint32_t reserved0;                       /* 0 = all six streams are in the blob; else a bit mask of the ZkbStreamKinds that are */
  uint64_t total_bytes;                     /* whole blob */
  uint64_t raw_bytes;                       /* canonical bytes it decodes to (sum over streams) */
  uint64_t counts_offset, offsets_offset;   /* from the start of the blob */
  uint64_t payload_offset[ZKB_N_STREAMS];
  uint64_t payload_bytes[ZKB_N_STREAMS];
  uint64_t reserved1[2];
} ZkbEncodedHeader;

/* words per canonical record / per presence bitmap of each stream */
static const uint32_t ZKB_CODEC_REC_WORDS[ZKB_N_STREAMS] = {64, 12, 32, 12, 32, 2};
static const uint32_t ZKB_CODEC_MASK_WORDS[ZKB_N_STREAMS] = {2, 1, 1, 1, 1, 0};

#ifdef __cplusplus
}
#endif

#if defined(__cplusplus) && !defined(ZKB_CODEC_NO_HOST)
/* ---- host side: scalar predictor + decoder (+ the reference encoder used by the CPU oracle / tests) ---------------- */
#ifndef ZK_TABLE_QUALIFIER
#define ZK_TABLE_QUALIFIER static const
#endif
#include "../era_zk_evm_b200/csrc/isa_tables.inc"

namespace zkb_codec {

/* prediction of word i of a record of stream `kind`; cur[] holds the words of the SAME record with index < i (decoded
 * in ascending order), prev[] the previous record of that VM and stream (zeros before the first) */
static inline uint32_t predict(uint32_t kind, uint32_t i, const uint32_t* prev, const uint32_t* cur) {
  switch (kind) {
    case ZKB_STREAM_ROWS:
      if (i == 0) return prev[0] + 1u;
      if (i == 1) return prev[1] + ZK_TIME_DELTA_PER_CYCLE;
      if (i == 2 || i == 3) return 0u;
      if (i == 4) return (cur[2] & ((1u << ZK_VARIANT_BITS) - 1u)) | 1u << 16;
      if (i == 5) {
        const uint32_t p = prev[5] >> 16;
        return p | ((p + 1u) & 0xFFFFu) << 16;
      }
      if (i == 7) return prev[7] - ZK_OPCODE_PRICES[cur[2] & ((1u << ZK_VARIANT_BITS) - 1u)];
      if ((i >= 8 && i < 40) || i == 43) return 0u;
      return prev[i];
    case ZKB_STREAM_MEM:
      if (i == 2) return prev[2] + 1u;
      return i < 4 ? prev[i] : 0u;
    case ZKB_STREAM_LOG: return i < 8 ? prev[i] : 0u;
    case ZKB_STREAM_DECOMMIT: return i < 4 ? prev[i] : 0u;
    case ZKB_STREAM_FRAME: return prev[i];
    default: return 0u;
  }
}

/* encodes n records (canonical bytes at `src`) of one VM and stream; returns the number of bytes written to `dst`
 * (dst == NULL: size only).  The scalar restatement of zkb_encode_kernel. */
static inline uint64_t encode_records(uint32_t kind, const void* src, uint64_t n, uint8_t* dst) {
  const uint32_t nw = ZKB_CODEC_REC_WORDS[kind], mw = ZKB_CODEC_MASK_WORDS[kind];
  const uint32_t* in = (const uint32_t*)src;
  uint32_t prev[64];
  memset(prev, 0, sizeof(prev));
  uint64_t at = 0;
  for (uint64_t r = 0; r < n; r++, in += nw) {
    if (mw == 0) {
      if (dst) memcpy(dst + at, in, (size_t)nw * 4);
      at += (uint64_t)nw * 4;
      continue;
    }
    uint32_t resid[64];
    uint64_t mask = 0;
    uint32_t k = 0;
    for (uint32_t i = 0; i < nw; i++) {
      const uint32_t x = in[i] ^ predict(kind, i, prev, in);
      if (x) {
        mask |= 1ull << i;
        resid[k++] = x;
      }
    }
    if (dst) {
      const uint32_t m32[2] = {(uint32_t)mask, (uint32_t)(mask >> 32)};
      memcpy(dst + at, m32, (size_t)mw * 4);
      memcpy(dst + at + (size_t)mw * 4, resid, (size_t)k * 4);
    }
    at += ((uint64_t)mw + k) * 4;
    memcpy(prev, in, (size_t)nw * 4);
  }
  return at;
}

/* decodes n records of one VM and stream from `src` (n_src bytes) into canonical bytes at `dst`; returns the number of
 * encoded bytes consumed, or UINT64_MAX when the input is truncated.  decode_records_ref is the word-by-word inverse of
 * encode_records (one predict() call per word); decode_records below is what the library runs: the same function with
 * the predictions written as whole-record block operations and the residuals applied by walking the set bits of the
 * presence mask -- ~10x faster, checked against the reference decoder in tests/test_codec.py. */
static inline uint64_t decode_records_ref(uint32_t kind, const uint8_t* src, uint64_t n_src, uint64_t n, void* dst) {
  const uint32_t nw = ZKB_CODEC_REC_WORDS[kind], mw = ZKB_CODEC_MASK_WORDS[kind];
  uint32_t* out = (uint32_t*)dst;
  uint32_t zero[64];
  memset(zero, 0, sizeof(zero));
  const uint32_t* prev = zero;
  uint64_t at = 0;
  for (uint64_t r = 0; r < n; r++, out += nw) {
    if (mw == 0) {
      if (at + (uint64_t)nw * 4 > n_src) return UINT64_MAX;
      memcpy(out, src + at, (size_t)nw * 4);
      at += (uint64_t)nw * 4;
      continue;
    }
    if (at + (uint64_t)mw * 4 > n_src) return UINT64_MAX;
    uint32_t m32[2] = {0, 0};
    memcpy(m32, src + at, (size_t)mw * 4);
    at += (uint64_t)mw * 4;
    const uint64_t mask = (uint64_t)m32[0] | (uint64_t)m32[1] << 32;
    if (at + (uint64_t)__builtin_popcountll(mask) * 4 > n_src) return UINT64_MAX;
    for (uint32_t i = 0; i < nw; i++) {
      uint32_t x = 0;
      if ((mask >> i) & 1u) {
        memcpy(&x, src + at, 4);
        at += 4;
      }
      out[i] = x ^ predict(kind, i, prev, out);
    }
    prev = out;
  }
  return at;
}

/* XORs the residual words named by `mask` (ascending bit order) into out[]; returns the advanced read position */
static inline const uint8_t* apply_residuals(uint64_t mask, const uint8_t* p, uint32_t* out) {
  while (mask) {
    uint32_t x;
    memcpy(&x, p, 4);
    p += 4;
    out[__builtin_ctzll(mask)] ^= x;
    mask &= mask - 1;
  }
  return p;
}

static inline uint64_t decode_records(uint32_t kind, const uint8_t* src, uint64_t n_src, uint64_t n, void* dst) {
  const uint32_t nw = ZKB_CODEC_REC_WORDS[kind], mw = ZKB_CODEC_MASK_WORDS[kind];
  uint32_t* out = (uint32_t*)dst;
  if (mw == 0) {  /* REFUND: raw */
    const uint64_t need = n * nw * 4;
    if (need > n_src) return UINT64_MAX;
    memcpy(out, src, (size_t)need);
    return need;
  }
  uint32_t zero[64];
  memset(zero, 0, sizeof(zero));
  const uint32_t* prev = zero;
  const uint8_t* p = src;
  const uint8_t* const end = src + n_src;
  for (uint64_t r = 0; r < n; r++, out += nw) {
    if ((uint64_t)(end - p) < (uint64_t)mw * 4) return UINT64_MAX;
    uint32_t m32[2] = {0, 0};
    memcpy(m32, p, (size_t)mw * 4);
    p += (size_t)mw * 4;
    const uint64_t mask = (uint64_t)m32[0] | (uint64_t)m32[1] << 32;
    if ((uint64_t)(end - p) < (uint64_t)__builtin_popcountll(mask) * 4) return UINT64_MAX;
    switch (kind) {
      case ZKB_STREAM_ROWS: {
        /* every prediction that does not look at the record itself, as block operations */
        out[0] = prev[0] + 1u;
        out[1] = prev[1] + ZK_TIME_DELTA_PER_CYCLE;
        out[2] = out[3] = 0u;
        const uint32_t pc = prev[5] >> 16;
        out[5] = pc | ((pc + 1u) & 0xFFFFu) << 16;
        out[6] = prev[6];
        memset(out + 8, 0, 32 * 4);
        memcpy(out + 40, prev + 40, 24 * 4);
        out[43] = 0u;
        /* words 4 and 7 are predicted from the decoded raw opcode (word 2): residuals of words 0..3 first */
        p = apply_residuals(mask & 0xFull, p, out);
        const uint32_t v = out[2] & ((1u << ZK_VARIANT_BITS) - 1u);
        out[4] = v | 1u << 16;
        out[7] = prev[7] - ZK_OPCODE_PRICES[v];
        p = apply_residuals(mask & ~0xFull, p, out);
        break;
      }
      case ZKB_STREAM_MEM:
      case ZKB_STREAM_DECOMMIT:
        memcpy(out, prev, 16);
        if (kind == ZKB_STREAM_MEM) out[2] = prev[2] + 1u;
        memset(out + 4, 0, 32);
        p = apply_residuals(mask, p, out);
        break;
      case ZKB_STREAM_LOG:
        memcpy(out, prev, 32);
        memset(out + 8, 0, 96);
        p = apply_residuals(mask, p, out);
        break;
      default: /* FRAME */
        memcpy(out, prev, 128);
        p = apply_residuals(mask, p, out);
        break;
    }
    prev = out;
  }
  return (uint64_t)(p - src);
}

/* a received blob: validates the header, gives per-VM access */
struct EncodedView {
  const uint8_t* base = nullptr;
  const ZkbEncodedHeader* h = nullptr;
  bool open(const void* blob, uint64_t n_bytes) {
    if (!blob || n_bytes < sizeof(ZkbEncodedHeader)) return false;
    base = (const uint8_t*)blob;
    h = (const ZkbEncodedHeader*)blob;
    if (h->magic != ZKB_CODEC_MAGIC || h->version != ZKB_CODEC_VERSION || h->total_bytes > n_bytes) return false;
    return true;
  }
  uint32_t n_vms() const { return h->n_vms; }
  const uint32_t* counts(uint32_t vm) const { return (const uint32_t*)(base + h->counts_offset) + (size_t)vm * 8; }
  const uint64_t* offsets(uint32_t kind) const { return (const uint64_t*)(base + h->offsets_offset) + (size_t)kind * (h->n_vms + 1); }
  /* canonical bytes of VM `vm`'s stream `kind` -> dst (capacity max_bytes); returns the canonical length, or
   * UINT64_MAX on a malformed blob.  dst == NULL: length only. */
  uint64_t decode(uint32_t vm, uint32_t kind, void* dst, uint64_t max_bytes, bool reference_decoder = false) const {
    if (vm >= h->n_vms || kind >= ZKB_N_STREAMS) return UINT64_MAX;
    const bool present = h->reserved0 == 0 || ((h->reserved0 >> kind) & 1u);   /* a blob may carry a subset of the streams */
    const uint64_t n = present ? counts(vm)[kind] : 0, need = n * ZKB_CODEC_REC_WORDS[kind] * 4;
    if (!dst) return need;
    if (need > max_bytes) return UINT64_MAX;
    const uint64_t lo = offsets(kind)[vm], hi = offsets(kind)[vm + 1];
    if (hi < lo || hi > h->payload_bytes[kind]) return UINT64_MAX;
    const uint8_t* src = base + h->payload_offset[kind] + lo;
    const uint64_t used = reference_decoder ? decode_records_ref(kind, src, hi - lo, n, dst) : decode_records(kind, src, hi - lo, n, dst);
    return used == hi - lo ? need : UINT64_MAX;
  }
};

}  // namespace zkb_codec
#endif
#endif
