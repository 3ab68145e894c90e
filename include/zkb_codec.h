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
 * Record encodings (u32 words), format version 2:
 *   LOG / FRAME   mask (32 bits), residual words with mask bit set, ascending word index
 *   DECOMMIT      mask (12 bits), residuals            REFUND    the two raw words
 *   predictions (prev = the previous record of the same VM and stream, all-zero before the first):
 *     LOG w0-7 (timestamp, tx, aux, shard, address, flags): prev;  key / read: 0;  written: the same record's read value
 *     DECOMMIT  w0-3: prev;  hash: 0            FRAME  every word: prev
 *   ROWS + MEM are coded JOINTLY, cycle by cycle (row r, then the n_mem memory queries row r announces), because most of
 *   a memory query is a function of the cycle that issued it and the raw opcode of a cycle is a slice of the code word the
 *   cycle's instruction fetch returned (cycle.rs:59-100).  State per VM: the previous row, the previous memory query, the
 *   current code word `cw`, a 32-set x 4-way FIFO cache of code words seen so far (set = word index mod 32).
 *   ROWS   mask_lo, mask_hi, residuals.  mask bits 0..54 = presence of words 0..54 (words 55..63 are reserved-zero and not
 *          transmitted); bits 55-56 = dst0 predictor (0 zero, 1 src0, 2 src0 + src1, 3 src0 - src1, 256-bit); bits 57-59
 *          = w43 (per-cycle record counts) when it is < 7, else 7 and w43 travels as a residual.
 *          w0 cycle: prev + 1;  w1 timestamp: prev + TIME_DELTA_PER_CYCLE;  w2-3 raw opcode: cw[6 - 2 sub_pc], cw[7 - 2 sub_pc];
 *          w4 variant | resolved << 16 | err << 24: (w2 & 0x7FF) | 1 << 16;  w5 pc_before | pc_after << 16: p | (p + 1) << 16
 *          with p = prev.pc_after;  w6 sp | flags | bits: prev;  w7 ergs_after: prev - OPCODES_PRICES[w2 & 0x7FF];
 *          w8-23 src0, src1: 0;  w24-31 dst0: by the predictor bits;  w32-39 dst1: 0;  w48: prev.tx_number |
 *          (pc_before >> 2) << 16;  every other word (the frame tail): prev
 *   MEM    one word: bits 0..11 presence of w0..w11, bits 12-14 value predictor, bits 16-19 code of the flags word w3
 *          (index into ZKB_MEM_FLAG_CODES; 15 = escape: w3 predicted by prev and carried as a residual), then residuals.
 *          w0 timestamp: row.timestamp (+ 3 for a write);  w1 page: by memory type from the frame of the previous row
 *          (code page / base page + 1, 2, 3; fat pointer: src0 limb 1);  w2 index: previous query's + 1 inside a run of
 *          equal flags, else instruction fetch: pc_before >> 2, code / stack operand: imm0 / imm1 of the raw opcode,
 *          heap: src0 >> 5, fat pointer: (offset + start) >> 5;  value: zero | src0 | src1 | dst0 | dst1 of the row,
 *          or (code reads) one of the four cached code words of the index's set
 *   Order inside a cycle (the decoder's data dependencies): row words that do not depend on the opcode -> if an
 *   instruction fetch is expected (super-pc or code page moved) the cycle's FIRST memory query, which refreshes `cw` ->
 *   w2, w3, w4, w7 -> the remaining memory queries.  zkb_codec::JointCoder below is that walk, written once for both
 *   directions.
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
#define ZKB_CODEC_VERSION 2u

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
/* format v2, joint ROWS + MEM coding (see above): the flags words of a memory query that travel as a 4-bit code, the
 * geometry of the code-word cache, the number of row words that are transmitted */
#define ZKB_MEM_FLAG_CODES_INIT                                                                                                      \
  {0x00000004u, 0x00000001u, 0x00000101u, 0x00000003u, 0x00000000u, 0x00000100u, 0x01000003u, 0x02000101u, 0x00010000u, 0x00010100u, \
   0x00000002u, 0x00000102u, 0x01000001u, 0x02000102u, 0x00010003u}
enum { ZKB_CW_SETS = 32, ZKB_CW_WAYS = 4, ZKB_ROW_TX_WORDS = 55 /* words 55..63 of a row are reserved-zero */ };

#ifdef __cplusplus
}
#endif

#if defined(__cplusplus) && !defined(ZKB_CODEC_NO_HOST)
/* ---- host side: scalar predictor + decoder (+ the reference encoder used by the CPU oracle / tests) ---------------- */
#ifndef ZK_TABLE_QUALIFIER
#define ZK_TABLE_QUALIFIER static const
#endif
#include "../era_zk_evm_b200/csrc/isa_tables.inc"
#include <algorithm>
#include <vector>

namespace zkb_codec {

/* per-stream coding (LOG, DECOMMIT, FRAME; REFUND travels raw): prediction of word i of a record of stream `kind`; prev[]
 * is the previous record of that VM and stream (zeros before the first).  ROWS and MEM are coded by JointCoder below. */
static inline uint32_t predict(uint32_t kind, uint32_t i, const uint32_t* prev, const uint32_t* cur) {
  switch (kind) {
    case ZKB_STREAM_LOG: return i < 8 ? prev[i] : i >= 24 ? cur[i - 8] : 0u;   /* written_value: the read value (a read writes back what it read) */
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
    if (nw < 64 && (mask >> nw)) return UINT64_MAX;   /* presence bits beyond the record (a 12-word record carries a 32-bit mask word) */
    if ((uint64_t)(end - p) < (uint64_t)__builtin_popcountll(mask) * 4) return UINT64_MAX;
    switch (kind) {
      case ZKB_STREAM_DECOMMIT:
        memcpy(out, prev, 16);
        memset(out + 4, 0, 32);
        p = apply_residuals(mask, p, out);
        break;
      case ZKB_STREAM_LOG:
        memcpy(out, prev, 32);
        memset(out + 8, 0, 96);
        p = apply_residuals(mask, p, out);
        for (int i = 24; i < 32; i++) out[i] ^= out[i - 8];   /* written_value is predicted by the (now final) read value */
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

/* ---- format version 2: joint coding of the cycle rows and the memory queries of one VM ----------------------------- */
/* flags words (memory_type | rw << 8 | value_is_pointer << 16 | origin << 24) that travel as a 4-bit code */
static const uint32_t ZKB_MEM_FLAG_CODES[15] = ZKB_MEM_FLAG_CODES_INIT;

static inline void limbs_add(const uint32_t* a, const uint32_t* b, uint32_t* r) {
  uint64_t c = 0;
  for (int i = 0; i < 8; i++) {
    c += (uint64_t)a[i] + b[i];
    r[i] = (uint32_t)c;
    c >>= 32;
  }
}
static inline void limbs_sub(const uint32_t* a, const uint32_t* b, uint32_t* r) {
  uint64_t br = 0;
  for (int i = 0; i < 8; i++) {
    const uint64_t d = (uint64_t)a[i] - b[i] - br;
    r[i] = (uint32_t)d;
    br = d >> 63;
  }
}
static inline uint32_t count_diff8(const uint32_t* a, const uint32_t* b) {
  uint32_t n = 0;
  for (int i = 0; i < 8; i++) n += a[i] != b[i];
  return n;
}

/* The walk over one VM's rows and memory queries.  ENC = true: reads canonical records, writes the two payloads
 * (rdst / mdst may be NULL: sizes only).  ENC = false: reads the payloads, writes canonical records.  One body for both
 * directions, so a prediction can never differ between the encoder and the decoder; what is checked from outside is
 * decode(encode(x)) == x against the oracle's streams and CUDA blob == this encoder's blob (tests/test_codec.py). */
struct JointCoder {
  const uint32_t* prev_row;   /* the previous row / memory query, where they lie in the caller's canonical buffers (a zero record before the first) */
  const uint32_t* prev_mem;
  uint32_t cw[8], rr[ZKB_CW_SETS], cache[ZKB_CW_SETS][ZKB_CW_WAYS][8];
  uint32_t prevprev_w50;
  /* payload cursors */
  uint8_t* rp;        /* rows payload (written when ENC, read when !ENC) */
  uint8_t* mp;
  const uint8_t* rend;
  const uint8_t* mend;
  uint64_t rbytes, mbytes;
  bool ok;

  void reset() {
    static const uint32_t zero_record[64] = {0};
    prev_row = prev_mem = zero_record;
    memset(cw, 0, sizeof(cw));
    memset(rr, 0, sizeof(rr));
    memset(cache, 0, sizeof(cache));
    prevprev_w50 = 0;
    rbytes = mbytes = 0;
    ok = true;
  }

  /* candidate v of the value predictor of a memory query */
  const uint32_t* value_candidate(uint32_t v, uint32_t type, uint32_t index, const uint32_t* row, const uint32_t* zero8) const {
    switch (v) {
      case 0: return zero8;
      case 1: return row + 8;
      case 2: return row + 16;
      case 3: return row + 24;
      default:
        if (type == 4u) return cache[index % ZKB_CW_SETS][v - 4u];
        return v == 4u ? row + 32 : zero8;
    }
  }

  /* one memory query.  row: the cycle's row (w2, w3 valid only when `post`), j: position inside the cycle */
  template <bool ENC>
  void mem_record(uint32_t* rec, const uint32_t* row, bool post, bool fe, uint32_t j) {
    static const uint32_t zero8[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const uint32_t pc_before = row[5] & 0xFFFFu;
    uint32_t head = 0, resid[12], pres = 0, fcode = 15, vsel = 0;
    if (ENC) {
      for (uint32_t c = 0; c < 15; c++)
        if (ZKB_MEM_FLAG_CODES[c] == rec[3]) {
          fcode = c;
          break;
        }
    } else {
      if ((uint64_t)(mend - mp) < 4) {
        ok = false;
        return;
      }
      memcpy(&head, mp, 4);
      mp += 4;
      pres = head & 0xFFFu;
      vsel = (head >> 12) & 7u;
      fcode = (head >> 16) & 15u;
      if ((uint64_t)(mend - mp) < (uint64_t)__builtin_popcount(pres) * 4) {
        ok = false;
        return;
      }
      memset(rec, 0, 48);
      mp = (uint8_t*)apply_residuals(pres, mp, rec);   /* rec[i] = residual of word i (0 where absent) */
    }
    /* w3 flags first: the other predictions look at the memory type and the rw bit */
    const uint32_t p3 = fcode < 15 ? ZKB_MEM_FLAG_CODES[fcode] : prev_mem[3];
    if (ENC) resid[3] = fcode < 15 ? 0u : rec[3] ^ p3;
    else rec[3] = fcode < 15 ? p3 : rec[3] ^ p3;
    const uint32_t flags = rec[3], type = flags & 0xFFu, rw = (flags >> 8) & 1u;
    const uint32_t p0 = row[1] + (rw ? 3u : 0u);
    const uint32_t p1 = type == 4u ? prev_row[50] : type <= 2u ? prev_row[51] + 1u + type : row[9];
    uint32_t p2;
    if (j > 0 && prev_mem[3] == flags) p2 = prev_mem[2] + 1u;
    else if (type == 4u) p2 = (j == 0 && fe) || !post ? pc_before >> 2 : row[3] & 0xFFFFu;
    else if (type == 0u) p2 = post ? (rw ? row[3] >> 16 : row[3] & 0xFFFFu) : 0u;
    else if (type <= 2u) p2 = row[8] >> 5;
    else p2 = (row[8] + row[10]) >> 5;
    if (ENC) {
      resid[0] = rec[0] ^ p0;
      resid[1] = rec[1] ^ p1;
      resid[2] = rec[2] ^ p2;
      uint32_t best = 9;
      for (uint32_t v = 0; v < 8; v++) {
        const uint32_t c = count_diff8(rec + 4, value_candidate(v, type, rec[2], row, zero8));
        if (c < best) {
          best = c;
          vsel = v;
        }
      }
      const uint32_t* cand = value_candidate(vsel, type, rec[2], row, zero8);
      for (int i = 0; i < 8; i++) resid[4 + i] = rec[4 + i] ^ cand[i];
      uint32_t words[12], k = 0;
      for (uint32_t i = 0; i < 12; i++)
        if (resid[i]) {
          pres |= 1u << i;
          words[k++] = resid[i];
        }
      head = pres | vsel << 12 | fcode << 16;
      if (mp) {
        memcpy(mp, &head, 4);
        memcpy(mp + 4, words, (size_t)k * 4);
        mp += 4 + (size_t)k * 4;
      }
      mbytes += 4 + (uint64_t)k * 4;
    } else {
      rec[0] ^= p0;
      rec[1] ^= p1;
      rec[2] ^= p2;
      const uint32_t* cand = value_candidate(vsel, type, rec[2], row, zero8);
      for (int i = 0; i < 8; i++) rec[4 + i] ^= cand[i];
    }
    /* state updates (both directions see the canonical record here) */
    if (type == 4u) {
      uint32_t(*set)[8] = cache[rec[2] % ZKB_CW_SETS];
      bool hit = false;
      for (uint32_t w = 0; w < ZKB_CW_WAYS; w++) hit = hit || count_diff8(rec + 4, set[w]) == 0;
      if (!hit) {
        uint32_t& r = rr[rec[2] % ZKB_CW_SETS];
        memcpy(set[r], rec + 4, 32);
        r = (r + 1u) % ZKB_CW_WAYS;
      }
      if (j == 0 && rw == 0 && rec[2] == pc_before >> 2) memcpy(cw, rec + 4, 32);
    }
    prev_mem = rec;
  }

  /* rows[n_rows][64], mem[n_mem][12]: canonical records (input when ENC, output otherwise) */
  template <bool ENC>
  void run(uint32_t* rows, uint64_t n_rows, uint32_t* mem, uint64_t n_mem) {
    uint64_t mi = 0;
    for (uint64_t r = 0; r < n_rows && ok; r++) {
      uint32_t* row = rows + r * 64;
      uint64_t mask = 0;
      uint32_t pred[64], dsel = 0, ncode = 7;
      const uint32_t pc = prev_row[5] >> 16;
      const uint32_t p0 = prev_row[0] + 1u, p1 = prev_row[1] + ZK_TIME_DELTA_PER_CYCLE, p5 = pc | ((pc + 1u) & 0xFFFFu) << 16;
      uint32_t cand[4][8];
      if (ENC) {
        /* predictions that look at nothing of this row */
        pred[0] = p0;
        pred[1] = p1;
        pred[2] = pred[3] = pred[4] = pred[7] = 0u;
        pred[5] = p5;
        pred[6] = prev_row[6];
        memset(pred + 8, 0, 32 * 4);
        memcpy(pred + 40, prev_row + 40, 15 * 4);
        pred[43] = 0u;
        memset(pred + 55, 0, 9 * 4);
      } else {
        if ((uint64_t)(rend - rp) < 8) {
          ok = false;
          return;
        }
        uint32_t m32[2];
        memcpy(m32, rp, 8);
        rp += 8;
        mask = (uint64_t)m32[0] | (uint64_t)m32[1] << 32;
        const uint64_t pres = mask & ((1ull << ZKB_ROW_TX_WORDS) - 1ull);
        dsel = (uint32_t)(mask >> 55) & 3u;
        ncode = (uint32_t)(mask >> 57) & 7u;
        if ((uint64_t)(rend - rp) < (uint64_t)__builtin_popcountll(pres) * 4) {
          ok = false;
          return;
        }
        /* the same predictions written straight into the row; words whose prediction looks at the row itself (2, 3, 4, 7,
         * dst0, 43, 48) start from zero and hold their bare residual until they are fixed up below */
        row[0] = p0;
        row[1] = p1;
        row[2] = row[3] = row[4] = row[7] = 0u;
        row[5] = p5;
        row[6] = prev_row[6];
        memset(row + 8, 0, 32 * 4);
        memcpy(row + 40, prev_row + 40, 15 * 4);
        row[43] = row[48] = 0u;
        memset(row + 55, 0, 9 * 4);
        rp = (uint8_t*)apply_residuals(pres, rp, row);
        if (ncode < 7) row[43] = ncode;
      }
      /* w48 looks at w5, dst0 at src0 / src1 */
      const uint32_t pc_before = row[5] & 0xFFFFu;
      const uint32_t p48 = (prev_row[48] & 0xFFFFu) | (pc_before >> 2) << 16;
      if (ENC) {
        pred[48] = p48;
        memset(cand[0], 0, 32);
        memcpy(cand[1], row + 8, 32);
        limbs_add(row + 8, row + 16, cand[2]);
        limbs_sub(row + 8, row + 16, cand[3]);
        uint32_t best = 9;
        for (uint32_t v = 0; v < 4; v++) {
          const uint32_t c = count_diff8(row + 24, cand[v]);
          if (c < best) {
            best = c;
            dsel = v;
          }
        }
        ncode = row[43] < 7u ? row[43] : 7u;
        memcpy(pred + 24, cand[dsel], 32);
      } else {
        row[48] ^= p48;
        if (dsel == 1) {
          for (int i = 0; i < 8; i++) row[24 + i] ^= row[8 + i];
        } else if (dsel >= 2) {   /* only the selected predictor is computed when decoding */
          if (dsel == 2) limbs_add(row + 8, row + 16, cand[0]);
          else limbs_sub(row + 8, row + 16, cand[0]);
          for (int i = 0; i < 8; i++) row[24 + i] ^= cand[0][i];
        }
      }
      /* the cycle's memory queries: the first one ahead of the opcode when an instruction fetch is expected */
      const bool fe = r == 0 || (pc_before >> 2) != (prev_row[48] >> 16) || prev_row[50] != prevprev_w50;
      const uint64_t nm = std::min<uint64_t>(row[43] & 0xFFFFu, n_mem - mi);
      uint32_t j = 0;
      if (fe && nm >= 1) {
        mem_record<ENC>(mem + mi * 12, row, false, fe, 0);
        mi++;
        j = 1;
        if (!ok) return;
      }
      const uint32_t sub = pc_before & 3u;
      pred[2] = cw[6 - 2 * sub];
      pred[3] = cw[7 - 2 * sub];
      if (!ENC) {
        row[2] ^= pred[2];
        row[3] ^= pred[3];
      }
      const uint32_t vidx = row[2] & ((1u << ZK_VARIANT_BITS) - 1u);
      pred[4] = vidx | 1u << 16;
      pred[7] = prev_row[7] - ZK_OPCODE_PRICES[vidx];
      if (!ENC) {
        row[4] ^= pred[4];
        row[7] ^= pred[7];
      } else {
        uint32_t words[64], k = 0;
        for (uint32_t i = 0; i < ZKB_ROW_TX_WORDS; i++) {
          if (i == 43 && ncode < 7) continue;
          const uint32_t x = row[i] ^ pred[i];
          if (x) {
            mask |= 1ull << i;
            words[k++] = x;
          }
        }
        mask |= (uint64_t)dsel << 55 | (uint64_t)ncode << 57;
        if (rp) {
          const uint32_t m32[2] = {(uint32_t)mask, (uint32_t)(mask >> 32)};
          memcpy(rp, m32, 8);
          memcpy(rp + 8, words, (size_t)k * 4);
          rp += 8 + (size_t)k * 4;
        }
        rbytes += 8 + (uint64_t)k * 4;
      }
      for (; j < nm && ok; j++, mi++) mem_record<ENC>(mem + mi * 12, row, true, fe, j);
      prevprev_w50 = prev_row[50];
      prev_row = row;
    }
    /* memory queries no row announces (a cycle that stopped the VM emits its queries but no row) */
    static const uint32_t zero_row[64] = {0};
    for (uint32_t j = 0; mi < n_mem && ok; mi++, j++) mem_record<ENC>(mem + mi * 12, zero_row, true, false, j);
  }
};

/* sizes (dst NULL) or payloads of one VM's rows + memory queries */
static inline void encode_joint(const void* rows, uint64_t n_rows, const void* mem, uint64_t n_mem, uint8_t* rows_dst, uint8_t* mem_dst,
                                uint64_t* rows_bytes, uint64_t* mem_bytes) {
  JointCoder* jc = new JointCoder;
  jc->reset();
  jc->rp = rows_dst;
  jc->mp = mem_dst;
  jc->rend = jc->mend = nullptr;
  jc->run<true>((uint32_t*)rows, n_rows, (uint32_t*)mem, n_mem);   /* (ENC only reads the records) */
  *rows_bytes = jc->rbytes;
  *mem_bytes = jc->mbytes;
  delete jc;
}

/* canonical rows_out[n_rows * 256 bytes] / mem_out[n_mem * 48 bytes] from the two payload slices; false = malformed */
static inline bool decode_joint(const uint8_t* rsrc, uint64_t rlen, uint64_t n_rows, const uint8_t* msrc, uint64_t mlen, uint64_t n_mem,
                                void* rows_out, void* mem_out) {
  static thread_local JointCoder coder;
  JointCoder* jc = &coder;
  jc->reset();
  jc->rp = (uint8_t*)rsrc;
  jc->rend = rsrc + rlen;
  jc->mp = (uint8_t*)msrc;
  jc->mend = msrc + mlen;
  jc->run<false>((uint32_t*)rows_out, n_rows, (uint32_t*)mem_out, n_mem);
  return jc->ok && jc->rp == jc->rend && jc->mp == jc->mend;
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
  bool present(uint32_t kind) const { return h->reserved0 == 0 || ((h->reserved0 >> kind) & 1u); }   /* a blob may carry a subset of the streams */
  /* the slice of stream `kind`'s payload that belongs to VM `vm`; false on a malformed offset table */
  bool slice(uint32_t vm, uint32_t kind, const uint8_t** src, uint64_t* len) const {
    const uint64_t lo = offsets(kind)[vm], hi = offsets(kind)[vm + 1];
    if (hi < lo || hi > h->payload_bytes[kind]) return false;
    *src = base + h->payload_offset[kind] + lo;
    *len = hi - lo;
    return true;
  }
  /* ROWS and MEM of one VM in one walk (they are coded jointly): canonical bytes into rows_dst / mem_dst, whose
   * capacities must cover counts(vm)[0] * 256 and counts(vm)[1] * 48 bytes; false on a malformed blob */
  bool decode_rows_mem(uint32_t vm, void* rows_dst, uint64_t rows_cap, void* mem_dst, uint64_t mem_cap) const {
    if (vm >= h->n_vms || !present(ZKB_STREAM_ROWS) || !present(ZKB_STREAM_MEM)) return false;
    const uint64_t nr = counts(vm)[ZKB_STREAM_ROWS], nm = counts(vm)[ZKB_STREAM_MEM];
    if (nr * ZKB_ROW_BYTES > rows_cap || nm * ZKB_MEM_BYTES > mem_cap) return false;
    const uint8_t *rs, *ms;
    uint64_t rl, ml;
    if (!slice(vm, ZKB_STREAM_ROWS, &rs, &rl) || !slice(vm, ZKB_STREAM_MEM, &ms, &ml)) return false;
    return decode_joint(rs, rl, nr, ms, ml, nm, rows_dst, mem_dst);
  }
  /* canonical bytes of VM `vm`'s stream `kind` -> dst (capacity max_bytes); returns the canonical length, or
   * UINT64_MAX on a malformed blob.  dst == NULL: length only.  (ROWS or MEM alone still walks both: prefer
   * decode_rows_mem when both are wanted.) */
  uint64_t decode(uint32_t vm, uint32_t kind, void* dst, uint64_t max_bytes, bool reference_decoder = false) const {
    if (vm >= h->n_vms || kind >= ZKB_N_STREAMS) return UINT64_MAX;
    const uint64_t n = present(kind) ? counts(vm)[kind] : 0, need = n * ZKB_CODEC_REC_WORDS[kind] * 4;
    if (!dst) return need;
    if (need > max_bytes) return UINT64_MAX;
    if (kind <= ZKB_STREAM_MEM) {
      if (n == 0 && !present(kind)) return 0;
      const uint32_t other = kind ^ 1u;
      static thread_local std::vector<uint8_t> tmp;   /* the other stream of the walk: scratch, reused from VM to VM */
      const size_t other_bytes = (size_t)counts(vm)[other] * ZKB_CODEC_REC_WORDS[other] * 4 + 1;
      if (tmp.size() < other_bytes) tmp.resize(other_bytes);
      const bool good = kind == ZKB_STREAM_ROWS ? decode_rows_mem(vm, dst, max_bytes, tmp.data(), other_bytes)
                                                : decode_rows_mem(vm, tmp.data(), other_bytes, dst, max_bytes);
      return good ? need : UINT64_MAX;
    }
    const uint8_t* src;
    uint64_t len;
    if (!slice(vm, kind, &src, &len)) return UINT64_MAX;
    const uint64_t used = reference_decoder ? decode_records_ref(kind, src, len, n, dst) : decode_records(kind, src, len, n, dst);
    return used == len ? need : UINT64_MAX;
  }
};

}  // namespace zkb_codec
#endif
#endif
