// The batched EraVM interpreter: ONE OCTET (8 lanes, a quarter warp) = ONE VM for the whole run, FOUR VMs per warp.
// A U256 has eight 32-bit limbs, so eight lanes are what one VM can keep busy; the four octets of a warp run four VMs of
// the batch side by side -- converged (one issue slot per instruction for all four) while they follow the same path
// through the interpreter, which batches of transactions against the same contracts do, and diverged (each octet
// syncs only with itself, u256.cuh) where they do not.
//
// Replaces (reference, /root/reference/src): vm_state/cycle.rs:19-429 (read_and_decode + cycle),
// vm_state/mem_ops.rs:14-125, vm_state/helpers.rs:10-338, every handler in opcodes/execution/*.rs and the
// backends reference_impls/{memory,decommitter}.rs + testing/storage.rs, re-designed for sm_100a:
//   * opcode dispatch is octet-uniform; the 8 lanes of a VM are used for limb-parallel U256 arithmetic (u256.cuh),
//     the column-per-lane keccak state (keccak.cuh) and vectorised record stores;
//   * architectural state lives in shared memory (15 x 256-bit registers, current frame, the "live" tail of
//     the cycle row) and octet-uniform registers (pc, sp, ergs, flags, timestamp ...);
//   * SimpleMemory's growable pages become bounded per-VM slabs in HBM (stack per far-call level, a pool of heap
//     slabs with a free mask, a 16-entry page-indirection table), InMemoryStorage becomes a per-VM open-addressed
//     table + one rollback journal whose frame marks give the reference's rollback semantics;
//   * every VmWitnessTracer callback becomes a packed record (include/zkb_records.h) appended to the VM's stream
//     slab: the 256-byte cycle row is two 16-byte stores per lane (two full 128-byte lines per VM and cycle).
#pragma once
#include <stdint.h>

#include "../../include/zkb.h"
#define ZK_TABLE_QUALIFIER __device__ __constant__
#include "isa_tables.inc"
#include "keccak.cuh"
#include "sha256.cuh"
#include "secp256k1.cuh"
#include "u256.cuh"

namespace zkb {

// ---------------------------------------------------------------------------------------------------
// per-VM hot state as stored in HBM between runs (loaded into shared memory / registers by run_vm)
// ---------------------------------------------------------------------------------------------------
// frame layout (32 words) == ZkbFrameRec words 2..28, then device-only fields
enum {
  F_THIS = 0, F_SENDER = 5, F_CODE_ADDR = 10, F_BASE_PAGE = 15, F_CODE_PAGE = 16, F_SP_PC = 17, F_EH_SHARDS = 18,
  F_ERGS = 19, F_MISC = 20 /* code_shard | is_static << 8 | is_local << 16 */, F_CTX = 21, F_HEAP_BOUND = 25,
  F_AUX_BOUND = 26, F_CODE_ID = 27, F_JOURNAL_MARK = 28, F_FAR_LEVEL = 29
};
// live tail of the cycle row (row words 40..55), kept current in shared memory
enum {
  L_DEPTH = 40, L_SPENT_PUBDATA = 41, L_PAGE_COUNTER = 42, L_COUNTS = 43, L_CTX = 44, L_TX_PSP = 48, L_EPP = 49,
  L_CODE_PAGE = 50, L_BASE_PAGE = 51, L_HEAP_BOUND = 52, L_AUX_BOUND = 53, L_EH_BITS = 54
};
enum {
  X_TIMESTAMP = 0, X_CYCLE = 1, X_FLAGS = 2, X_PENDING = 3, X_STATUS = 4, X_PTRMASK = 5, X_PREV_CODE_PAGE = 6,
  X_FAR_DEPTH = 7, X_JOURNAL_LEN = 8, X_N_DECOMMIT = 9, X_SLAB_FREE = 10, X_COUNT0 = 11 /* ..16 */, X_ABS_STEP = 17,
  X_DEFER = 18 /* 0, or the DEFER_* kind of a precompile result the FAST kernel left for the FULL one */
};
// DevBatch.defer[vm][ZKB_DEFER_WORDS]: what a parked VM's pending precompile needs (written by the FAST kernel, read by FULL)
enum { DEFER_NONE = 0, DEFER_ECRECOVER = 1, DEFER_KECCAK = 2, DEFER_CONTINUE = 3 /* FULL kernel: between a deferred keccak phase and its re-entry */ };
enum { DF_EC_INPUT = 0 /* 32 words */, DF_EC_DESC = 32 /* kbuf[60..63] */, DF_KC = 36 /* kc[8] */, DF_CYCLES_RUN = 44 };
#define ZKB_DEFER_WORDS 48u

struct VmHot {
  uint32_t regs[16][8];  // regs[0] is the constant-zero r0
  uint32_t F[32];
  uint32_t live[16];
  uint32_t prev_word[8];
  uint32_t x[32];
};

#ifndef ZKB_LOCKSTEP_PERIOD
#define ZKB_LOCKSTEP_PERIOD 32
#endif
// internal status (never leaves the kernel): the VM finished a cycle whose ecrecover is still to be computed; the
// schedulers park the VM state in HBM, run the recovery (DeferredEcrecover) and resume
#define ZKB_VM_YIELD_ECRECOVER 0x100u
// ... or whose keccak256 digest is still to be computed: the schedulers run the sponge batched over the VMs of the CTA,
// one THREAD per state (run_deferred_keccak), and the VM continues in place -- nothing of it has to be parked
#define ZKB_VM_YIELD_KECCAK 0x101u
// FAST kernel only: the VM waits (between two cycles) for the FULL kernel to finish its pending precompile
#define ZKB_VM_PARKED 0x102u
// kbuf layout while an ecrecover is pending
enum { KB_EC_INPUT = 0 /* 4 x 8 limbs */, KB_EC_PENDING = 60, KB_EC_MEM_INDEX = 61, KB_EC_SLAB = 62, KB_EC_OUT_WORD = 63 };  // 48..63: unused by sha256 / pop_frame
// ... and while a keccak256 is pending (the descriptor of the deferred sponge)
// (VmSmem.kc[]: its own words, the in-cycle keccak uses most of kbuf as scratch)
enum { KB_KC_PENDING = 0, KB_KC_IN_OFF = 1, KB_KC_IN_LEN = 2, KB_KC_SRC_SLAB = 3, KB_KC_OUT_SLAB = 4, KB_KC_OUT_WORD = 5, KB_KC_MEM_INDEX = 6, KB_KC_VM = 7 };
// block-placement hint: keeps the per-cycle hot path contiguous in the instruction stream (the interpreter's first-order
// cost is instruction fetch)
#define ZK_UNLIKELY(x) __builtin_expect(!!(x), 0)
#define ZKB_NO_SLAB 0xFFu
#define ZKB_NO_CODE 0xFFFFFFFFu
#define ZKB_PT_ENTRIES 32u   // four entries per octet lane
#define ZKB_DEC_ENTRIES 16u
#define ZKB_PT_FREE 0xFFFFFFFFu
enum { PT_HEAP_LIVE = 1, PT_AUX_LIVE = 2, PT_EXT = 3 };

struct DevBatch {
  uint32_t n_vms, witness;
  uint32_t chunk;              // lockstep schedule: VMs a CTA pulls per turn (<= VMs per CTA; balanced over the waves by zkb_run)
  const uint32_t* order;       // schedule slot -> VM index (VMs regrouped by bootloader code, zkb_run); nullptr = identity
  uint32_t warm_refund_bytes;  // ZkbConfig.reserved[1]: 0 = RefundType::None always (storage.rs:80-86), else the f-3 oracle
  uint32_t cap[ZKB_N_STREAMS];
  uint32_t stack_words, heap_words, n_slabs, max_far_depth, max_depth, storage_slots, journal_entries;
  const uint32_t* code_words;  // 8 u32 (LE limbs) per 256-bit code word, all bytecodes back to back
  const uint32_t* code_meta;   // per bytecode: offset_words, len_words, hash limbs[8]
  uint32_t n_codes;
  const uint32_t* code_index;  // open-addressed hash index over code_meta: slot -> bytecode id or ZKB_NO_CODE
  uint32_t code_index_mask;    // slots - 1 (power of two, >= 2 x n_codes)
  uint32_t default_aa[8];
  uint32_t zkporter;
  VmHot* hot;
  uint32_t* callstack;  // [vm][max_depth][32]
  uint32_t* stack_mem;  // [vm][max_far_depth + 1][stack_words][8]
  uint8_t* stack_ptr;   // [vm][max_far_depth + 1][stack_words]
  uint32_t* heap_mem;   // [vm][n_slabs][heap_words][8]
  uint32_t* lvl;        // [vm][max_far_depth + 1][4]: heap slab, aux slab, stack hwm, -
  uint32_t* slab_hwm;   // [vm][n_slabs]   words touched
  uint32_t* pt;         // [vm][16][2]     page, kind | slab << 8 | cleanup_level << 16
  uint32_t* dec;        // [vm][16][2]     code id, page
  uint32_t* st_tags;    // [vm][slots]
  uint32_t* st_keys;    // [vm][slots][8]
  uint32_t* st_addr;    // [vm][slots][8]  5 address words, shard, -, -
  uint32_t* st_vals;    // [vm][slots][8]
  uint32_t* j_slot;     // [vm][journal]
  uint32_t* j_val;      // [vm][journal][8]
  uint8_t* streams[ZKB_N_STREAMS];
  unsigned int* queue;    // [2] VM work queues of the FAST and the FULL interpreter launch of one zkb_run
  uint32_t* defer;        // [vm][ZKB_DEFER_WORDS]
  uint32_t* host_counts;  // [vm][8] in mapped pinned HOST memory: 6 stream counts, status, cycles (written once per run)
};

__device__ __forceinline__ uint32_t rec_bytes(int kind) {
  return kind == ZKB_STREAM_ROWS ? ZKB_ROW_BYTES : kind == ZKB_STREAM_MEM ? ZKB_MEM_BYTES : kind == ZKB_STREAM_LOG ? ZKB_LOG_BYTES
         : kind == ZKB_STREAM_DECOMMIT ? ZKB_DECOMMIT_BYTES : kind == ZKB_STREAM_FRAME ? ZKB_FRAME_BYTES : ZKB_REFUND_BYTES;
}

// shared memory per VM.  Size = 360 words = 8 (mod 32): the four VMs of a warp sit 8 banks apart, so one 4-byte access
// per lane by all four octets (same field, same limb index) is bank-conflict free.
struct __align__(16) VmSmem {
  uint32_t regs[16][8];
  uint32_t row[64];
  uint32_t F[32];
  uint32_t kbuf[64];   // precompile scratch: message staging / sha256 schedule / ecrecover stash; the keccak-f[1600] lane scratch
                       // (25 x u64, ks()) aliases its first 200 bytes -- staging is consumed before the permutation starts
  uint32_t pw[8];      // previous_code_word (cycle.rs:59-100): the 4 opcodes of the current code word
  uint32_t pt[ZKB_PT_ENTRIES * 2];  // page indirections (page, kind | slab << 8 | cleanup_level << 16); HBM copy: DevBatch.pt
  uint32_t hwm[32];    // words touched per heap slab; HBM copy: DevBatch.slab_hwm
  uint32_t lv[4];      // the current far level's entry of DevBatch.lvl: heap slab, aux slab, stack high-water mark, -
  // Octet-uniform cold state, kept here to leave the interpreter's register file to the hot state.
  // u[] / gp[]: written by every lane (same value) only inside load_frame_from_F / vm_load, i.e. between two octet syncs.
  // ul[k][l]: scalars that are read-modify-written in the middle of a cycle (stream counts, journal length, ...) --
  // every lane keeps a PRIVATE copy (lane l only ever touches ul[k][l]), so they need no sync and cannot race.
  uint32_t u[4];       // U_* below
  uint64_t gp[6];      // GP_* below: per-VM base pointers into the HBM slabs
  uint32_t ul[8][8];   // UL_* below
  uint32_t kc[8];      // KB_KC_*: descriptor of a pending (deferred) keccak256
  uint32_t pad[4];
  // keccak lane scratch (25 x u64).  The four VMs of a warp sit 8 banks apart and an 8-byte access is served per
  // half-warp (two VMs): with both at their natural offset their ten-bank windows overlap in two banks (1.9 G bank
  // conflicts on the keccak workload, ncu r01); the odd VM of each pair starts 8 words later, which clears them.
#ifdef ZKB_NO_KS_SKEW
  __device__ __forceinline__ uint64_t* ks() { return reinterpret_cast<uint64_t*>(kbuf); }
#else
  __device__ __forceinline__ uint64_t* ks() { return reinterpret_cast<uint64_t*>(kbuf + ((threadIdx.x >> 3) & 1u) * 8u); }
#endif
};
enum { U_FAR_DEPTH = 0, U_CODE_LEN = 1 };
enum { UL_JOURNAL_LEN = 0, UL_N_DECOMMIT, UL_SLAB_FREE, UL_COUNT2 /* LOG, DECOMMIT, FRAME, REFUND */ };
enum { GP_STACK = 0 /* stack page of the current far level */, GP_STACK_PTR, GP_HEAP, GP_LVL, GP_CODE };
static_assert(sizeof(VmSmem) == 1952 && (sizeof(VmSmem) / 4) % 32 == 8 && sizeof(VmSmem) % 16 == 0, "VmSmem bank skew");
typedef VmSmem WarpSmem;

// KD ("full"): this instantiation contains the out-of-line precompile routines -- the deferred keccak256 sponge for long
// inputs and the ecrecover recovery.  Two instantiations, because the mere PRESENCE of such a callee in the kernel costs
// the interpreter loop 6 % (ecrecover) to 15 % (keccak) through ptxas' register allocation around the call ABI
// (measured): every zkb_run launches the FAST kernel (KD = false: no call anywhere; a VM whose cycle left such a result
// pending is parked, between two cycles, with its descriptor in DevBatch.defer) and then the FULL kernel, which picks up
// exactly the parked VMs, finishes the pending result and runs them on.  Batches without such precompiles (the ERC-20
// workload hashes 64-byte preimages in the cycle) never leave the fast kernel.
template <bool KD>
struct Vm {
  const DevBatch& B;
  WarpSmem& S;
  const uint32_t vm;
  const uint32_t lane;  // lane within the VM's octet (0..7)
  // octet-uniform registers (the hot state; the cold state sits in S.u / S.gp)
  uint32_t pc, sp, ergs, flags, timestamp, cycle, pending, ptr_mask, status;
  uint32_t prev_code_page;
  uint32_t n_rows, n_mem;   // records in the ROWS / MEM streams (the two streams written every cycle)
  uint8_t* row_ptr;         // next free slot of this VM's ROWS / MEM slab
  uint8_t* mem_ptr;
  uint32_t rowbits;
  uint32_t ccount;  // per-cycle record counts, packed as CycleRow.n_mem | n_log << 16 | n_dfr << 24
  uint32_t forbid;  // ZK_E_* bits an opcode must not have in the current frame (kernel-only / not-in-static)
  uint32_t tx_psp;  // tx_number_in_block | previous_super_pc << 16 (row word L_TX_PSP, written back at row emission)
  // decoded opcode (octet-uniform): table entry + the operand fields of the (masked) instruction
  uint32_t entry, ops_lo, ops_hi;
  uint32_t dst_loc;  // stack destination of dst0: index | 1 << 16 when valid
  uint32_t defer_kind;  // DEFER_* of a parked VM (FAST kernel), written to x[X_DEFER] by vm_store

  __device__ Vm(const DevBatch& b, VmSmem& s, uint32_t vm_, uint32_t lane_) : B(b), S(s), vm(vm_), lane(lane_) {}
  // cold state accessors
  __device__ __forceinline__ uint32_t& far_depth() const { return S.u[U_FAR_DEPTH]; }
  __device__ __forceinline__ uint32_t& journal_len() const { return S.ul[UL_JOURNAL_LEN][lane]; }
  __device__ __forceinline__ uint32_t& n_decommit() const { return S.ul[UL_N_DECOMMIT][lane]; }
  __device__ __forceinline__ uint32_t& slab_free() const { return S.ul[UL_SLAB_FREE][lane]; }
  __device__ __forceinline__ uint32_t& count(int kind) const { return S.ul[UL_COUNT2 + kind - 2][lane]; }  // kind >= ZKB_STREAM_LOG
  __device__ __forceinline__ uint32_t stream_count(int kind) const { return kind == ZKB_STREAM_ROWS ? n_rows : kind == ZKB_STREAM_MEM ? n_mem : S.ul[UL_COUNT2 + kind - 2][lane]; }
  __device__ __forceinline__ uint32_t code_len() const { return S.u[U_CODE_LEN]; }
  __device__ __forceinline__ const uint32_t* code() const { return reinterpret_cast<const uint32_t*>(S.gp[GP_CODE]); }
  __device__ __forceinline__ uint32_t* g_stack() const { return reinterpret_cast<uint32_t*>(S.gp[GP_STACK]); }
  __device__ __forceinline__ uint8_t* g_stack_ptr() const { return reinterpret_cast<uint8_t*>(S.gp[GP_STACK_PTR]); }
  __device__ __forceinline__ uint32_t* g_heap() const { return reinterpret_cast<uint32_t*>(S.gp[GP_HEAP]); }
  __device__ __forceinline__ uint32_t* g_lvl() const { return reinterpret_cast<uint32_t*>(S.gp[GP_LVL]); }
  __device__ __forceinline__ uint32_t dst0_reg() const { return (ops_lo >> 24) & 15u; }
  __device__ __forceinline__ uint32_t dst1_reg() const { return ops_lo >> 28; }
  __device__ __forceinline__ uint32_t imm0() const { return ops_hi & 0xFFFFu; }
  __device__ __forceinline__ uint32_t imm1() const { return ops_hi >> 16; }
  // per-VM bases (vm_load); the stack bases are re-pointed at the current far level by load_frame_from_F
  __device__ __forceinline__ void set_vm_pointers() {
    S.gp[GP_HEAP] = reinterpret_cast<uint64_t>(B.heap_mem + (size_t)vm * B.n_slabs * B.heap_words * 8);
    S.gp[GP_LVL] = reinterpret_cast<uint64_t>(B.lvl + (size_t)vm * (B.max_far_depth + 1) * 4);
  }
  __device__ __forceinline__ void set_level_pointers(uint32_t level) {
    const size_t page = (size_t)vm * (B.max_far_depth + 1) + level;
    S.gp[GP_STACK] = reinterpret_cast<uint64_t>(B.stack_mem + page * B.stack_words * 8);
    S.gp[GP_STACK_PTR] = reinterpret_cast<uint64_t>(B.stack_ptr + page * B.stack_words);
  }

  // ---- small helpers -------------------------------------------------------------------------------
  __device__ __forceinline__ void fail(uint32_t code) {
    if (status == ZKB_VM_RUNNING) status = code;
  }
  __device__ __forceinline__ uint32_t L(int w) const { return S.row[w]; }
  __device__ __forceinline__ void setL(int w, uint32_t v) {
    if (lane == 0) S.row[w] = v;
    osync();
  }
  __device__ __forceinline__ bool is_kernel() const { return (S.row[L_EH_BITS] >> 16) & ZKB_FRAMEBIT_KERNEL; }
  __device__ __forceinline__ bool is_static() const { return (S.row[L_EH_BITS] >> 16) & ZKB_FRAMEBIT_STATIC; }
  __device__ __forceinline__ bool is_local() const { return (S.row[L_EH_BITS] >> 16) & ZKB_FRAMEBIT_LOCAL; }
  __device__ __forceinline__ u256l reg_read(uint32_t idx) const { return S.regs[idx][lane]; }
  __device__ __forceinline__ void reg_write(uint32_t idx, u256l v, bool is_ptr) {
    // limb l is written and read back by lane l only; the one cross-lane reader (limb 0 in operand addressing of a LATER
    // cycle) sits behind the octet sync of the row emission, so no sync is needed here
    if (idx != 0) {
      S.regs[idx][lane] = v;
      ptr_mask = (ptr_mask & ~(1u << idx)) | ((is_ptr ? 1u : 0u) << idx);
    }
  }
  // address (5 LE-loaded words holding 20 BE bytes, in lanes 0..4) <-> U256 (address_to_u256, utils.rs:29-41)
  __device__ __forceinline__ u256l addr_words_to_u256(uint32_t aw) const {
    uint32_t v = oshfl(aw, (4 - (int)lane) & 7);
    return lane < 5 ? bswap32(v) : 0u;
  }
  __device__ __forceinline__ uint32_t u256_to_addr_words(u256l v) const {
    uint32_t x = oshfl(v, (4 - (int)lane) & 7);
    return lane < 5 ? bswap32(x) : 0u;
  }

  // ---- record emission (VmWitnessTracer callbacks -> packed records) ---------------------------------
  // Every record is written straight from the registers that hold its fields: the scalar header by lane 0 as one
  // 8/16-byte vector store, each U256 by lanes 0..7 (one 32-byte segment) -- no shuffles, no lane-select chains.
  __device__ __forceinline__ uint8_t* stream_slot(int kind) {
    uint32_t n = count(kind);
    if (n >= B.cap[kind]) {
      fail(ZKB_VM_CAP_STREAM);
      return nullptr;
    }
    count(kind) = n + 1;
    if (!B.witness) return nullptr;
    return B.streams[kind] + ((size_t)vm * B.cap[kind] + n) * rec_bytes(kind);
  }

  // witness_tracer.add_memory_query (helpers.rs:36,70,111) / precompile memory witness (helpers.rs:215-221)
  __device__ __forceinline__ void emit_mem(uint32_t ts, uint32_t page, uint32_t index, uint32_t mtype, uint32_t rw, uint32_t is_ptr,
                                           uint32_t origin, u256l value) {
    ccount += 1u;
    if (ZK_UNLIKELY(n_mem >= B.cap[ZKB_STREAM_MEM])) {
      fail(ZKB_VM_CAP_STREAM);
      return;
    }
    n_mem++;
    uint32_t* p = reinterpret_cast<uint32_t*>(mem_ptr);
    mem_ptr += ZKB_MEM_BYTES;
    if (!B.witness) return;
    if (lane == 0) *reinterpret_cast<uint4*>(p) = make_uint4(ts, page, index, mtype | rw << 8 | is_ptr << 16 | origin << 24);
    p[4 + lane] = value;
  }

  // witness_tracer.add_log_query (helpers.rs:151,161,208); aw = address words in lanes 0..4
  __device__ __forceinline__ void emit_log(uint32_t ts, uint32_t aux, uint32_t shard, uint32_t aw, uint32_t rw, uint32_t is_service,
                                           u256l key, u256l read_value, u256l written_value) {
    ccount += 1u << 16;
    uint32_t* p = (uint32_t*)stream_slot(ZKB_STREAM_LOG);
    uint32_t tx = tx_psp & 0xFFFFu;
    if (p) {
      if (lane == 0) {
        *reinterpret_cast<uint2*>(p) = make_uint2(ts, tx | aux << 16 | shard << 24);
        p[7] = rw | is_service << 16;
      }
      if (lane < 5) p[2 + lane] = aw;
      p[8 + lane] = key;
      p[16 + lane] = read_value;
      p[24 + lane] = written_value;
    }
  }

  // witness_tracer.add_decommittment (helpers.rs:185-191)
  __device__ __forceinline__ void emit_decommit(uint32_t ts, uint32_t page, uint32_t len, uint32_t fresh, u256l hash) {
    ccount += 1u << 24;
    uint32_t* p = (uint32_t*)stream_slot(ZKB_STREAM_DECOMMIT);
    if (p) {
      if (lane == 0) *reinterpret_cast<uint4*>(p) = make_uint4(ts, page, (len & 0xFFFFu) | fresh << 16, 0u);
      p[4 + lane] = hash;
    }
  }

  // witness_tracer.record_refund_for_query (helpers.rs:130-134): type 0 = RefundType::None, 1 = RepeatedWrite(value)
  __device__ __forceinline__ void emit_refund(uint32_t type, uint32_t value) {
    ccount += 1u << 28;
    uint32_t* p = (uint32_t*)stream_slot(ZKB_STREAM_REFUND);
    if (p && lane < 2) p[lane] = lane == 0 ? type : value;
  }

  // start_new_execution_context (helpers.rs:237-241): the new frame must already be in S.F
  __device__ __forceinline__ void emit_frame_start(uint32_t prev_ergs, uint32_t prev_pc, uint32_t prev_sp, uint32_t bound_kind,
                                                   uint32_t bound_value) {
    ccount += 1u << 26;
    uint32_t* p = (uint32_t*)stream_slot(ZKB_STREAM_FRAME);
    if (p) {
#pragma unroll
      for (uint32_t q = 0; q < 4; q++) {
        const uint32_t i = lane + 8u * q;
        uint32_t w = i == 0 ? (ZKB_FRAMEKIND_START | bound_kind << 16) : i == 1 ? cycle : i < 29 ? S.F[(i - 2) & 31]
                     : i == 29 ? prev_ergs : i == 30 ? (prev_pc | prev_sp << 16) : bound_value;
        p[i] = w;
      }
    }
  }
  // finish_execution_context (helpers.rs:258-259)
  __device__ __forceinline__ void emit_frame_finish(bool panicked) {
    ccount += 1u << 26;
    uint32_t* p = (uint32_t*)stream_slot(ZKB_STREAM_FRAME);
    if (p) {
      const uint32_t head = lane == 0 ? (ZKB_FRAMEKIND_FINISH | (panicked ? 1u : 0u) << 8) : lane == 1 ? cycle : 0u;
      p[lane] = head;
      p[8 + lane] = 0u;
      p[16 + lane] = 0u;
      p[24 + lane] = 0u;
    }
  }

  // ---- stack page of the current far level (SimpleMemory stack_pages, memory.rs:412-437) ----------
  __device__ __forceinline__ u256l stack_read(uint32_t index, uint32_t& is_ptr) {
    is_ptr = 0;
    if (index >= B.stack_words) return 0u;  // never written (writes beyond the cap stop the VM) => still zero
    const uint32_t off = index;  // g_stack() / g_stack_ptr() point at the current far level's page
    is_ptr = g_stack_ptr()[off];
    return g_stack()[off * 8 + lane];
  }
  __device__ __forceinline__ void stack_write(uint32_t index, u256l v, uint32_t is_ptr) {
    if (index >= B.stack_words) {
      fail(ZKB_VM_CAP_STACK);
      return;
    }
    const uint32_t off = index;  // g_stack() / g_stack_ptr() point at the current far level's page
    g_stack()[off * 8 + lane] = v;
    if (lane == 1) g_stack_ptr()[off] = (uint8_t)is_ptr;
    if (lane == 0 && index + 1 > S.lv[2]) S.lv[2] = index + 1;
    // no octet sync: limb l is read back by lane l; the pointer byte and the high-water mark are read by other lanes only
    // in later cycles, i.e. behind the sync of the row emission
  }

  // ---- heap slabs (SimpleMemory heaps / pages_with_extended_lifetime, memory.rs:439-521) ----------
  __device__ __forceinline__ uint32_t slab_alloc() {
    if (slab_free() == 0) {
      fail(ZKB_VM_CAP_HEAP);
      return ZKB_NO_SLAB;
    }
    uint32_t s = __ffs(slab_free()) - 1;
    slab_free() &= ~(1u << s);
    return s;
  }
  __device__ __forceinline__ void slab_release(uint32_t s) {
    if (s == ZKB_NO_SLAB) return;
    uint32_t hwm = S.hwm[s];
    uint32_t* base = g_heap() + (size_t)s * B.heap_words * 8;
    uint4* base4 = reinterpret_cast<uint4*>(base);
    for (uint32_t i = lane; i < hwm * 2; i += 8) base4[i] = make_uint4(0u, 0u, 0u, 0u);  // == heap_on_return fill (memory.rs:181-183)
    osync();  // every lane has read the mark (loop bound above) before lane 0 resets it
    if (lane == 0) S.hwm[s] = 0;
    slab_free() |= 1u << s;
    osync();
  }
  __device__ __forceinline__ u256l slab_read(uint32_t s, uint32_t word) {
    if (s == ZKB_NO_SLAB || word >= B.heap_words) return 0u;
    return g_heap()[((size_t)s * B.heap_words + word) * 8 + lane];
  }
  __device__ __forceinline__ void slab_write(uint32_t s, uint32_t word, u256l v) {
    g_heap()[((size_t)s * B.heap_words + word) * 8 + lane] = v;
    if (lane == 0 && word + 1 > S.hwm[s]) S.hwm[s] = word + 1;  // lane 0 owns the mark (as in stack_write: no octet sync needed)
  }
  // heap (which = 0) / aux heap (which = 1) slab of far level x: the current level's entry lives in shared memory, the
  // callers' entries are current in HBM (written back when the callee's level started)
  __device__ __forceinline__ uint32_t level_slab(uint32_t x, uint32_t which) const { return x == far_depth() ? S.lv[which] : g_lvl()[x * 4 + which]; }
  // slab of the current frame's heap (which = 0) / aux heap (which = 1); allocate lazily on first write
  __device__ __forceinline__ uint32_t cur_slab(uint32_t which, bool for_write, uint32_t word) {
    uint32_t s = S.lv[which];
    if (for_write) {
      if (word >= B.heap_words) {
        fail(ZKB_VM_CAP_HEAP);
        return ZKB_NO_SLAB;
      }
      if (s == ZKB_NO_SLAB) {
        s = slab_alloc();
        osync();  // every lane has read the old entry
        if (lane == 0) S.lv[which] = s;
        osync();
      }
    }
    return s;
  }

  // ---- page indirections (SimpleMemory.page_numbers_indirections, memory.rs:160-171,475-521) ------
  __device__ __forceinline__ int pt_find(uint32_t page) {
    // octet lane l looks at entries 4l .. 4l+3 (two 16-byte loads)
    const uint4 a = reinterpret_cast<const uint4*>(S.pt)[lane * 2], b = reinterpret_cast<const uint4*>(S.pt)[lane * 2 + 1];
    const uint32_t h = (a.x == page ? 1u : 0u) | (a.z == page ? 2u : 0u) | (b.x == page ? 4u : 0u) | (b.z == page ? 8u : 0u);
    const uint32_t m = oballot(h != 0);
    if (!m) return -1;
    const int l = __ffs(m) - 1;
    return l * 4 + (__ffs(oshfl(h, l)) - 1);
  }
  __device__ __forceinline__ void pt_upsert(uint32_t page, uint32_t kind, uint32_t slab_or_level, uint32_t cleanup_level) {
    int e = pt_find(page);
    if (e < 0) e = pt_find(ZKB_PT_FREE);
    if (e < 0) {
      fail(ZKB_VM_CAP_PAGES);
      return;
    }
    if (lane == 0) {
      S.pt[e * 2] = page;
      S.pt[e * 2 + 1] = kind | slab_or_level << 8 | cleanup_level << 16;
    }
    osync();
  }
  // fat-pointer read of one word (memory.rs:475-521); ok = false => reference panic (unreachable page)
  __device__ __forceinline__ u256l fatptr_read(uint32_t page, uint32_t word, bool& ok) {
    ok = true;
    if (page == 0) return 0u;  // Indirection::Empty (memory.rs:226)
    int e = pt_find(page);
    if (e < 0) {
      ok = false;
      return 0u;
    }
    uint32_t info = S.pt[e * 2 + 1];
    uint32_t kind = info & 0xFFu, x = (info >> 8) & 0xFFu;
    uint32_t s = kind == PT_EXT ? x : level_slab(x, kind == PT_AUX_LIVE ? 1 : 0);
    return slab_read(s, word);
  }

  // ---- storage (InMemoryStorage, storage.rs:88-186) ----------------------------------------------
  // aw: address words in lanes 0..4.  Returns the previous value; on writes stores `nv` and journals.
  // mode (always a literal at the call site):
  //   ST_READ / ST_WRITE  execute_partial_query (storage.rs:88-139); with the refund-aware oracle enabled they also set
  //                       the slot's cold/warm marker (storage.rs:105-110,126-131; word 6 of the slot's address row --
  //                       a read of an absent key then claims a slot with value 0 to carry its marker)
  //   ST_POPULATE         InMemoryStorage::populate (storage.rs:26-32): no marker, no journal
  //   ST_IS_WARM          the probe behind estimate_refunds_for_write: returns the marker (0 / 1), changes nothing
  enum { ST_READ = 0, ST_WRITE = 1, ST_POPULATE = 2, ST_IS_WARM = 3 };
  __device__ __forceinline__ u256l storage_access(uint32_t shard, uint32_t aw, u256l key, const uint32_t mode, u256l nv) {
    const bool is_write = mode == ST_WRITE || mode == ST_POPULATE, journal = mode == ST_WRITE;
    const bool track_warm = B.warm_refund_bytes != 0u;
    uint32_t* tags = B.st_tags + (size_t)vm * B.storage_slots;
    uint32_t* keys = B.st_keys + (size_t)vm * B.storage_slots * 8;
    uint32_t* addrs = B.st_addr + (size_t)vm * B.storage_slots * 8;
    uint32_t* vals = B.st_vals + (size_t)vm * B.storage_slots * 8;
    // identity of a slot: key limb l in lane l, address word l in lanes 0..4, shard in lane 5
    const uint32_t extra = lane < 5 ? aw : lane == 5 ? shard : 0u;
    uint32_t h = (key + 0x9E3779B9u * (lane + 1u)) * 0x85EBCA6Bu;
    h ^= h >> 15;
    h = (h ^ extra) * 0xC2B2AE35u;
    h ^= h >> 13;
#pragma unroll
    for (int o = 4; o >= 1; o >>= 1) h += oshfl_xor(h, o) * 0x27D4EB2Fu + (h >> 13);
    h = oshfl(h, 0);
    uint32_t tag = h | 0x80000000u;
    uint32_t mask = B.storage_slots - 1;
    int found = -1, insert_at = -1;
    u256l old = 0u;
    uint32_t marker_w = 0u;
    for (uint32_t base = 0; base < B.storage_slots && found < 0 && insert_at < 0; base += ZK_OCT) {
      uint32_t idx = (h + base + lane) & mask;
      uint32_t t = tags[idx];
      uint32_t m_match = oballot(t == tag);
      uint32_t m_empty = oballot(t == 0u);
      uint32_t before_empty = m_empty ? ((1u << (__ffs(m_empty) - 1)) - 1u) : 0xFFFFFFFFu;
      m_match &= before_empty;
      while (m_match) {
        int c = __ffs(m_match) - 1;
        m_match &= m_match - 1;
        uint32_t ci = (h + base + c) & mask;
        // key, address and value of the candidate are fetched together (one HBM round trip instead of two)
        const uint32_t k = keys[ci * 8 + lane], a = lane < 7 ? addrs[ci * 8 + lane] : 0u, val = vals[ci * 8 + lane];
        if (oall(k == key && (lane >= 6 || a == extra))) {
          found = (int)ci;
          old = val;
          marker_w = a;  // lane 6: the slot's cold/warm marker
          break;
        }
      }
      if (found < 0 && m_empty) insert_at = (int)((h + base + __ffs(m_empty) - 1) & mask);
    }
    if (mode == ST_IS_WARM) return found >= 0 ? oshfl(marker_w, 6) : 0u;
    const uint32_t new_marker = (track_warm && mode != ST_POPULATE) ? 1u : 0u;
    if (is_write || (track_warm && found < 0)) {  // claim a slot: a new key, or (refund-aware oracle) the marker of an absent key
      int slot = found;
      if (slot < 0) {
        if (insert_at < 0) {
          fail(ZKB_VM_CAP_STORAGE);
          return 0u;
        }
        slot = insert_at;
        keys[slot * 8 + lane] = key;
        if (lane < 6) addrs[slot * 8 + lane] = extra;
        if (lane == 6) addrs[slot * 8 + 6] = new_marker;
        if (lane == 7) tags[slot] = tag;
        if (!is_write) vals[slot * 8 + lane] = 0u;
      } else if (new_marker && lane == 6) {
        addrs[slot * 8 + 6] = 1u;
      }
    } else if (new_marker && lane == 6) {
      addrs[found * 8 + 6] = 1u;  // read of a present key
    }
    if (is_write) {
      const int slot = found >= 0 ? found : insert_at;
      vals[slot * 8 + lane] = nv;
      if (journal) {
        if (journal_len() >= B.journal_entries) {
          fail(ZKB_VM_CAP_STORAGE);
        } else {
          uint32_t* js = B.j_slot + (size_t)vm * B.journal_entries;
          uint32_t* jv = B.j_val + (size_t)vm * B.journal_entries * 8;
          jv[journal_len() * 8 + lane] = old;
          if (lane == 0) js[journal_len()] = (uint32_t)slot;
          journal_len()++;
        }
      }
      osync();
    }
    return old;
  }
  // storage.finish_frame(panicked = true): undo the frame's writes in reverse (storage.rs:156-176)
  __device__ __forceinline__ void storage_rollback(uint32_t mark) {
    uint32_t* vals = B.st_vals + (size_t)vm * B.storage_slots * 8;
    const uint32_t* js = B.j_slot + (size_t)vm * B.journal_entries;
    const uint32_t* jv = B.j_val + (size_t)vm * B.journal_entries * 8;
    while (journal_len() > mark) {
      journal_len()--;
      uint32_t slot = js[journal_len()];
      vals[slot * 8 + lane] = jv[journal_len() * 8 + lane];
      osync();
    }
  }

  // ---- frames --------------------------------------------------------------------------------------
  // sync the register/row-resident fields of the current frame into S.F
  __device__ __forceinline__ void sync_frame_to_F() {
    if (lane == 0) {
      S.F[F_SP_PC] = sp | pc << 16;
      S.F[F_ERGS] = ergs;
      S.F[F_BASE_PAGE] = S.row[L_BASE_PAGE];
      S.F[F_CODE_PAGE] = S.row[L_CODE_PAGE];
      S.F[F_HEAP_BOUND] = S.row[L_HEAP_BOUND];
      S.F[F_AUX_BOUND] = S.row[L_AUX_BOUND];
      S.F[F_EH_SHARDS] = (S.F[F_EH_SHARDS] & 0xFFFF0000u) | (S.row[L_EH_BITS] & 0xFFFFu);
    }
    osync();
  }
  // derive the register/row-resident fields from S.F (after a push of a new frame or a pop)
  __device__ __forceinline__ void load_frame_from_F() {
    osync();
    uint32_t sp_pc = S.F[F_SP_PC];
    sp = sp_pc & 0xFFFFu;
    pc = sp_pc >> 16;
    ergs = S.F[F_ERGS];
    uint32_t misc = S.F[F_MISC];
    bool kernel = S.F[0] == 0 && S.F[1] == 0 && S.F[2] == 0 && S.F[3] == 0 && (S.F[4] & 0xFFFFu) == 0;  // execution_stack.rs:83-87
    uint32_t bits = (((misc >> 8) & 1u) ? ZKB_FRAMEBIT_STATIC : 0u) | (((misc >> 16) & 1u) ? ZKB_FRAMEBIT_LOCAL : 0u) | (kernel ? ZKB_FRAMEBIT_KERNEL : 0u);
    forbid = (kernel ? 0u : (uint32_t)ZK_E_KERNEL_ONLY) | (((misc >> 8) & 1u) ? (uint32_t)ZK_E_STATIC_FORBIDDEN : 0u);
    if (lane == 0) {
      S.row[L_BASE_PAGE] = S.F[F_BASE_PAGE];
      S.row[L_CODE_PAGE] = S.F[F_CODE_PAGE];
      S.row[L_HEAP_BOUND] = S.F[F_HEAP_BOUND];
      S.row[L_AUX_BOUND] = S.F[F_AUX_BOUND];
      S.row[L_EH_BITS] = (S.F[F_EH_SHARDS] & 0xFFFFu) | bits << 16;
    }
    uint32_t id = S.F[F_CODE_ID];
    if (id == ZKB_NO_CODE) {
      S.gp[GP_CODE] = 0;
      S.u[U_CODE_LEN] = 0;
    } else {
      S.gp[GP_CODE] = reinterpret_cast<uint64_t>(B.code_words + (size_t)B.code_meta[id * 10] * 8);
      S.u[U_CODE_LEN] = B.code_meta[id * 10 + 1];
    }
    const uint32_t level = S.F[F_FAR_LEVEL];
    far_depth() = level;
    set_level_pointers(level);
    osync();
  }
  // vm_state.start_frame (helpers.rs:225-246) in two halves.  push_begin saves the caller's frame (it stays readable
  // in S.F); the handler then turns S.F into the callee's frame and calls push_end.
  __device__ __forceinline__ bool push_begin() {
    const uint32_t depth = S.row[L_DEPTH];
    if (depth >= B.max_depth) {
      fail(ZKB_VM_CAP_DEPTH);
      return false;
    }
    sync_frame_to_F();
    reinterpret_cast<uint4*>(B.callstack + ((size_t)vm * B.max_depth + depth) * 32)[lane] = reinterpret_cast<const uint4*>(S.F)[lane];
    osync();
    return true;
  }
  // prev_*: the caller's ergs / pc / sp as the tracer sees the previous frame (after the call's own updates)
  __device__ __forceinline__ void push_end(uint32_t prev_ergs, uint32_t prev_pc, uint32_t prev_sp, uint32_t bound_kind = 0,
                                           uint32_t bound_value = 0) {
    osync();
    if (lane == 0) {
      S.F[F_JOURNAL_MARK] = journal_len();  // storage.start_frame / event_sink.start_frame
      S.row[L_DEPTH] = S.row[L_DEPTH] + 1;
    }
    osync();
    emit_frame_start(prev_ergs, prev_pc, prev_sp, bound_kind, bound_value);
    load_frame_from_F();
  }
  // vm_state.finish_frame (helpers.rs:248-264); leaves the parent in S.F
  __device__ __forceinline__ void pop_frame(bool panicked) {
    sync_frame_to_F();
    if (panicked) storage_rollback(S.F[F_JOURNAL_MARK]);
    emit_frame_finish(panicked);
    uint32_t depth = S.row[L_DEPTH];
    osync();
    reinterpret_cast<uint4*>(S.F)[lane] = reinterpret_cast<const uint4*>(B.callstack + ((size_t)vm * B.max_depth + depth - 1) * 32)[lane];
    if (lane == 0) S.row[L_DEPTH] = depth - 1;
    osync();
    load_frame_from_F();
  }

  // ---- operand write-back (helpers.rs:266-287) ------------------------------------------------------
  __device__ __forceinline__ void dst0_update(u256l v, bool is_ptr) {
    S.row[24 + lane] = v;
    rowbits |= ZKB_ROWBIT_DST0_VALID | (is_ptr ? ZKB_ROWBIT_DST0_PTR : 0u);
    if (dst_loc) {
      const uint32_t dst_loc_index = dst_loc & 0xFFFFu;
      stack_write(dst_loc_index, v, is_ptr ? 1u : 0u);
      emit_mem(timestamp + 3, L(L_BASE_PAGE) + 1, dst_loc_index, ZK_MEM_STACK, 1, is_ptr ? 1u : 0u, ZKB_MEMORIGIN_VM, v);
    } else {
      reg_write(dst0_reg(), v, is_ptr);
    }
  }
  __device__ __forceinline__ void dst1_update(u256l v, bool is_ptr) {
    S.row[32 + lane] = v;
    rowbits |= ZKB_ROWBIT_DST1_VALID | (is_ptr ? ZKB_ROWBIT_DST1_PTR : 0u);
    reg_write(dst1_reg(), v, is_ptr);
  }

  // handlers
  __device__ void op_context(uint32_t sub, u256l src0);
  __device__ void op_shift(uint32_t sub, u256l src0, u256l src1);
  __device__ void op_ptr(uint32_t sub, u256l src0, u256l src1, bool src0_ptr, bool src1_ptr);
  __device__ void op_near_call(u256l src0, uint32_t new_pc);
  __device__ void op_log(uint32_t sub, u256l src0, u256l src1);
  __device__ void op_far_call(uint32_t sub, u256l src0, u256l src1, bool src0_ptr, uint32_t new_pc, bool kernel_mode);
  __device__ void op_ret(uint32_t sub, u256l src0, bool src0_ptr);
  __device__ void op_uma(uint32_t sub, u256l src0, u256l src1, bool src0_ptr);
  __device__ void keccak_precompile(u256l abi);
  __device__ void keccak_precompile_inline(u256l abi);
  __device__ void sha256_precompile(u256l abi);
  __device__ void ecrecover_precompile(u256l abi);
  __device__ void memory_start_global_frame(uint32_t caller_level, uint32_t caller_base, uint32_t calldata_page);
  __device__ void memory_finish_global_frame(uint32_t level, uint32_t base_page, uint32_t returndata_page);
  __device__ void cycle_once();
};

// ===================================================================================================
// one VM cycle (cycle.rs:19-429)
// ===================================================================================================
template <bool KD>
__device__ __forceinline__ void Vm<KD>::cycle_once() {
  const uint32_t row_cycle = cycle, row_ts = timestamp, pc_before = pc;
  S.row[24 + lane] = 0u;  // dst0 / dst1 fields default to zero (lane l owns limb l of both, here and in dst*_update)
  S.row[32 + lane] = 0u;
  rowbits = 0;
  ccount = 0;
  dst_loc = 0;

  // ---- fetch (cycle.rs:46-130) ----
  const uint32_t code_page = S.row[L_CODE_PAGE];
  const uint32_t super_pc = pc >> 2, sub_pc = pc & 3u;
  uint32_t prev_super_pc = tx_psp >> 16;
  uint32_t raw_lo, raw_hi;
  if (!ZK_UNLIKELY(pending)) {
    if (code_page != prev_code_page || prev_super_pc != super_pc) {
      u256l w = super_pc < code_len() ? __ldg(code() + (size_t)super_pc * 8 + lane) : 0u;
      S.pw[lane] = w;
      osync();
      prev_super_pc = super_pc;
      emit_mem(timestamp, code_page, super_pc, ZK_MEM_CODE, 0, 0, ZKB_MEMORIGIN_VM, w);
    }
    const uint2 raw = *reinterpret_cast<const uint2*>(&S.pw[6 - 2 * (int)sub_pc]);  // broadcast read
    raw_lo = raw.x;
    raw_hi = raw.y;
  } else {
    pending = 0;
    prev_super_pc = super_pc;
    raw_lo = (uint32_t)ZK_EXCEPTION_REVERT_ENCODING;
    raw_hi = 0;
  }
  prev_code_page = code_page;

  // ---- decode, price, exceptions, condition (cycle.rs:132-217) ----
  uint32_t vidx = raw_lo & ((1u << ZK_VARIANT_BITS) - 1u);
  entry = ZK_OPCODE_TABLE[vidx];
  const uint32_t price = ZK_OPCODE_PRICES[vidx];
  uint32_t err = (entry & ZK_E_INVALID) ? 1u : 0u;
  if (ZK_UNLIKELY(ergs < price)) {
    ergs = 0;
    err |= 2u;
  } else {
    ergs -= price;
  }
  const bool kernel_mode = !(forbid & ZK_E_KERNEL_ONLY);
  if (ZK_UNLIKELY(entry & forbid)) err |= ((entry & forbid & ZK_E_KERNEL_ONLY) ? 4u : 0u) | ((entry & forbid & ZK_E_STATIC_FORBIDDEN) ? 8u : 0u);
  if (ZK_UNLIKELY(S.row[L_DEPTH] == ZK_VM_MAX_STACK_DEPTH)) err |= 16u;
  // flags: bit0 LT/OF, bit1 EQ, bit2 GT.  Byte `cond` of the table = the set of flag values that satisfy the condition
  // {Always, Gt, Lt, Eq, Ge, Le, Ne, GtOrLt} (cycle.rs:193-210)
  const uint32_t cond = (raw_lo >> ZK_COND_SHIFT) & 7u;
  const uint64_t kCondTable = 0xFA33EEFCCCAAF0FFull;
  bool resolved = (uint32_t)(kCondTable >> (cond * 8u + (flags & 7u))) & 1u;
  // mask_into_panic / mask_into_nop (cycle.rs:187-217): the masked opcode has all-zero operands
  ops_lo = raw_lo;
  ops_hi = raw_hi;
  if (ZK_UNLIKELY(err)) {
    vidx = ZK_PANIC_VARIANT_IDX;
    resolved = true;
  } else if (!resolved) {
    vidx = ZK_NOP_VARIANT_IDX;
  }
  if (err || !resolved) {
    entry = ZK_OPCODE_TABLE[vidx];
    ops_lo = ops_hi = 0u;
  }
  const uint32_t src0_reg = (ops_lo >> 16) & 15u, src1_reg = (ops_lo >> 20) & 15u;
  // delayed changes (mod.rs:134-153): previous_super_pc goes back to the row tail with the row head
  tx_psp = (tx_psp & 0xFFFFu) | prev_super_pc << 16;

  // ---- operand addressing (mem_ops.rs:14-125, cycle.rs:275-345) ----
  const uint32_t family = entry & 15u, sub = (entry >> ZK_E_SUB_SHIFT) & 15u;
  const uint32_t src_mode = (entry >> ZK_E_SRC_SHIFT) & 7u, dst_mode = (entry >> ZK_E_DST_SHIFT) & 3u;
  u256l src0 = reg_read(src0_reg);
  uint32_t src0_ptr = (ptr_mask >> src0_reg) & 1u;
  if (src_mode != ZK_SRC_REG) {
    uint32_t vaddr = (S.regs[src0_reg][0] + imm0()) & 0xFFFFu;
    if (src_mode == ZK_SRC_IMM) {
      src0 = lane == 0 ? imm0() : 0u;
      src0_ptr = 0;
    } else {
      uint32_t index;
      if (src_mode == ZK_SRC_STACK_POP) {
        sp = (sp - vaddr) & 0xFFFFu;
        index = sp;
      } else if (src_mode == ZK_SRC_STACK_REL) {
        index = (sp - vaddr) & 0xFFFFu;
      } else {
        index = vaddr;  // absolute stack or code page
      }
      if (family == ZK_OP_NOP) {
        src0 = 0u;  // NOP moves SP but never reads (cycle.rs:298-301)
        src0_ptr = 0;
      } else if (src_mode == ZK_SRC_CODE) {
        src0 = index < code_len() ? __ldg(code() + (size_t)index * 8 + lane) : 0u;
        src0_ptr = 0;
        emit_mem(timestamp, code_page, index, ZK_MEM_CODE, 0, 0, ZKB_MEMORIGIN_VM, src0);
      } else {
        src0 = stack_read(index, src0_ptr);
        emit_mem(timestamp, L(L_BASE_PAGE) + 1, index, ZK_MEM_STACK, 0, src0_ptr, ZKB_MEMORIGIN_VM, src0);
      }
    }
  }
  if (dst_mode != ZK_DST_REG) {
    uint32_t vaddr = (S.regs[dst0_reg()][0] + imm1()) & 0xFFFFu;
    uint32_t dst_loc_index;
    if (dst_mode == ZK_DST_STACK_PUSH) {
      dst_loc_index = sp;
      sp = (sp + vaddr) & 0xFFFFu;
    } else if (dst_mode == ZK_DST_STACK_REL) {
      dst_loc_index = (sp - vaddr) & 0xFFFFu;
    } else {
      dst_loc_index = vaddr;
    }
    dst_loc = dst_loc_index | 1u << 16;
  }
  u256l src1 = reg_read(src1_reg);
  uint32_t src1_ptr = (ptr_mask >> src1_reg) & 1u;
  if (entry & ZK_E_SWAP) {
    u256l t = src0;
    src0 = src1;
    src1 = t;
    uint32_t tp = src0_ptr;
    src0_ptr = src1_ptr;
    src1_ptr = tp;
  }
  const uint32_t new_pc = (pc + 1u) & 0xFFFFu;
  rowbits |= (src0_ptr ? ZKB_ROWBIT_SRC0_PTR : 0u) | (src1_ptr ? ZKB_ROWBIT_SRC1_PTR : 0u);
  // erase_fat_pointer_metadata (cycle.rs:374-396)
  if (!(entry & ZK_E_SRC0_PTR_OK) && src0_ptr && !kernel_mode) {
    src0 = lane < 4 ? src0 : 0u;
    src0_ptr = 0;
  }
  if (!(entry & ZK_E_SRC1_PTR_OK) && src1_ptr && !kernel_mode) {
    src1 = lane < 4 ? src1 : 0u;
    src1_ptr = 0;
  }
  S.row[8 + lane] = src0;
  S.row[16 + lane] = src1;
  const bool set_flags = entry & ZK_E_FLAG0;

  // ---- dispatch (parsing.rs:47-79) ----
  switch (family) {
    case ZK_OP_NOP:  // noop.rs
      pc = new_pc;
      break;
    case ZK_OP_ADD: {  // add.rs:35-43 (no flags reset; all three assigned)
      pc = new_pc;
      bool of;
      u256l r;
      // `add x, r0, dst` is the ISA's move: with r0 as the second operand there is no carry to resolve (two votes saved;
      // +2 % on ERC-20, where a quarter of the cycles are such moves)
      if (src1_reg == 0 && !(entry & ZK_E_SWAP)) {
        r = src0;
        of = false;
      } else {
        r = u_add(src0, src1, lane, of);
      }
      if (set_flags) {
        bool eq = u_is_zero(r);
        flags = (of ? 1u : 0u) | (eq ? 2u : 0u) | ((!eq && !of) ? 4u : 0u);
      }
      dst0_update(r, false);
      break;
    }
    case ZK_OP_SUB: {  // sub.rs:35-44
      pc = new_pc;
      bool of;
      u256l r = u_sub(src0, src1, lane, of);
      if (set_flags) {
        bool eq = u_is_zero(r);
        flags = (of ? 1u : 0u) | (eq ? 2u : 0u) | ((!eq && !of) ? 4u : 0u);
      }
      dst0_update(r, false);
      break;
    }
    case ZK_OP_MUL: {  // mul.rs:35-65
      pc = new_pc;
      u256l lo, hi;
      u_mul(src0, src1, lane, lo, hi);
      if (set_flags) {
        bool of = !u_is_zero(hi), eq = u_is_zero(lo);
        flags = (of ? 1u : 0u) | (eq ? 2u : 0u) | ((!of && !eq) ? 4u : 0u);
      }
      dst0_update(lo, false);
      dst1_update(hi, false);
      break;
    }
    case ZK_OP_DIV: {  // div.rs:36-75
      pc = new_pc;
      if (u_is_zero(src1)) {
        if (set_flags) flags = 1u;
        dst0_update(0u, false);
        dst1_update(0u, false);
      } else {
        u256l q, r;
        u_divmod(src0, src1, lane, q, r);
        if (set_flags) flags = (u_is_zero(q) ? 2u : 0u) | (u_is_zero(r) ? 4u : 0u);
        dst0_update(q, false);
        dst1_update(r, false);
      }
      break;
    }
    case ZK_OP_JUMP:  // jump.rs:24-25
      pc = oshfl(src0, 0) & 0xFFFFu;
      break;
    case ZK_OP_CONTEXT:
      pc = new_pc;
      op_context(sub, src0);
      break;
    case ZK_OP_SHIFT:
      pc = new_pc;
      op_shift(sub, src0, src1);
      break;
    case ZK_OP_BINOP: {  // binop.rs:42-51
      pc = new_pc;
      u256l r = sub == ZK_XOR ? (src0 ^ src1) : sub == ZK_AND ? (src0 & src1) : (src0 | src1);
      if (set_flags) flags = u_is_zero(r) ? 2u : 0u;
      dst0_update(r, false);
      break;
    }
    case ZK_OP_PTR:
      pc = new_pc;
      op_ptr(sub, src0, src1, src0_ptr, src1_ptr);
      break;
    case ZK_OP_NEAR_CALL:
      op_near_call(src0, new_pc);
      break;
    case ZK_OP_LOG:
      pc = new_pc;
      op_log(sub, src0, src1);
      break;
    case ZK_OP_FAR_CALL:
      op_far_call(sub, src0, src1, src0_ptr, new_pc, kernel_mode);
      break;
    case ZK_OP_RET:
      op_ret(sub, src0, src0_ptr);
      break;
    case ZK_OP_UMA:
      pc = new_pc;
      op_uma(sub, src0, src1, src0_ptr);
      break;
    default:
      fail(ZKB_VM_REFERENCE_PANIC);
      break;
  }
  // the reference returns Err / panics before end_execution_cycle (cycle.rs:406): no row for a cycle that stopped the VM
  if (ZK_UNLIKELY(status != ZKB_VM_RUNNING)) return;

  timestamp += ZK_TIME_DELTA_PER_CYCLE;
  cycle += 1;

  // ---- end_execution_cycle: emit the 256-byte row, two 16-byte stores per lane (each a full 128-byte line per VM) ----
  // lane 0 drops the scalar head of the row next to the operand values already staged in shared memory
  if (lane == 0) {
    *reinterpret_cast<uint4*>(&S.row[0]) = make_uint4(row_cycle, row_ts, raw_lo, raw_hi);
    *reinterpret_cast<uint4*>(&S.row[4]) = make_uint4(vidx | (resolved ? 1u : 0u) << 16 | err << 24, pc_before | pc << 16,
                                                      sp | flags << 16 | (rowbits | (pending ? ZKB_ROWBIT_PENDING : 0u)) << 24, ergs);
    S.row[L_COUNTS] = ccount;
    S.row[L_TX_PSP] = tx_psp;
  }
  if (ZK_UNLIKELY(n_rows >= B.cap[ZKB_STREAM_ROWS])) {
    fail(ZKB_VM_CAP_STREAM);
    return;
  }
  n_rows++;
  osync();
  if (B.witness) {
    const uint4 v0 = reinterpret_cast<const uint4*>(S.row)[lane], v1 = reinterpret_cast<const uint4*>(S.row)[8 + lane];
    uint4* dst = reinterpret_cast<uint4*>(row_ptr);
    dst[lane] = v0;
    dst[8 + lane] = v1;
  }
  row_ptr += ZKB_ROW_BYTES;
  // rare end-of-cycle state changes, keyed on the opcode family so that ordinary cycles pay one compare:
  if (family - ZK_OP_LOG <= 2u) {  // LOG, FAR_CALL, RET
    if (family == ZK_OP_LOG) {  // the cycle is complete; its precompile's result is not
      if (S.kc[KB_KC_PENDING]) status = ZKB_VM_YIELD_KECCAK;
      else if (S.kbuf[KB_EC_PENDING]) status = ZKB_VM_YIELD_ECRECOVER;
    }
    if (family == ZK_OP_RET && S.row[L_DEPTH] == 0) status = ZKB_VM_ENDED;              // execution_has_ended (mod.rs:96-98)
  }
}

// context.rs:36-99
template <bool KD>
__device__ __forceinline__ void Vm<KD>::op_context(uint32_t sub, u256l src0) {
  if (sub == ZK_CTX_SET_U128) {
    if (lane < 4) S.row[L_CTX + lane] = src0;
    osync();
    return;
  }
  if (sub == ZK_CTX_SET_ERGS_PER_PUBDATA) {
    setL(L_EPP, oshfl(src0, 0));
    return;
  }
  if (sub == ZK_CTX_INC_TX) {
    tx_psp = (tx_psp & 0xFFFF0000u) | ((tx_psp + 1u) & 0xFFFFu);
    return;
  }
  u256l v = 0u;
  switch (sub) {
    case ZK_CTX_THIS: v = addr_words_to_u256(lane < 5 ? S.F[F_THIS + lane] : 0u); break;
    case ZK_CTX_CALLER: v = addr_words_to_u256(lane < 5 ? S.F[F_SENDER + lane] : 0u); break;
    case ZK_CTX_CODE_ADDRESS: v = addr_words_to_u256(lane < 5 ? S.F[F_CODE_ADDR + lane] : 0u); break;
    case ZK_CTX_META: {  // VmMetaParameters::to_u256 (external layout; same reconstruction as the ISA table)
      uint32_t sh = S.F[F_EH_SHARDS], misc = S.F[F_MISC];
      uint32_t this_shard = (sh >> 16) & 0xFFu, caller_shard = sh >> 24, code_shard = misc & 0xFFu;
      v = lane == 0 ? L(L_EPP) : lane == 4 ? L(L_HEAP_BOUND) : lane == 5 ? L(L_AUX_BOUND)
          : lane == 7 ? (this_shard << 24 | caller_shard << 16 | code_shard << 8) : 0u;
      break;
    }
    case ZK_CTX_ERGS_LEFT: v = lane == 0 ? ergs : 0u; break;
    case ZK_CTX_SP: v = lane == 0 ? sp : 0u; break;
    case ZK_CTX_GET_U128: v = lane < 4 ? S.F[F_CTX + lane] : 0u; break;
    default: fail(ZKB_VM_REFERENCE_PANIC); return;
  }
  dst0_update(v, false);
}

// shift.rs:44-67
template <bool KD>
__device__ __forceinline__ void Vm<KD>::op_shift(uint32_t sub, u256l src0, u256l src1) {
  uint32_t n = oshfl(src1, 0) & 0xFFu;
  bool cyclic = sub == ZK_ROL || sub == ZK_ROR, right = sub == ZK_SHR || sub == ZK_ROR;
  u256l r;
  if (right) {
    r = u_shr(src0, n, lane);
    if (cyclic) r |= u_shl(src0, 256u - n, lane);
  } else {
    r = u_shl(src0, n, lane);
    if (cyclic) r |= u_shr(src0, 256u - n, lane);
  }
  if (entry & ZK_E_FLAG0) flags = u_is_zero(r) ? 2u : 0u;
  dst0_update(r, false);
}

// ptr.rs:32-193
template <bool KD>
__device__ __forceinline__ void Vm<KD>::op_ptr(uint32_t sub, u256l src0, u256l src1, bool src0_ptr, bool src1_ptr) {
  if (!src0_ptr || src1_ptr) {
    pending = 1;
    return;
  }
  uint32_t s1_nz = oballot(src1 != 0);
  uint32_t off1 = oshfl(src1, 0);
  if (sub == ZK_PTR_ADD || sub == ZK_PTR_SUB) {
    if (s1_nz & 0xFEu) {  // src1 >= 2^32 (MAX_OFFSET_FOR_ADD_SUB, ptr.rs:47)
      pending = 1;
      return;
    }
    uint32_t off0 = oshfl(src0, 0);
    uint32_t r = sub == ZK_PTR_ADD ? off0 + off1 : off0 - off1;
    bool of = sub == ZK_PTR_ADD ? r < off0 : off0 < off1;
    if (of) {
      pending = 1;
      return;
    }
    dst0_update(lane == 0 ? r : src0, true);
  } else if (sub == ZK_PTR_PACK) {
    if (s1_nz & 0x0Fu) {  // src1.low_u128() != 0 (ptr.rs:110)
      pending = 1;
      return;
    }
    dst0_update(lane < 4 ? src0 : src1, true);
  } else {  // Shrink (ptr.rs:140-192)
    uint32_t len = oshfl(src0, 3);
    if (len < off1) {
      pending = 1;
      return;
    }
    dst0_update(lane == 3 ? len - off1 : src0, true);
  }
}

// near_call.rs:6-68
template <bool KD>
__device__ __forceinline__ void Vm<KD>::op_near_call(u256l src0, uint32_t new_pc) {
  flags = 0;
  uint32_t abi_ergs = oshfl(src0, 0);
  uint32_t passed, remaining;
  if (abi_ergs == 0 || ergs < abi_ergs) {
    passed = ergs;
    remaining = 0;
  } else {
    passed = abi_ergs;
    remaining = ergs - abi_ergs;
  }
  ergs = remaining;
  pc = new_pc;
  if (!push_begin()) return;
  // the callee's frame is a clone of the caller's with is_local_frame set (near_call.rs:59-63)
  if (lane == 0) {
    S.F[F_SP_PC] = sp | imm0() << 16;
    S.F[F_EH_SHARDS] = (S.F[F_EH_SHARDS] & 0xFFFF0000u) | imm1();
    S.F[F_ERGS] = passed;
    S.F[F_MISC] |= 1u << 16;
  }
  push_end(ergs, pc, sp);
}

// log.rs:11-330
template <bool KD>
__device__ __forceinline__ void Vm<KD>::op_log(uint32_t sub, u256l src0, u256l src1) {
  const bool is_first = entry & ZK_E_FLAG0;
  const uint32_t shard = (S.F[F_EH_SHARDS] >> 16) & 0xFFu;
  const uint32_t ergs_available = ergs;
  const uint32_t ts_log = timestamp + 1;
  const uint32_t aw = lane < 5 ? S.F[F_THIS + lane] : 0u;
  const uint32_t epp = L(L_EPP);
  uint32_t ergs_on_pubdata = 0;
  if (sub == ZK_LOG_SSTORE) {
    // refund_for_partial_query (log.rs:99-102): Storage::estimate_refunds_for_write, asked BEFORE the write executes.
    // The reference's InMemoryStorage answers None (storage.rs:80-86); the refund-aware oracle (row f-3) answers
    // RepeatedWrite for a slot of the rollup shard whose cold/warm marker is already set.
    uint32_t refund = 0;
    if (ZK_UNLIKELY(B.warm_refund_bytes != 0u) && shard == 0) {
      if (oshfl(storage_access(shard, aw, src0, ST_IS_WARM, 0u), 0)) refund = B.warm_refund_bytes;
    }
    emit_refund(refund ? 1u : 0u, refund);
    uint32_t net = shard == 0 ? ZK_INITIAL_STORAGE_WRITE_PUBDATA_BYTES - refund : 0u;
    ergs_on_pubdata = epp * net;
  } else if (sub == ZK_LOG_TO_L1) {
    ergs_on_pubdata = epp * ZK_L1_MESSAGE_PUBDATA_BYTES;
  }
  uint32_t extra = sub == ZK_LOG_PRECOMPILE ? oshfl(src1, 0) : 0u;
  uint32_t total = extra + ergs_on_pubdata;
  bool not_enough = ergs_available < total;
  uint32_t spent = S.row[L_SPENT_PUBDATA];
  osync();
  if (not_enough) {
    ergs = 0;
    setL(L_SPENT_PUBDATA, spent + min(ergs_available, ergs_on_pubdata));
  } else {
    ergs = ergs_available - total;
    if (ergs_on_pubdata) setL(L_SPENT_PUBDATA, spent + ergs_on_pubdata);
  }
  switch (sub) {
    case ZK_LOG_SLOAD: {
      u256l v = storage_access(shard, aw, src0, ST_READ, 0u);
      emit_log(ts_log, ZK_STORAGE_AUX_BYTE, shard, aw, 0, is_first, src0, v, v);  // written := read (helpers.rs:145-148)
      dst0_update(v, false);
      break;
    }
    case ZK_LOG_SSTORE: {
      if (not_enough) return;
      u256l old = storage_access(shard, aw, src0, ST_WRITE, src1);
      emit_log(ts_log, ZK_STORAGE_AUX_BYTE, shard, aw, 1, is_first, src0, old, src1);
      break;
    }
    case ZK_LOG_EVENT:
    case ZK_LOG_TO_L1: {
      if (not_enough) return;
      emit_log(ts_log, sub == ZK_LOG_EVENT ? ZK_EVENT_AUX_BYTE : ZK_L1_MESSAGE_AUX_BYTE, shard, aw, 1, is_first, src0, 0u, src1);
      break;
    }
    default: {  // PrecompileCall (log.rs:252-328)
      if (not_enough) {
        dst0_update(0u, false);
        return;
      }
      uint32_t heap_page = L(L_BASE_PAGE) + 2;
      u256l abi = src0;
      if (lane == 4 && abi == 0) abi = heap_page;  // memory_page_to_read
      if (lane == 5 && abi == 0) abi = heap_page;  // memory_page_to_write
      emit_log(ts_log, ZK_PRECOMPILE_AUX_BYTE, shard, aw, 0, is_first, abi, 0u, 0u);
      uint32_t addr_low = bswap32(S.F[F_THIS + 4]) & 0xFFFFu;
      if (addr_low == ZK_KECCAK256_PRECOMPILE_ADDRESS) {
        keccak_precompile(abi);
      } else if (addr_low == ZK_SHA256_PRECOMPILE_ADDRESS) {
        sha256_precompile(abi);
      } else if (addr_low == ZK_ECRECOVER_PRECOMPILE_ADDRESS) {
        ecrecover_precompile(abi);
      }
      if (status != ZKB_VM_RUNNING) return;
      u256l one = lane == 0 ? 1u : 0u;
      dst0_update(one, false);
      break;
    }
  }
}

// keccak256 precompile (external DefaultPrecompilesProcessor; memory ABI pinned by keccak256.rs:100-139):
// byte offset/length in, one output word (word index) out; one FatPointer-type read per distinct input word.
// The cycle itself only moves data, like ecrecover's: it emits the read witness (one record per input word, in order),
// reserves the output word and its write record (placeholder value) and leaves a descriptor in the VM's scratch; the
// sponge runs AFTER the cycle (run_deferred_keccak, one thread per state, batched over the CTA) and patches the digest
// into the heap word and the record.  Nothing inside the cycle depends on the digest.
template <bool KD>
__device__ __forceinline__ void Vm<KD>::keccak_precompile(u256l abi) {
  const uint32_t in_off = oshfl(abi, 0), in_len = oshfl(abi, 1);
  const uint32_t out_word = oshfl(abi, 2);
  const uint32_t page_read = oshfl(abi, 4), page_write = oshfl(abi, 5);
  const uint32_t ts_read = timestamp + 1, ts_write = timestamp + 2;
  if (in_len < 2u * 136u) {  // at most two permutations: in place
    keccak_precompile_inline(abi);
    return;
  }
  // resolve the source page once (fat-pointer indirection, memory.rs:475-521)
  uint32_t src_slab = ZKB_NO_SLAB;
  if (in_len > 0 && page_read != 0) {
    int e = pt_find(page_read);
    if (e < 0) {
      fail(ZKB_VM_REFERENCE_PANIC);
      return;
    }
    uint32_t info = S.pt[e * 2 + 1];
    uint32_t kind = info & 0xFFu, x = (info >> 8) & 0xFFu;
    src_slab = kind == PT_EXT ? x : level_slab(x, kind == PT_AUX_LIVE ? 1 : 0);
  }
  if (in_len > 0) {
    const uint64_t w_first = in_off / 32, w_last = ((uint64_t)in_off + in_len - 1) / 32;
    for (uint64_t w = w_first; w <= w_last; w++) {
      const u256l word = slab_read(src_slab, (uint32_t)min(w, (uint64_t)0xFFFFFFFFu));
      emit_mem(ts_read, page_read, (uint32_t)w, ZK_MEM_FAT_PTR, 0, 0, ZKB_MEMORIGIN_PRECOMPILE_IN, word);
      if (ZK_UNLIKELY(status != ZKB_VM_RUNNING)) return;
    }
  }
  // the write goes through MemoryType::Heap: the reference checks the page only by debug_assert (memory.rs:447)
  if (page_write != L(L_BASE_PAGE) + 2) {
    fail(ZKB_VM_REFERENCE_PANIC);
    return;
  }
  uint32_t s = cur_slab(0, true, out_word);
  if (status != ZKB_VM_RUNNING) return;
  const uint32_t record = n_mem;
  slab_write(s, out_word, 0u);
  emit_mem(ts_write, page_write, out_word, ZK_MEM_HEAP, 1, 0, ZKB_MEMORIGIN_PRECOMPILE_OUT, 0u);
  if (status != ZKB_VM_RUNNING) return;
  osync();
  if (lane == 0) {
    S.kc[KB_KC_PENDING] = 1u;
    S.kc[KB_KC_IN_OFF] = in_off;
    S.kc[KB_KC_IN_LEN] = in_len;
    S.kc[KB_KC_SRC_SLAB] = src_slab;
    S.kc[KB_KC_OUT_SLAB] = s;
    S.kc[KB_KC_OUT_WORD] = out_word;
    S.kc[KB_KC_MEM_INDEX] = record;
    S.kc[KB_KC_VM] = vm;
  }
  osync();
}

// The in-cycle variant for SHORT inputs (at most two rate blocks, e.g. the 64-byte mapping-slot preimages of a token
// contract): the octet absorbs and permutes on the spot (keccak.cuh, octet-cooperative layout).  One or two
// permutations are cheaper here than a CTA-wide deferred phase, whose cost is the single-thread latency of a permutation.
template <bool KD>
__device__ __forceinline__ void Vm<KD>::keccak_precompile_inline(u256l abi) {
  const uint32_t in_off = oshfl(abi, 0), in_len = oshfl(abi, 1);
  const uint32_t out_word = oshfl(abi, 2);
  const uint32_t page_read = oshfl(abi, 4), page_write = oshfl(abi, 5);
  const uint32_t ts_read = timestamp + 1, ts_write = timestamp + 2;
  // resolve the source page once (fat-pointer indirection, memory.rs:475-521)
  uint32_t src_slab = ZKB_NO_SLAB;
  if (in_len > 0 && page_read != 0) {
    int e = pt_find(page_read);
    if (e < 0) {
      fail(ZKB_VM_REFERENCE_PANIC);
      return;
    }
    uint32_t info = S.pt[e * 2 + 1];
    uint32_t kind = info & 0xFFu, x = (info >> 8) & 0xFFu;
    src_slab = kind == PT_EXT ? x : level_slab(x, kind == PT_AUX_LIVE ? 1 : 0);
  }
  const KeccakLanes kl = keccak_lanes(lane);
  KeccakState st;
#pragma unroll
  for (int y = 0; y < 5; y++) st.a[y] = 0;
  const uint64_t end = (uint64_t)in_off + in_len;
  uint64_t next_emit_word = in_off / 32;
  const uint32_t n_blocks = in_len / 136 + 1;
  for (uint32_t blk = 0; blk < n_blocks; blk++) {
    const uint64_t a0 = (uint64_t)in_off + (uint64_t)blk * 136;
    const uint32_t nb = (uint32_t)min((uint64_t)136, end - a0);  // valid bytes in this block
    const uint64_t w0 = a0 / 32;
    if (nb > 0) {
      const uint64_t w1 = (a0 + nb - 1) / 32;
      for (uint64_t w = w0; w <= w1; w++) {
        u256l word = slab_read(src_slab, (uint32_t)min(w, (uint64_t)0xFFFFFFFFu));
        if (w >= next_emit_word) {
          emit_mem(ts_read, page_read, (uint32_t)w, ZK_MEM_FAT_PTR, 0, 0, ZKB_MEMORIGIN_PRECOMPILE_IN, word);
          next_emit_word = w + 1;
        }
        S.kbuf[(uint32_t)(w - w0) * 8 + (7 - lane)] = bswap32(word);  // byte stream order
      }
    }
    osync();
    // absorb: rate word i (bytes [8 i, 8 i + 8) of the block, little-endian) belongs to column i % 5, row i / 5
    if (lane < 5) {
#pragma unroll
      for (int y = 0; y < 4; y++) {
        const uint32_t wi = lane + 5u * (uint32_t)y;
        if (wi < 17) {
          uint32_t o = (uint32_t)(a0 - w0 * 32) + 8 * wi;  // byte offset into kbuf
          uint32_t i = o >> 2, sh = (o & 3u) * 8;
          uint32_t x0 = S.kbuf[i], x1 = S.kbuf[i + 1], x2 = S.kbuf[(i + 2) & 63];
          uint32_t lo = __funnelshift_r(x0, x1, sh), hi = __funnelshift_r(x1, x2, sh);
          uint64_t v = ((uint64_t)hi << 32) | lo;
          uint32_t b0 = 8 * wi;
          if (b0 >= nb) v = 0;
          else if (b0 + 8 > nb) v &= (1ull << (8 * (nb - b0))) - 1ull;
          if (nb < 136) {  // final block: pad10*1 with the keccak domain byte 0x01
            if (wi == nb / 8) v ^= 1ull << (8 * (nb % 8));
            if (wi == 16) v ^= 0x80ull << 56;
          }
          st.a[y] ^= v;
        }
      }
    }
    osync();
    keccak_f1600(st, kl, S.ks(), lane);
  }
  // digest = first 32 bytes of the state (row 0 of columns 0..3, little-endian lanes) read as one big-endian word
  int t = 7 - (int)lane;
  uint32_t lo = oshfl((uint32_t)st.a[0], (t >> 1) & 7), hi = oshfl((uint32_t)(st.a[0] >> 32), (t >> 1) & 7);
  u256l digest = bswap32((t & 1) ? hi : lo);
  // the write goes through MemoryType::Heap: the reference checks the page only by debug_assert (memory.rs:447)
  if (page_write != L(L_BASE_PAGE) + 2) {
    fail(ZKB_VM_REFERENCE_PANIC);
    return;
  }
  uint32_t s = cur_slab(0, true, out_word);
  if (status != ZKB_VM_RUNNING) return;
  slab_write(s, out_word, digest);
  emit_mem(ts_write, page_write, out_word, ZK_MEM_HEAP, 1, 0, ZKB_MEMORIGIN_PRECOMPILE_OUT, digest);
}

// sha256 precompile (external DefaultPrecompilesProcessor; memory ABI reconstructed): the caller
// passes pre-padded 64-byte blocks: input offset in WORDS, number of rounds in precompile_interpreted_data, two
// Heap-type word reads per round at timestamp+1, one digest word written at timestamp+2 after the last round.
template <bool KD>
__device__ __forceinline__ void Vm<KD>::sha256_precompile(u256l abi) {
  const uint32_t in_word = oshfl(abi, 0), out_word = oshfl(abi, 2);
  const uint32_t page_read = oshfl(abi, 4), page_write = oshfl(abi, 5);
  const uint64_t rounds = (uint64_t)oshfl(abi, 6) | (uint64_t)oshfl(abi, 7) << 32;
  const uint32_t ts_read = timestamp + 1, ts_write = timestamp + 2;
  if (rounds == 0) return;
  const uint32_t heap_page = L(L_BASE_PAGE) + 2;
  // MemoryType::Heap queries address the current frame's heap; the reference checks the page number (memory.rs:447)
  if (page_read != heap_page || page_write != heap_page) {
    fail(ZKB_VM_REFERENCE_PANIC);
    return;
  }
  const uint32_t src_slab = S.lv[0];
  S.kbuf[lane] = c_sha256_iv[lane];
  for (uint64_t r = 0; r < rounds; r++) {
#pragma unroll
    for (uint32_t k = 0; k < 2; k++) {
      const uint32_t idx = in_word + 2u * (uint32_t)r + k;
      u256l word = slab_read(src_slab, idx);
      emit_mem(ts_read, page_read, idx, ZK_MEM_HEAP, 0, 0, ZKB_MEMORIGIN_PRECOMPILE_IN, word);
      if (status != ZKB_VM_RUNNING) return;
      S.kbuf[8 + 8 * k + (7 - lane)] = word;  // big-endian message words: M[j] = limb[7 - j]
    }
#ifndef ZKB_NO_SHA_CALL   // experiment: what the presence of the callee costs the loop
    sha256_compress_smem(S.kbuf, lane);
#endif
  }
  osync();
  u256l digest = S.kbuf[7 - lane];
  uint32_t s = cur_slab(0, true, out_word);
  if (status != ZKB_VM_RUNNING) return;
  slab_write(s, out_word, digest);
  emit_mem(ts_write, page_write, out_word, ZK_MEM_HEAP, 1, 0, ZKB_MEMORIGIN_PRECOMPILE_OUT, digest);
}

// ecrecover precompile (external DefaultPrecompilesProcessor; memory ABI reconstructed): four
// Heap-type word reads at input_memory_offset (digest, v in {0, 1}, r, s) at timestamp+1, two Heap-type word writes at
// output_memory_offset (success marker, address; both zero on failure) at timestamp+2.
// The cycle itself only moves data: it emits the six memory queries (the two writes with placeholder values), reserves
// the output words and stashes the inputs; the ~6 000 modular multiplications of the recovery run AFTER the cycle, with
// the VM state parked in HBM (run_deferred_ecrecover), and patch the two values in place.  Nothing inside the cycle
// depends on them, and the interpreter's hot loop carries no live state across the call.
template <bool KD>
__device__ __forceinline__ void Vm<KD>::ecrecover_precompile(u256l abi) {
  const uint32_t in_word = oshfl(abi, 0), out_word = oshfl(abi, 2);
  const uint32_t page_read = oshfl(abi, 4), page_write = oshfl(abi, 5);
  const uint32_t ts_read = timestamp + 1, ts_write = timestamp + 2;
  const uint32_t heap_page = L(L_BASE_PAGE) + 2;
  if (page_read != heap_page) {  // MemoryType::Heap queries address the current frame's heap (memory.rs:447)
    fail(ZKB_VM_REFERENCE_PANIC);
    return;
  }
  const uint32_t src_slab = S.lv[0];
  u256l v_word = 0u;
#pragma unroll
  for (uint32_t i = 0; i < 4; i++) {
    u256l w = slab_read(src_slab, in_word + i);
    emit_mem(ts_read, page_read, in_word + i, ZK_MEM_HEAP, 0, 0, ZKB_MEMORIGIN_PRECOMPILE_IN, w);
    S.kbuf[KB_EC_INPUT + 8 * i + lane] = w;
    if (i == 1) v_word = w;
  }
  if (status != ZKB_VM_RUNNING) return;
  const uint32_t v_nz = oballot(v_word != 0), v0 = oshfl(v_word, 0);
  if ((v_nz & 0xFEu) || v0 > 1u || page_write != heap_page) {  // the external precompile asserts v == 0 || v == 1
    fail(ZKB_VM_REFERENCE_PANIC);
    return;
  }
  uint32_t s = cur_slab(0, true, out_word + 1);
  if (status != ZKB_VM_RUNNING) return;
  const uint32_t first_record = n_mem;
  slab_write(s, out_word, 0u);
  emit_mem(ts_write, page_write, out_word, ZK_MEM_HEAP, 1, 0, ZKB_MEMORIGIN_PRECOMPILE_OUT, 0u);
  slab_write(s, out_word + 1, 0u);
  emit_mem(ts_write, page_write, out_word + 1, ZK_MEM_HEAP, 1, 0, ZKB_MEMORIGIN_PRECOMPILE_OUT, 0u);
  if (status != ZKB_VM_RUNNING) return;
  if (lane == 0) {
    S.kbuf[KB_EC_PENDING] = 1u;
    S.kbuf[KB_EC_MEM_INDEX] = first_record;
    S.kbuf[KB_EC_SLAB] = s;
    S.kbuf[KB_EC_OUT_WORD] = out_word;
  }
  osync();
}

// SimpleMemory::start_global_frame (memory.rs:573-657) for the level far_depth() (already incremented)
template <bool KD>
__device__ __forceinline__ void Vm<KD>::memory_start_global_frame(uint32_t caller_level, uint32_t caller_base, uint32_t calldata_page) {
  uint32_t level = far_depth();
  // the caller's level entry goes back to HBM, the callee starts with no heaps and an untouched stack page
  if (lane < 4) {
    g_lvl()[caller_level * 4 + lane] = S.lv[lane];
    S.lv[lane] = lane < 2 ? ZKB_NO_SLAB : 0u;
  }
  osync();
  // the root "heaps" entry has page numbers 0/0 (memory.rs:230-233)
  uint32_t cur_heap = caller_level == 0 ? 0u : caller_base + 2, cur_aux = caller_level == 0 ? 0u : caller_base + 3;
  if (calldata_page == 0) {
  } else if (calldata_page == cur_heap) {
    pt_upsert(cur_heap, PT_HEAP_LIVE, caller_level, level);
  } else if (calldata_page == cur_aux) {
    pt_upsert(cur_aux, PT_AUX_LIVE, caller_level, level);
  } else {
    int e = pt_find(calldata_page);
    if (e < 0) {
      fail(ZKB_VM_REFERENCE_PANIC);  // "fat pointer must only point to reachable memory" (memory.rs:641)
      return;
    }
    uint32_t kind = S.pt[e * 2 + 1] & 0xFFu;
    if (kind != PT_HEAP_LIVE && kind != PT_AUX_LIVE) fail(ZKB_VM_REFERENCE_PANIC);  // memory.rs:645
  }
}

// SimpleMemory::finish_global_frame (memory.rs:660-758)
template <bool KD>
__device__ __forceinline__ void Vm<KD>::memory_finish_global_frame(uint32_t level, uint32_t base_page, uint32_t returndata_page) {
  // stack page goes back to the pool: clear what was touched (stack_on_return, memory.rs:185-188)
  uint32_t hwm = S.lv[2];  // S.lv still holds the finished level's entry (the parent's is reloaded at the end)
  const size_t fpage = (size_t)vm * (B.max_far_depth + 1) + level;  // g_stack() already points at the caller's page
  uint32_t* sbase = B.stack_mem + fpage * B.stack_words * 8;
  uint8_t* pbase = B.stack_ptr + fpage * B.stack_words;
  uint4* sbase4 = reinterpret_cast<uint4*>(sbase);
  for (uint32_t i = lane; i < hwm * 2; i += 8) sbase4[i] = make_uint4(0u, 0u, 0u, 0u);
  for (uint32_t i = lane; i < hwm; i += 8) pbase[i] = 0;
  uint32_t heap_slab = S.lv[0], aux_slab = S.lv[1];
  osync();
  uint32_t heap_page = base_page + 2, aux_page = base_page + 3;
  if (returndata_page == heap_page) {
    pt_upsert(heap_page, PT_EXT, heap_slab, level - 1);
    slab_release(aux_slab);
  } else if (returndata_page == aux_page) {
    pt_upsert(aux_page, PT_EXT, aux_slab, level - 1);
    slab_release(heap_slab);
  } else {
    if (returndata_page != 0) {
      int e = pt_find(returndata_page);
      if (e < 0) {
        fail(ZKB_VM_REFERENCE_PANIC);  // memory.rs:735
        return;
      }
      if (lane == 0) S.pt[e * 2 + 1] = (S.pt[e * 2 + 1] & 0xFFFFu) | (level - 1) << 16;
      osync();
    }
    slab_release(heap_slab);
    slab_release(aux_slab);
  }
  // drop every indirection still owned by the finished level (octet lane l holds entries l, l + 8, l + 16, l + 24)
#pragma unroll 1
  for (uint32_t q = 0; q < ZKB_PT_ENTRIES / ZK_OCT; q++) {
    const uint32_t e0 = lane + ZK_OCT * q;
    const uint32_t page = S.pt[e0 * 2], info = S.pt[e0 * 2 + 1];
    uint32_t drop = oballot(page != ZKB_PT_FREE && (info >> 16) == level);
    while (drop) {
      int l = __ffs(drop) - 1;
      drop &= drop - 1;
      uint32_t einfo = oshfl(info, l);
      if ((einfo & 0xFFu) == PT_EXT) slab_release((einfo >> 8) & 0xFFu);
      if ((int)lane == l) S.pt[e0 * 2] = ZKB_PT_FREE;  // the lane that read the entry frees it (no cross-lane write)
    }
  }
  osync();
  if (lane < 4) S.lv[lane] = g_lvl()[far_depth() * 4 + lane];  // back in the caller's far level
  osync();
}

// far_call.rs:35-613
template <bool KD>
__device__ __forceinline__ void Vm<KD>::op_far_call(uint32_t sub, u256l src0, u256l src1, bool abi_is_ptr, uint32_t new_pc, bool kernel_mode) {
  enum { EX_NOT_PTR = 1, EX_HASH_FORMAT = 2, EX_ERGS_DECOMMIT = 4, EX_ERGS_GROW = 8, EX_MALFORMED_PTR = 16, EX_CONSTRUCTED_SYSTEM = 32 };
  flags = 0;
  const bool is_call_shard = entry & ZK_E_FLAG0, is_static_call = entry & ZK_E_FLAG1;
  const uint32_t eh = imm0();
  // called address / kernel test
  const uint32_t dest_nz = oballot(src1 != 0) & 0x1Fu;  // limbs 0..4 = low 160 bits
  const bool dst_is_kernel = (dest_nz & 0x1Eu) == 0 && oshfl(src1, 0) < 65536u;
  const u256l dest_key = lane < 5 ? src1 : 0u;             // value & U256_TO_ADDRESS_MASK
  const uint32_t dest_aw = u256_to_addr_words(src1);       // lanes 0..4
  // FarCallABI::from_u256
  uint32_t p_off = oshfl(src0, 0), p_page = oshfl(src0, 1);
  uint32_t p_start = oshfl(src0, 2), p_len = oshfl(src0, 3);
  const uint32_t abi_ergs = oshfl(src0, 6), top = oshfl(src0, 7);
  const uint32_t fwd_byte = top & 0xFFu, abi_shard = (top >> 8) & 0xFFu;
  const uint32_t fwd = fwd_byte == ZK_FWD_FORWARD_FAT_POINTER ? ZK_FWD_FORWARD_FAT_POINTER : fwd_byte == ZK_FWD_USE_AUX_HEAP ? ZK_FWD_USE_AUX_HEAP : ZK_FWD_USE_HEAP;
  const bool constructor_call = ((top >> 16) & 0xFFu) != 0 && kernel_mode;
  const bool to_system = (top >> 24) != 0 && dst_is_kernel;

  const uint32_t cur_base = L(L_BASE_PAGE);
  const uint32_t shards = S.F[F_EH_SHARDS];
  const uint32_t caller_shard = (shards >> 16) & 0xFFu;
  const uint32_t remaining_ergs = ergs;
  const uint32_t new_code_shard = is_call_shard ? abi_shard : caller_shard;
  const uint32_t new_this_shard = sub == ZK_FC_DELEGATE ? caller_shard : new_code_shard;
  const uint32_t new_base = L(L_PAGE_COUNTER);
  const uint32_t ts1 = timestamp + 1;

  u256l code_hash;
  bool map_to_trivial;
  if (new_code_shard != 0 && !B.zkporter) {
    code_hash = 0u;
    map_to_trivial = true;
  } else {
    if (new_code_shard >= 2) {
      fail(ZKB_VM_REFERENCE_PANIC);  // InMemoryStorage has NUM_SHARDS = 2 (storage.rs:93 index)
      return;
    }
    const uint32_t deployer_aw = lane == 4 ? bswap32(ZK_DEPLOYER_SYSTEM_CONTRACT_ADDRESS) : 0u;
    u256l v = storage_access(new_code_shard, deployer_aw, dest_key, ST_READ, 0u);
    emit_log(ts1, ZK_STORAGE_AUX_BYTE, new_code_shard, deployer_aw, 0, 0, dest_key, v, v);
    bool mask_aa = u_is_zero(v) && !dst_is_kernel;
    code_hash = mask_aa ? B.default_aa[lane] : v;
    map_to_trivial = false;
  }
  const uint32_t page_candidate = map_to_trivial ? ZK_UNMAPPED_PAGE : new_base;

  uint32_t ex = 0;
  uint32_t code_len_words = 0;
  {
    uint32_t top_limb = oshfl(code_hash, 7);
    uint32_t version = top_limb >> 24, marker = (top_limb >> 16) & 0xFFu;
    if (version == ZK_CODE_HASH_VERSION_BYTE) {
      bool at_rest = marker == ZK_CODE_AT_REST_MARKER, constructed_now = marker == ZK_YET_CONSTRUCTED_MARKER;
      if (!(at_rest || constructed_now)) {
        ex |= EX_HASH_FORMAT;
        code_hash = 0u;
      } else if ((!constructor_call && at_rest) || (constructor_call && constructed_now)) {
        if (lane == 7) code_hash &= 0xFF00FFFFu;  // serialize_to_stored: marker := at rest
        code_len_words = top_limb & 0xFFFFu;
      } else if (!dst_is_kernel) {
        uint32_t aa_top = B.default_aa[7];
        if ((aa_top >> 24) != ZK_CODE_HASH_VERSION_BYTE || ((aa_top >> 16) & 0xFFu) != ZK_CODE_AT_REST_MARKER) {
          fail(ZKB_VM_REFERENCE_PANIC);  // far_call.rs:222-227
          return;
        }
        code_hash = B.default_aa[lane];
        code_len_words = aa_top & 0xFFFFu;
      } else {
        ex |= EX_CONSTRUCTED_SYSTEM;
        code_hash = 0u;
      }
    } else {
      ex |= EX_HASH_FORMAT;
      code_hash = 0u;
    }
  }
  if (fwd == ZK_FWD_FORWARD_FAT_POINTER && !abi_is_ptr) ex |= EX_NOT_PTR;
  // FatPointer::validate / validate_as_slice
  const bool fresh = fwd != ZK_FWD_FORWARD_FAT_POINTER;
  const bool deref_beyond = (uint64_t)p_start + (uint64_t)p_len > 0xFFFFFFFFull;
  if ((fresh && p_off != 0) || deref_beyond) ex |= EX_MALFORMED_PTR;
  if (p_off > p_len) ex |= EX_MALFORMED_PTR;
  if (fwd == ZK_FWD_FORWARD_FAT_POINTER) {
    p_start += p_off;
    p_len -= p_off;
    p_off = 0;
  } else {
    p_page = cur_base + (fwd == ZK_FWD_USE_HEAP ? 2u : 3u);
  }
  if (ex) p_off = p_page = p_start = p_len = 0;

  uint32_t growth = 0, bound_kind = 0, bound_value = 0;
  if (fwd != ZK_FWD_FORWARD_FAT_POINTER) {
    uint32_t upper = p_start + p_len;
    if (deref_beyond) upper = 0xFFFFFFFFu;
    int w = fwd == ZK_FWD_USE_HEAP ? L_HEAP_BOUND : L_AUX_BOUND;
    uint32_t bound = S.row[w];
    osync();
    if (upper >= bound) {
      growth = upper - bound;
      setL(w, upper);
      bound = upper;
    }
    bound_kind = fwd == ZK_FWD_USE_HEAP ? 1u : 2u;  // FrameRec.prev_bound_*: the caller's bound as the tracer sees it
    bound_value = bound;
  }
  uint32_t ergs_after_growth;
  if (remaining_ergs >= growth) {
    ergs_after_growth = remaining_ergs - growth;
  } else {
    ex |= EX_ERGS_GROW;
    ergs_after_growth = 0;
  }
  const uint32_t decommit_cost = ZK_ERGS_PER_CODE_WORD_DECOMMITTMENT * code_len_words;
  uint32_t ergs_after_decommit;
  if (ergs_after_growth >= decommit_cost) {
    ergs_after_decommit = ergs_after_growth - decommit_cost;
  } else {
    ex |= EX_ERGS_DECOMMIT;
    ergs_after_decommit = ergs_after_growth;
  }
  uint32_t mapped_code_page, new_code_id = ZKB_NO_CODE;
  if (ex) {
    pending = 1;
    mapped_code_page = ZK_UNMAPPED_PAGE;
  } else {
    // SimpleDecommitter::decommit_into_memory (decommitter.rs:32-99)
    // lookup by hash: open-addressed index over the loaded bytecodes (built by the host at upload: power-of-two slots, at
    // most half full, keyed by the hash's low limb) -- round 1 scanned all bytecodes linearly on every fresh far call
    int id = -1;
    for (uint32_t slot = oshfl(code_hash, 0) & B.code_index_mask, probes = 0; probes <= B.code_index_mask; slot = (slot + 1u) & B.code_index_mask, probes++) {
      const uint32_t c = B.code_index[slot];
      if (c == ZKB_NO_CODE) break;
      if (u_eq(B.code_meta[c * 10 + 2 + lane], code_hash)) {
        id = (int)c;
        break;
      }
    }
    uint32_t* dec = B.dec + (size_t)vm * ZKB_DEC_ENTRIES * 2;
    // history is keyed by hash; entries created by populate_code are flagged (bit 31) and not part of it.
    // Octet lane l looks at entries l and l + 8 (ZKB_DEC_ENTRIES = 16).
    const uint32_t e_lo = lane < n_decommit() ? dec[lane * 2] : ZKB_NO_CODE, e_hi = lane + 8 < n_decommit() ? dec[(lane + 8) * 2] : ZKB_NO_CODE;
    const bool have_id = id >= 0;
    uint32_t hist = oballot(have_id && e_lo == (uint32_t)id) | oballot(have_id && e_hi == (uint32_t)id) << 8;
    uint32_t fresh_flag, len16;
    if (hist) {
      mapped_code_page = dec[(__ffs(hist) - 1) * 2 + 1];
      fresh_flag = 0;
      len16 = B.code_meta[id * 10 + 1] & 0xFFFFu;
      ergs_after_decommit += decommit_cost;  // refund (far_call.rs:450-453)
    } else {
      if (id < 0) {
        status = ZKB_VM_UNKNOWN_CODE_HASH;  // anyhow::Err (decommitter.rs:50-56)
        return;
      }
      if (n_decommit() >= ZKB_DEC_ENTRIES) {
        fail(ZKB_VM_CAP_PAGES);
        return;
      }
      if (lane == 0) {
        dec[n_decommit() * 2] = (uint32_t)id;
        dec[n_decommit() * 2 + 1] = page_candidate;
      }
      n_decommit()++;
      osync();
      mapped_code_page = page_candidate;
      fresh_flag = 1;
      len16 = B.code_meta[id * 10 + 1] & 0xFFFFu;
    }
    new_code_id = (uint32_t)id;
    emit_decommit(ts1, mapped_code_page, len16, fresh_flag, code_hash);
  }
  // 63/64 rule (far_call.rs:468-487)
  const uint32_t max_passable = (ergs_after_decommit / 64u) * 63u;
  const uint32_t leftover = ergs_after_decommit - max_passable;
  uint32_t passed, remaining_for_this;
  if (max_passable < abi_ergs) {
    passed = max_passable;
    remaining_for_this = leftover;
  } else {
    passed = abi_ergs;
    remaining_for_this = leftover + (max_passable - abi_ergs);
  }
  ergs = remaining_for_this;
  pc = new_pc;
  const bool new_static = is_static() || is_static_call;
  uint32_t page_counter = S.row[L_PAGE_COUNTER];
  osync();
  setL(L_PAGE_COUNTER, page_counter + ZK_NEW_MEMORY_PAGES_PER_FAR_CALL);

  // new frame, lane i = word i
  const uint32_t r15_aw = u256_to_addr_words(reg_read(ZK_CALL_IMPLICIT_PARAMETER_REG_IDX + 1));
  const uint32_t caller_level = far_depth();
  if (far_depth() + 1 > B.max_far_depth) {
    fail(ZKB_VM_CAP_DEPTH);
    return;
  }
  if (!push_begin()) return;  // the caller's frame is saved and still readable in S.F
  {
    const uint32_t prev_ergs = ergs, prev_pc = pc, prev_sp = sp;
    const uint32_t this_w = lane < 5 ? S.F[F_THIS + lane] : 0u;  // lanes 0..4
    const uint32_t sender_w = lane < 5 ? S.F[F_SENDER + lane] : 0u;
    const uint32_t next_this = sub == ZK_FC_DELEGATE ? this_w : dest_aw;
    const uint32_t next_sender = sub == ZK_FC_NORMAL ? this_w : sub == ZK_FC_DELEGATE ? sender_w : r15_aw;
    const uint32_t ctx = lane < 4 ? (sub == ZK_FC_DELEGATE ? S.F[F_CTX + lane] : S.row[L_CTX + lane]) : 0u;
    osync();
    if (lane < 5) {
      S.F[F_THIS + lane] = next_this;
      S.F[F_SENDER + lane] = next_sender;
      S.F[F_CODE_ADDR + lane] = dest_aw;
    }
    if (lane < 4) {
      S.F[F_CTX + lane] = ctx;
      S.row[L_CTX + lane] = 0u;  // context_u128_register = 0 (far_call.rs:558)
    }
    if (lane == 5) {
      S.F[F_BASE_PAGE] = new_base;
      S.F[F_CODE_PAGE] = mapped_code_page;
      S.F[F_SP_PC] = ZK_INITIAL_SP_ON_FAR_CALL;
      S.F[F_EH_SHARDS] = eh | new_this_shard << 16 | caller_shard << 24;
      S.F[F_ERGS] = passed;
      S.F[F_MISC] = new_code_shard | (new_static ? 1u : 0u) << 8;
    }
    if (lane == 6) {
      S.F[F_HEAP_BOUND] = ZK_NEW_FRAME_MEMORY_STIPEND;
      S.F[F_AUX_BOUND] = ZK_NEW_FRAME_MEMORY_STIPEND;
      S.F[F_CODE_ID] = new_code_id;
      S.F[F_JOURNAL_MARK] = 0u;
      S.F[F_FAR_LEVEL] = caller_level + 1;
      S.F[30] = 0u;
      S.F[31] = 0u;
    }
    push_end(prev_ergs, prev_pc, prev_sp, bound_kind, bound_value);
  }
  if (status != ZKB_VM_RUNNING) return;
  memory_start_global_frame(caller_level, cur_base, p_page);

  // register ABI (far_call.rs:573-610)
  u256l r1 = lane == 0 ? p_off : lane == 1 ? p_page : lane == 2 ? p_start : lane == 3 ? p_len : 0u;
  u256l r2 = lane == 0 ? ((constructor_call ? 1u : 0u) | (to_system ? 2u : 0u)) : 0u;
  S.regs[1][lane] = r1;
  S.regs[2][lane] = r2;
  S.row[24 + lane] = r1;
  S.row[32 + lane] = r2;
  // registers[] index i <-> r(i+1): system ABI regs r3..r12, reserved r13,r14, implicit r15
  uint32_t clear_mask = 0xE000u | (to_system ? 0u : 0x1FF8u);
  for (uint32_t r = 3; r < 16; r++)
    if ((clear_mask >> r) & 1u) S.regs[r][lane] = 0u;
  ptr_mask = 0x0002u;  // only r1 is a pointer: r2 plain, r3..r12 markers removed or zeroed, r13..r15 zeroed
  rowbits |= ZKB_ROWBIT_DST0_VALID | ZKB_ROWBIT_DST0_PTR | ZKB_ROWBIT_DST1_VALID;
  osync();
}

// ret.rs:9-265
template <bool KD>
__device__ __forceinline__ void Vm<KD>::op_ret(uint32_t sub, u256l src0, bool src0_ptr) {
  uint32_t variant = sub;
  flags = 0;
  if (variant == ZK_RET_PANIC) {
    src0 = 0u;
    src0_ptr = false;
  }
  uint32_t p_off = oshfl(src0, 0), p_page = oshfl(src0, 1);
  uint32_t p_start = oshfl(src0, 2), p_len = oshfl(src0, 3);
  const uint32_t fwd_byte = oshfl(src0, 7) & 0xFFu;
  const uint32_t fwd = fwd_byte == ZK_FWD_FORWARD_FAT_POINTER ? ZK_FWD_FORWARD_FAT_POINTER : fwd_byte == ZK_FWD_USE_AUX_HEAP ? ZK_FWD_USE_AUX_HEAP : ZK_FWD_USE_HEAP;
  bool to_label = entry & ZK_E_FLAG0;
  const uint32_t label_pc = imm0();
  const bool local = is_local();
  const uint32_t base = L(L_BASE_PAGE);
  bool deref_beyond = false;
  if (!local) {
    if (fwd == ZK_FWD_FORWARD_FAT_POINTER) {
      if (!src0_ptr) variant = ZK_RET_PANIC;
      if (p_page < base) variant = ZK_RET_PANIC;
    }
    const bool fresh = fwd != ZK_FWD_FORWARD_FAT_POINTER;
    deref_beyond = (uint64_t)p_start + (uint64_t)p_len > 0xFFFFFFFFull;
    if ((fresh && p_off != 0) || deref_beyond) variant = ZK_RET_PANIC;
    if (p_off > p_len) variant = ZK_RET_PANIC;
    if (variant == ZK_RET_PANIC) p_off = p_page = p_start = p_len = 0;
  }
  uint32_t ergs_remaining = ergs;
  if (!local) {
    if (variant != ZK_RET_PANIC) {
      if (fwd == ZK_FWD_FORWARD_FAT_POINTER) {
        p_start += p_off;
        p_len -= p_off;
        p_off = 0;
      } else {
        p_page = base + (fwd == ZK_FWD_USE_HEAP ? 2u : 3u);
      }
    }
    uint32_t growth = 0;
    if (fwd != ZK_FWD_FORWARD_FAT_POINTER) {
      uint32_t upper = p_start + p_len;
      if (deref_beyond) upper = 0xFFFFFFFFu;
      uint32_t bound = fwd == ZK_FWD_USE_HEAP ? L(L_HEAP_BOUND) : L(L_AUX_BOUND);
      if (upper >= bound) growth = upper - bound;
    }
    if (ergs_remaining >= growth) {
      ergs_remaining -= growth;
    } else {
      ergs_remaining = 0;
      variant = ZK_RET_PANIC;
      p_off = p_page = p_start = p_len = 0;
    }
  }
  const bool panicked = variant == ZK_RET_REVERT || variant == ZK_RET_PANIC;
  const uint32_t finished_level = far_depth();
  const uint32_t finished_eh = L(L_EH_BITS) & 0xFFFFu;
  const uint32_t fin_heap_bound = L(L_HEAP_BOUND), fin_aux_bound = L(L_AUX_BOUND);
  if (S.row[L_DEPTH] == 0) {
    fail(ZKB_VM_REFERENCE_PANIC);  // pop of the root frame (execution_stack.rs:113 unwrap)
    return;
  }
  osync();
  pop_frame(panicked);
  to_label = to_label && local;
  if (!local) {
    memory_finish_global_frame(finished_level, base, p_page);
    u256l r1 = lane == 0 ? p_off : lane == 1 ? p_page : lane == 2 ? p_start : lane == 3 ? p_len : 0u;
    S.regs[1][lane] = r1;
    S.row[24 + lane] = r1;
#pragma unroll
    for (int r = 2; r < 16; r++) S.regs[r][lane] = 0u;
    if (lane < 4) S.row[L_CTX + lane] = 0u;
    ptr_mask = 0x0002u;
    rowbits |= ZKB_ROWBIT_DST0_VALID | ZKB_ROWBIT_DST0_PTR;
    osync();
  }
  ergs += ergs_remaining;  // ret.rs:243
  if (to_label) pc = label_pc;
  else if (panicked) pc = finished_eh;
  if (local) {
    if (fin_heap_bound < L(L_HEAP_BOUND) || fin_aux_bound < L(L_AUX_BOUND)) {
      fail(ZKB_VM_REFERENCE_PANIC);  // ret.rs:255-256
      return;
    }
    osync();
    if (lane == 0) {
      S.row[L_HEAP_BOUND] = fin_heap_bound;
      S.row[L_AUX_BOUND] = fin_aux_bound;
    }
    osync();
  }
  if (variant == ZK_RET_PANIC) flags = 1u;
}

// uma.rs:26-425
template <bool KD>
__device__ __forceinline__ void Vm<KD>::op_uma(uint32_t sub, u256l src0, u256l src1, bool src0_ptr) {
  const bool inc = entry & ZK_E_FLAG0;
  uint32_t p_off = oshfl(src0, 0), p_page = oshfl(src0, 1);
  const uint32_t p_start = oshfl(src0, 2), p_len = oshfl(src0, 3);
  const bool is_ptr_read = sub == ZK_UMA_PTR_READ;
  const bool is_heap = sub == ZK_UMA_HEAP_READ || sub == ZK_UMA_HEAP_WRITE;
  const bool is_write = sub == ZK_UMA_HEAP_WRITE || sub == ZK_UMA_AUX_WRITE;
  bool ex = false, ex_deref = false, skip = false;
  if (is_ptr_read && !src0_ptr) ex = true;
  uint32_t mtype;
  if (is_ptr_read) {
    mtype = ZK_MEM_FAT_PTR;
  } else {
    p_page = L(L_BASE_PAGE) + (is_heap ? 2u : 3u);
    mtype = is_heap ? ZK_MEM_HEAP : ZK_MEM_AUX_HEAP;
  }
  uint32_t src_offset;
  if (is_ptr_read) {
    if (!(p_off < p_len)) skip = true;
    src_offset = p_start + p_off;
  } else {
    uint32_t hi_nz = oballot(src0 != 0) & 0xFEu;
    if (hi_nz || p_off > (uint32_t)ZK_MAX_OFFSET_TO_DEREF) {
      ex = ex_deref = true;
      skip = true;
    }
    src_offset = p_off;
  }
  const uint32_t incremented = p_off + 32u;
  if (incremented < p_off) ex = true;
  uint32_t growth = 0;
  if (!is_ptr_read) {
    int w = is_heap ? L_HEAP_BOUND : L_AUX_BOUND;
    uint32_t bound = S.row[w];
    osync();
    if (incremented >= bound) {
      growth = incremented - bound;
      setL(w, incremented);
    }
  }
  if (ex_deref) growth = 0xFFFFFFFFu;
  if (ergs < growth) {
    ergs = 0;
    ex = true;
  } else {
    ergs -= growth;
  }
  const bool set_panic = ex;
  const bool skip_access = skip || set_panic;
  const uint32_t word0 = src_offset >> 5, word1 = word0 + 1, un = src_offset & 31u;
  const bool unaligned = un != 0;
  u256l w0 = 0u, w1 = 0u;
  uint32_t slab = ZKB_NO_SLAB;
  if (!skip_access) {
    if (is_ptr_read) {
      bool ok;
      w0 = fatptr_read(p_page, word0, ok);
      if (!ok) {
        fail(ZKB_VM_REFERENCE_PANIC);
        return;
      }
      emit_mem(timestamp, p_page, word0, mtype, 0, 0, ZKB_MEMORIGIN_VM, w0);
      if (unaligned) {
        w1 = fatptr_read(p_page, word1, ok);
        emit_mem(timestamp, p_page, word1, mtype, 0, 0, ZKB_MEMORIGIN_VM, w1);
      }
    } else {
      slab = cur_slab(is_heap ? 0 : 1, is_write, unaligned ? word1 : word0);
      if (status != ZKB_VM_RUNNING) return;
      w0 = slab_read(slab, word0);
      emit_mem(timestamp, p_page, word0, mtype, 0, 0, ZKB_MEMORIGIN_VM, w0);
      if (unaligned) {
        w1 = slab_read(slab, word1);
        emit_mem(timestamp, p_page, word1, mtype, 0, 0, ZKB_MEMORIGIN_VM, w1);
      }
    }
  }
  if (!is_write) {
    u256l r = w0;  // aligned accesses (the common case) need no byte shuffling
    if (unaligned) r = u_shl(w0, un * 8u, lane) | u_shr(w1, (32u - un) * 8u, lane);
    if (is_ptr_read) {
      uint32_t beyond = incremented - p_len;
      if (incremented < p_len || skip_access) beyond = 0;
      beyond &= 31u;
      if (beyond) r = u_shl(u_shr(r, beyond * 8u, lane), beyond * 8u, lane);
    }
    if (!set_panic) {
      dst0_update(r, false);
      if (inc) dst1_update(lane == 0 ? incremented : src0, src0_ptr);
    } else {
      pending = 1;
    }
  } else {
    const uint32_t low0 = 32u - un;
    u256l n0 = src1, n1 = 0u;
    if (unaligned) {
      n0 = u_shl(u_shr(w0, low0 * 8u, lane), low0 * 8u, lane) | u_shr(src1, un * 8u, lane);
      n1 = u_shr(u_shl(w1, un * 8u, lane), un * 8u, lane) | u_shl(src1, (32u - un) * 8u, lane);
    }
    if (!skip_access) {
      slab_write(slab, word0, n0);
      emit_mem(timestamp + 3, p_page, word0, mtype, 1, 0, ZKB_MEMORIGIN_VM, n0);
      if (unaligned) {
        slab_write(slab, word1, n1);
        emit_mem(timestamp + 3, p_page, word1, mtype, 1, 0, ZKB_MEMORIGIN_VM, n1);
      }
    }
    if (!set_panic) {
      if (inc) dst0_update(lane == 0 ? incremented : src0, false);
    } else {
      pending = 1;
    }
  }
}

// ===================================================================================================
// load / run / store one VM
// ===================================================================================================
// load VM `v.vm`'s hot state from HBM into shared memory / octet-uniform registers
template <bool KD>
__device__ __forceinline__ void vm_load(Vm<KD>& v, const VmHot* hot) {
  VmSmem& S = v.S;
  const uint32_t lane = v.lane;
  uint4* sregs = reinterpret_cast<uint4*>(&S.regs[0][0]);
  const uint4* hregs = reinterpret_cast<const uint4*>(&hot->regs[0][0]);
#pragma unroll
  for (int i = 0; i < 4; i++) sregs[i * 8 + lane] = hregs[i * 8 + lane];
  reinterpret_cast<uint4*>(S.F)[lane] = reinterpret_cast<const uint4*>(hot->F)[lane];
  // row words 0..39 start as zero, words 40..55 are the live tail, 56..63 reserved (zero)
  uint4* srow = reinterpret_cast<uint4*>(S.row);
  const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
  srow[lane] = zero4;
  srow[8 + lane] = (lane >= 2 && lane < 6) ? reinterpret_cast<const uint4*>(hot->live)[lane - 2] : zero4;
  if (lane == 0) {
    S.kbuf[KB_EC_PENDING] = 0u;
    S.kc[KB_KC_PENDING] = 0u;
  }
  S.pw[lane] = hot->prev_word[lane];
  v.tx_psp = hot->live[L_TX_PSP - 40];
  const uint32_t* x = hot->x;
  v.timestamp = x[X_TIMESTAMP];
  v.cycle = x[X_CYCLE];
  v.flags = x[X_FLAGS];
  v.pending = x[X_PENDING];
  v.status = ZKB_VM_RUNNING;
  v.ptr_mask = x[X_PTRMASK];
  v.prev_code_page = x[X_PREV_CODE_PAGE];
  v.journal_len() = x[X_JOURNAL_LEN];
  v.n_decommit() = x[X_N_DECOMMIT];
  v.slab_free() = x[X_SLAB_FREE];
  v.n_rows = x[X_COUNT0 + ZKB_STREAM_ROWS];
  v.n_mem = x[X_COUNT0 + ZKB_STREAM_MEM];
#pragma unroll
  for (int k = ZKB_STREAM_LOG; k < ZKB_N_STREAMS; k++) v.count(k) = x[X_COUNT0 + k];
  v.row_ptr = v.B.streams[ZKB_STREAM_ROWS] + ((size_t)v.vm * v.B.cap[ZKB_STREAM_ROWS] + v.n_rows) * ZKB_ROW_BYTES;
  v.mem_ptr = v.B.streams[ZKB_STREAM_MEM] + ((size_t)v.vm * v.B.cap[ZKB_STREAM_MEM] + v.n_mem) * ZKB_MEM_BYTES;
  v.rowbits = 0;
  v.ccount = 0;
  v.entry = v.ops_lo = v.ops_hi = v.dst_loc = 0;
  v.defer_kind = DEFER_NONE;
  v.set_vm_pointers();
  osync();
  v.load_frame_from_F();
  // cold per-VM tables that the interpreter keeps in shared memory while the VM runs
  {
    const uint4* gpt = reinterpret_cast<const uint4*>(v.B.pt + (size_t)v.vm * ZKB_PT_ENTRIES * 2);
    uint4* spt = reinterpret_cast<uint4*>(S.pt);
    spt[lane] = gpt[lane];
    spt[8 + lane] = gpt[8 + lane];
    const uint32_t* ghwm = v.B.slab_hwm + (size_t)v.vm * v.B.n_slabs;
    for (uint32_t i = lane; i < v.B.n_slabs; i += ZK_OCT) S.hwm[i] = ghwm[i];
    if (lane < 4) S.lv[lane] = v.g_lvl()[v.far_depth() * 4 + lane];
  }
  osync();
}

// write the hot state back to HBM (the batch is resumable: zkb_run may be called again)
template <bool KD>
__device__ __forceinline__ void vm_store(Vm<KD>& v, VmHot* hot) {
  VmSmem& S = v.S;
  const uint32_t lane = v.lane;
  osync();
  v.sync_frame_to_F();
  const uint4* sregs = reinterpret_cast<const uint4*>(&S.regs[0][0]);
  uint4* gregs = reinterpret_cast<uint4*>(&hot->regs[0][0]);
#pragma unroll
  for (int i = 0; i < 4; i++) gregs[i * 8 + lane] = sregs[i * 8 + lane];
  reinterpret_cast<uint4*>(hot->F)[lane] = reinterpret_cast<const uint4*>(S.F)[lane];
  if (lane < 4) reinterpret_cast<uint4*>(hot->live)[lane] = reinterpret_cast<const uint4*>(S.row)[10 + lane];
  hot->prev_word[lane] = S.pw[lane];
  // x[]: octet lane l writes words 4l .. 4l+3
  {
    uint32_t w[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const uint32_t i = lane * 4 + q;
      uint32_t out = 0;
      out = i == X_TIMESTAMP ? v.timestamp : out;
      out = i == X_CYCLE ? v.cycle : out;
      out = i == X_FLAGS ? v.flags : out;
      out = i == X_PENDING ? v.pending : out;
      out = i == X_STATUS ? (v.status == ZKB_VM_PARKED ? (uint32_t)ZKB_VM_RUNNING : v.status) : out;
      out = i == X_DEFER ? v.defer_kind : out;
      out = i == X_PTRMASK ? v.ptr_mask : out;
      out = i == X_PREV_CODE_PAGE ? v.prev_code_page : out;
      out = i == X_FAR_DEPTH ? v.far_depth() : out;
      out = i == X_JOURNAL_LEN ? v.journal_len() : out;
      out = i == X_N_DECOMMIT ? v.n_decommit() : out;
      out = i == X_SLAB_FREE ? v.slab_free() : out;
#pragma unroll
      for (int k = 0; k < ZKB_N_STREAMS; k++) out = i == (uint32_t)(X_COUNT0 + k) ? v.stream_count(k) : out;
      w[q] = out;
    }
    reinterpret_cast<uint4*>(hot->x)[lane] = make_uint4(w[0], w[1], w[2], w[3]);
  }
  // the per-VM summary goes straight to mapped host memory (one 32-byte posted write per VM and run): the host can
  // size and enqueue the witness download without a D2H copy of its own queueing behind the copies already in flight
  const uint32_t st_now = v.status == ZKB_VM_PARKED ? (uint32_t)ZKB_VM_RUNNING : v.status;
  const uint32_t st_out = (st_now == ZKB_VM_RUNNING && S.row[L_DEPTH] == 0 && v.cycle > 0) ? (uint32_t)ZKB_VM_ENDED : st_now;
  uint32_t summary = lane == 6 ? st_out : v.cycle;
#pragma unroll
  for (int k = 0; k < ZKB_N_STREAMS; k++) summary = lane == (uint32_t)k ? v.stream_count(k) : summary;
  v.B.host_counts[(size_t)v.vm * 8 + lane] = summary;
  {
    uint4* gpt = reinterpret_cast<uint4*>(v.B.pt + (size_t)v.vm * ZKB_PT_ENTRIES * 2);
    const uint4* spt = reinterpret_cast<const uint4*>(S.pt);
    gpt[lane] = spt[lane];
    gpt[8 + lane] = spt[8 + lane];
    uint32_t* ghwm = v.B.slab_hwm + (size_t)v.vm * v.B.n_slabs;
    for (uint32_t i = lane; i < v.B.n_slabs; i += ZK_OCT) ghwm[i] = S.hwm[i];
    if (lane < 4) v.g_lvl()[v.far_depth() * 4 + lane] = S.lv[lane];
  }
  osync();
}

// the deferred half of the ecrecover precompile: runs between vm_store and vm_load, i.e. with no interpreter state in
// registers; reads its inputs from the VM's scratch, patches the heap words and the two memory-query records
// (plain scalar arguments: handing the launch-constant DevBatch to a non-inlined function by reference would force a
// copy of the whole struct onto the local-memory stack of every thread)
__device__ __noinline__ void run_deferred_ecrecover(uint32_t* kbuf, uint64_t* ks, uint32_t* vm_heap, uint32_t heap_words, uint8_t* vm_mem_stream,
                                                    uint32_t lane) {
  osync();
  u256l in[4];
#pragma unroll
  for (int i = 0; i < 4; i++) in[i] = kbuf[KB_EC_INPUT + 8 * i + lane];
  const uint32_t rec = kbuf[KB_EC_MEM_INDEX], slab = kbuf[KB_EC_SLAB], out_word = kbuf[KB_EC_OUT_WORD];
  osync();
  u256l address;
  const bool ok = secp::ecrecover_octet(in[0], in[2], in[3], oshfl(in[1], 0), lane, ks, address);
  const u256l marker = lane == 0 ? (ok ? 1u : 0u) : 0u;
  uint32_t* heap = vm_heap + ((size_t)slab * heap_words + out_word) * 8;
  heap[lane] = marker;
  heap[8 + lane] = address;
  if (vm_mem_stream) {
    uint32_t* r = reinterpret_cast<uint32_t*>(vm_mem_stream + (size_t)rec * ZKB_MEM_BYTES);
    r[4 + lane] = marker;
    r[ZKB_MEM_BYTES / 4 + 4 + lane] = address;
  }
  if (lane == 0) kbuf[KB_EC_PENDING] = 0u;
  osync();
}
__device__ __forceinline__ void deferred_ecrecover(const DevBatch& B, VmSmem& S, uint32_t vm, uint32_t lane) {
  run_deferred_ecrecover(S.kbuf, S.ks(), B.heap_mem + (size_t)vm * B.n_slabs * B.heap_words * 8, B.heap_words,
                         B.witness ? B.streams[ZKB_STREAM_MEM] + (size_t)vm * B.cap[ZKB_STREAM_MEM] * ZKB_MEM_BYTES : nullptr, lane);
}

// The deferred half of the keccak256 precompile for the VM in shared-memory slot `S` (one THREAD per VM; the caller
// has synchronised with the octet that left the descriptor).  Reads the message from the VM's heap slab, patches the
// digest into the reserved heap word and memory-query record.  Scalar arguments only (see run_deferred_ecrecover).
__device__ __noinline__ void run_deferred_keccak(uint32_t* kbuf, uint32_t* heap_all, uint32_t n_slabs, uint32_t heap_words, uint8_t* mem_stream_all,
                                                 uint32_t mem_cap) {
  if (!kbuf[KB_KC_PENDING]) return;
  const uint32_t vm = kbuf[KB_KC_VM], src_slab = kbuf[KB_KC_SRC_SLAB];
  uint32_t* vm_heap = heap_all + (size_t)vm * n_slabs * heap_words * 8;
  const uint8_t* src = src_slab == ZKB_NO_SLAB ? nullptr : reinterpret_cast<const uint8_t*>(vm_heap + (size_t)src_slab * heap_words * 8);
  uint64_t digest[4];
  keccak256_slab(src, heap_words, kbuf[KB_KC_IN_OFF], kbuf[KB_KC_IN_LEN], digest);
  // digest = 32 stream bytes = one big-endian word: stored byte-reversed (little-endian limbs)
  uint64_t out[4];
#pragma unroll
  for (int i = 0; i < 4; i++) out[i] = kc_bswap64(digest[3 - i]);
  uint64_t* hw = reinterpret_cast<uint64_t*>(vm_heap + ((size_t)kbuf[KB_KC_OUT_SLAB] * heap_words + kbuf[KB_KC_OUT_WORD]) * 8);
#pragma unroll
  for (int i = 0; i < 4; i++) hw[i] = out[i];
  if (mem_stream_all) {
    uint64_t* r = reinterpret_cast<uint64_t*>(mem_stream_all + ((size_t)vm * mem_cap + kbuf[KB_KC_MEM_INDEX]) * ZKB_MEM_BYTES + 16);
#pragma unroll
    for (int i = 0; i < 4; i++) r[i] = out[i];
  }
  kbuf[KB_KC_PENDING] = 0u;
}

// Runs the four VMs of one warp (VM `vm_idx` on the calling octet) to the end, or for max_cycles cycles each.
//
// Inside the warp the four octets re-converge at one warp vote per VM cycle, so VMs that follow the same path through
// the interpreter share every issue slot.  LOCKSTEP additionally lines up the W warps of the CTA: they meet at a CTA
// barrier every ZKB_LOCKSTEP_PERIOD cycles, so batches of transactions against the same contracts execute the same
// handler at the same time and share its instruction-cache lines (the interpreter is ~170 KB of SASS against a 32 KB
// L1.5 instruction cache; free-running warps spend most of their issue slots waiting for instruction fetch).
// All threads of the warp (LOCKSTEP: of the CTA) must call, also for vm_idx >= n_vms.
// kc_flags: three CTA-shared words, a ring indexed by the lockstep period: octets that yield a keccak raise
// kc_flags[period % 3] before the period's CTA barrier, everybody reads it after the barrier, thread 0 clears the word
// of period + 2 (last read before this barrier, next written after the following one): no race, no second barrier.
// Returns 0 when every VM of the group has ended (or used up max_cycles), 1 when VMs of the group wait for their deferred
// keccak256: the caller runs the sponges (deferred_keccak_phase) and calls again -- the VM state is parked in HBM across
// that phase (this function loads at entry and stores at exit anyway), so no interpreter register is live across the
// call of the ~5 000-instruction sponge.  n = cycles run so far in this launch (carried across re-entries).
template <bool LOCKSTEP, bool KD>
__device__ __forceinline__ uint32_t run_vm_group(const DevBatch& B, VmSmem& S, uint32_t* kc_flags, uint32_t vm_idx, uint32_t lane, uint32_t max_cycles,
                                                 uint32_t& n) {
  // the FAST kernel runs the VMs that are not parked, the FULL kernel exactly the parked ones (a VM that merely ran out
  // of max_cycles in the fast kernel is not touched again by this zkb_run)
  // vm_idx arrives as a SCHEDULE SLOT: the host may have regrouped the VMs so that the octets of a warp (and the warps of
  // a lockstep CTA) run the same bootloader code; every per-VM array and stream stays indexed by the real VM number, so
  // the emitted bytes do not depend on the grouping
  if (B.order != nullptr && vm_idx < B.n_vms) vm_idx = __ldg(B.order + vm_idx);
  const uint32_t* x0 = B.hot[vm_idx < B.n_vms ? vm_idx : 0].x;
  const bool valid = vm_idx < B.n_vms && x0[X_STATUS] == ZKB_VM_RUNNING && ((x0[X_DEFER] != DEFER_NONE) == KD);
  VmHot* hot = B.hot + (valid ? vm_idx : 0);
  Vm<KD> v(B, S, valid ? vm_idx : 0, lane);
  v.status = ZKB_VM_ENDED;
  if (valid) {
    const uint32_t kind = KD ? x0[X_DEFER] : (uint32_t)DEFER_NONE;
    vm_load(v, hot);
    if (S.row[L_DEPTH] == 0) v.status = ZKB_VM_ENDED;  // nothing to run; later ends are detected by the RET that pops the last frame
    if (KD) {  // pick the parked VM up where the fast kernel left it: descriptor back into the scratch, result still pending
      const uint32_t* df = B.defer + (size_t)vm_idx * ZKB_DEFER_WORDS;
      n = df[DF_CYCLES_RUN];
      if (kind == DEFER_ECRECOVER) {
#pragma unroll
        for (int i = 0; i < 4; i++) S.kbuf[KB_EC_INPUT + 8 * i + lane] = df[DF_EC_INPUT + 8 * i + lane];
        if (lane < 4) S.kbuf[KB_EC_PENDING + lane] = df[DF_EC_DESC + lane];
        v.status = ZKB_VM_YIELD_ECRECOVER;
      } else if (kind == DEFER_KECCAK) {
        S.kc[lane] = df[DF_KC + lane];
        v.status = ZKB_VM_YIELD_KECCAK;
      }
      osync();
    }
  }
  uint32_t period = 0, reason = 0;
  while (true) {
    // up to ZKB_LOCKSTEP_PERIOD cycles between two CTA barriers: the warps may drift by a few hundred instructions
    // (still inside the I-cache window) and the barrier waits for the slowest SUM of cycles, not the slowest cycle
    bool active = valid && v.status == ZKB_VM_RUNNING && !(max_cycles && n >= max_cycles);
#pragma unroll 1
    for (int k = 0; k < ZKB_LOCKSTEP_PERIOD; k++) {
      if (!__any_sync(ZK_FULL, active)) break;  // warp-uniform; also the per-cycle re-convergence point of the four octets
      if (active) {
        v.cycle_once();
        n++;
        active = v.status == ZKB_VM_RUNNING && !(max_cycles && n >= max_cycles);
      }
    }
    if (!KD) {
      // FAST kernel: a pending precompile result parks the VM for the FULL kernel (descriptor -> DevBatch.defer)
      if (valid && (v.status == ZKB_VM_YIELD_ECRECOVER || v.status == ZKB_VM_YIELD_KECCAK)) {
        uint32_t* df = B.defer + (size_t)v.vm * ZKB_DEFER_WORDS;
        osync();
        if (v.status == ZKB_VM_YIELD_ECRECOVER) {
#pragma unroll
          for (int i = 0; i < 4; i++) df[DF_EC_INPUT + 8 * i + lane] = S.kbuf[KB_EC_INPUT + 8 * i + lane];
          if (lane < 4) df[DF_EC_DESC + lane] = S.kbuf[KB_EC_PENDING + lane];
          v.defer_kind = DEFER_ECRECOVER;
        } else {
          df[DF_KC + lane] = S.kc[lane];
          v.defer_kind = DEFER_KECCAK;
        }
        if (lane == 0) df[DF_CYCLES_RUN] = n;
        v.status = ZKB_VM_PARKED;
      }
    } else if (valid && v.status == ZKB_VM_YIELD_ECRECOVER) {  // park the VM, finish the pending recovery, resume
      v.status = ZKB_VM_RUNNING;
      vm_store(v, hot);
      deferred_ecrecover(B, S, v.vm, lane);
      vm_load(v, hot);
      active = !(max_cycles && n >= max_cycles);
    }
    const bool kc_yield = KD && valid && v.status == ZKB_VM_YIELD_KECCAK;
    if (LOCKSTEP) {
      if (kc_yield) kc_flags[period] = 1u;   // (several octets may write the same 1)
      const int any = __syncthreads_or(active ? 1 : 0);
      if (KD && kc_flags[period]) {   // (only the FULL kernel ever raises it: the FAST one provably returns 0)
        reason = 1;
        break;
      }
      if (threadIdx.x == 0) kc_flags[(period + 2u) % 3u] = 0u;
      period = (period + 1u) % 3u;
      if (!any) break;
    } else {
      if (KD && __any_sync(ZK_FULL, kc_yield)) {
        reason = 1;
        break;
      }
      if (!__any_sync(ZK_FULL, active)) break;
    }
  }
  if (valid) {
    if (v.status == ZKB_VM_YIELD_KECCAK) v.status = ZKB_VM_RUNNING;  // it continues after the caller's deferred phase
    if (KD && reason) {  // the FULL kernel comes back for this VM right after the deferred phase
      v.defer_kind = DEFER_CONTINUE;
      if (lane == 0) B.defer[(size_t)v.vm * ZKB_DEFER_WORDS + DF_CYCLES_RUN] = n;
    }
    vm_store(v, hot);
  }
  return reason;
}

}  // namespace zkb
