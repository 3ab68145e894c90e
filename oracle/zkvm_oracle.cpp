// =====================================================================================================
// TEST INFRASTRUCTURE — CPU oracle for the batched EraVM witness generator.  NOT part of the product:
// only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
//
// What it is: a scalar C++17 restatement of the reference hot path of matter-labs/era-zk_evm (crate zk_evm 1.4.1)
//   VmState::cycle()                 /root/reference/src/vm_state/cycle.rs:19-429
//   operand addressing               src/vm_state/mem_ops.rs:14-125
//   query emission helpers           src/vm_state/helpers.rs:10-338
//   the 15 opcode handlers           src/opcodes/execution/*.rs
//   SimpleMemory / SimpleDecommitter / InMemoryEventSink / InMemoryStorage
//                                    src/reference_impls/{memory,decommitter,event_sink}.rs, src/testing/storage.rs
// written function-for-function after those files (each function cites the lines it follows), recording the
// VmWitnessTracer callbacks (src/witness_trace/mod.rs:11-72) as the canonical packed records of
// include/zkb_records.h.
//
// PARITY STATUS: the reference cannot be built here (no rustc/cargo; its two git dependencies
// zkevm_opcode_defs / zk_evm_abstractions @ branch v1.4.1 are absent).  The interpreter logic is pinned by
// source; the ISA data table (era_zk_evm_b200/isa.py -> isa_tables.inc), ABI bit layouts and the
// sha256/ecrecover precompile memory ABIs are reconstructions => "parity unpinned" for those parts.
// keccak256 is pinned by the reference's live tests, sha256 / ecrecover by its hash and known-answer vectors, the ALU by
// Python-int vectors, handler quirks by hand-derived programs (tests/test_oracle_golden.py, tests/test_semantics.py).
//
// It is deliberately KINDER to the CPU than the reference: stack/heap pages are grown lazily and cleared by
// high-water mark instead of the reference's 65 536-entry fill per far return (memory.rs:185-192), and code
// pages alias the shared bytecode instead of being copied word by word (decommitter.rs:81-97).
// =====================================================================================================
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../include/zkb.h"
#include "../era_zk_evm_b200/csrc/isa_tables.inc"
#include "../include/zkb_codec.h"
#include "hashes.hpp"
#include "secp256k1.hpp"
#include "u256.hpp"

namespace {

struct RefPanic : std::runtime_error {
  using std::runtime_error::runtime_error;
};
struct UnknownCodeHash : std::runtime_error {
  using std::runtime_error::runtime_error;
};
struct Unsupported : std::runtime_error {
  using std::runtime_error::runtime_error;
};
#define REF_ASSERT(cond, msg) \
  do {                        \
    if (!(cond)) throw RefPanic(msg); \
  } while (0)

struct Address {
  uint8_t b[20];
  bool operator==(const Address& o) const { return memcmp(b, o.b, 20) == 0; }
  static Address zero() {
    Address a;
    memset(a.b, 0, 20);
    return a;
  }
};

// utils.rs:29-34
static U256 address_to_u256(const Address& a) {
  uint8_t buf[32];
  memset(buf, 0, 32);
  memcpy(buf + 12, a.b, 20);
  return U256::from_be(buf);
}
// utils.rs:36-41
static Address u256_to_address_unchecked(const U256& v) {
  uint8_t buf[32];
  v.to_be(buf);
  Address a;
  memcpy(a.b, buf + 12, 20);
  return a;
}

// vm_state/mod.rs:32-35
struct PrimitiveValue {
  U256 value;
  bool is_pointer;
  static PrimitiveValue empty() { return PrimitiveValue{U256::zero(), false}; }
};

// flags.rs:4-8
struct Flags {
  bool lt = false, eq = false, gt = false;
  void reset() { lt = eq = gt = false; }
  uint8_t bits() const { return (uint8_t)(lt | (eq << 1) | (gt << 2)); }
};

// execution_stack.rs:6-24
struct CallStackEntry {
  Address this_address, msg_sender, code_address;
  uint32_t base_memory_page, code_page;
  uint16_t sp, pc, exception_handler_location;
  uint32_t ergs_remaining;
  uint8_t this_shard_id, caller_shard_id, code_shard_id;
  bool is_static, is_local_frame;
  uint32_t context_u128_value[4];
  uint32_t heap_bound, aux_heap_bound;

  // execution_stack.rs:35-55
  static CallStackEntry empty_context() {
    CallStackEntry e;
    memset(&e, 0, sizeof(e));
    e.base_memory_page = ZK_UNMAPPED_PAGE;
    e.code_page = ZK_UNMAPPED_PAGE;
    e.sp = ZK_INITIAL_SP_ON_FAR_CALL;
    e.ergs_remaining = ZK_VM_INITIAL_FRAME_ERGS;
    return e;
  }
  // execution_stack.rs:83-87
  static bool address_is_kernel(const Address& a) {
    for (int i = 0; i < 18; i++)
      if (a.b[i]) return false;
    return true;
  }
  bool is_kernel_mode() const { return address_is_kernel(this_address); }
};
static inline uint32_t stack_page_from_base(uint32_t b) { return b + 1; }     // execution_stack.rs:71
static inline uint32_t heap_page_from_base(uint32_t b) { return b + 2; }      // :75
static inline uint32_t aux_heap_page_from_base(uint32_t b) { return b + 3; }  // :79

// zk_evm_abstractions queries (layouts external; field lists cited by use in helpers.rs:20-26,143-152,171-177)
struct MemoryQuery {
  uint32_t timestamp;
  uint8_t memory_type;
  uint32_t page, index;
  U256 value;
  bool value_is_pointer, rw_flag;
};
struct LogQuery {
  uint32_t timestamp;
  uint16_t tx_number_in_block;
  uint8_t aux_byte, shard_id;
  Address address;
  U256 key, read_value, written_value;
  bool rw_flag, rollback, is_service;
};
struct DecommittmentQuery {
  U256 hash;
  uint32_t timestamp, memory_page;
  uint16_t decommitted_length;
  bool is_fresh;
};

// FatPointer (zkevm_opcode_defs; low-128 layout cited at ptr.rs:79-82, uma.rs:337)
struct FatPointer {
  uint32_t offset, memory_page, start, length;
  static FatPointer empty() { return FatPointer{0, 0, 0, 0}; }
  static FatPointer from_u256(const U256& v) {
    return FatPointer{(uint32_t)v.w[0], (uint32_t)(v.w[0] >> 32), (uint32_t)v.w[1], (uint32_t)(v.w[1] >> 32)};
  }
  U256 to_u256() const {
    return U256{{(uint64_t)offset | ((uint64_t)memory_page << 32), (uint64_t)start | ((uint64_t)length << 32), 0, 0}};
  }
  // FatPointerValidationException bits
  enum { OFFSET_IS_NOT_ZERO_WHEN_EXPECTED = 1, DEREF_BEYOND_HEAP_RANGE = 2 };
  uint32_t validate(bool is_fresh) const {
    uint32_t ex = 0;
    if (is_fresh && offset != 0) ex |= OFFSET_IS_NOT_ZERO_WHEN_EXPECTED;
    if ((uint64_t)start + (uint64_t)length > 0xFFFFFFFFull) ex |= DEREF_BEYOND_HEAP_RANGE;
    return ex;
  }
  bool validate_in_bounds() const { return offset < length; }
  bool validate_as_slice() const { return offset <= length; }
};

// ---------------------------------------------------------------------------------------------------
// recording tracer: serialises VmWitnessTracer callbacks (witness_trace/mod.rs:11-72) into zkb records
// ---------------------------------------------------------------------------------------------------
struct Recorder {
  bool enabled = true;
  std::vector<uint8_t> s[ZKB_N_STREAMS];
  uint64_t counts[ZKB_N_STREAMS] = {0, 0, 0, 0, 0, 0};
  template <class T>
  void push(int kind, const T& rec) {
    counts[kind]++;
    if (!enabled) return;
    const uint8_t* p = (const uint8_t*)&rec;
    s[kind].insert(s[kind].end(), p, p + sizeof(T));
  }
  void add_memory_query(const MemoryQuery& q, uint8_t origin) {
    ZkbMemoryQueryRec r;
    memset(&r, 0, sizeof(r));
    r.timestamp = q.timestamp;
    r.page = q.page;
    r.index = q.index;
    r.memory_type = q.memory_type;
    r.rw_flag = q.rw_flag;
    r.value_is_pointer = q.value_is_pointer;
    r.origin = origin;
    q.value.to_limbs32(r.value);
    push(ZKB_STREAM_MEM, r);
  }
  void add_log_query(const LogQuery& q) {
    ZkbLogQueryRec r;
    memset(&r, 0, sizeof(r));
    r.timestamp = q.timestamp;
    r.tx_number_in_block = q.tx_number_in_block;
    r.aux_byte = q.aux_byte;
    r.shard_id = q.shard_id;
    memcpy(r.address, q.address.b, 20);
    r.rw_flag = q.rw_flag;
    r.rollback = q.rollback;
    r.is_service = q.is_service;
    q.key.to_limbs32(r.key);
    q.read_value.to_limbs32(r.read_value);
    q.written_value.to_limbs32(r.written_value);
    push(ZKB_STREAM_LOG, r);
  }
  void add_decommittment(const DecommittmentQuery& q) {
    ZkbDecommitRec r;
    memset(&r, 0, sizeof(r));
    r.timestamp = q.timestamp;
    r.memory_page = q.memory_page;
    r.decommitted_length = q.decommitted_length;
    r.is_fresh = q.is_fresh;
    q.hash.to_limbs32(r.hash);
    push(ZKB_STREAM_DECOMMIT, r);
  }
  void record_refund(uint32_t type, uint32_t value) {
    ZkbRefundRec r{type, value};
    push(ZKB_STREAM_REFUND, r);
  }
  // set by the far_call handler right before start_frame: which bound of `prev` the call's memory growth touched
  uint16_t far_call_bound_kind = 0;
  void start_new_execution_context(uint32_t cycle, const CallStackEntry& prev, const CallStackEntry& e) {
    ZkbFrameRec r;
    memset(&r, 0, sizeof(r));
    r.kind = ZKB_FRAMEKIND_START;
    r.prev_bound_kind = far_call_bound_kind;
    r.prev_bound_value = far_call_bound_kind == 1 ? prev.heap_bound : far_call_bound_kind == 2 ? prev.aux_heap_bound : 0;
    far_call_bound_kind = 0;
    r.cycle = cycle;
    memcpy(r.this_address, e.this_address.b, 20);
    memcpy(r.msg_sender, e.msg_sender.b, 20);
    memcpy(r.code_address, e.code_address.b, 20);
    r.base_memory_page = e.base_memory_page;
    r.code_page = e.code_page;
    r.sp = e.sp;
    r.pc = e.pc;
    r.exception_handler_location = e.exception_handler_location;
    r.this_shard_id = e.this_shard_id;
    r.caller_shard_id = e.caller_shard_id;
    r.ergs_remaining = e.ergs_remaining;
    r.code_shard_id = e.code_shard_id;
    r.is_static = e.is_static;
    r.is_local_frame = e.is_local_frame;
    memcpy(r.context_u128_value, e.context_u128_value, 16);
    r.heap_bound = e.heap_bound;
    r.aux_heap_bound = e.aux_heap_bound;
    r.prev_ergs_remaining = prev.ergs_remaining;
    r.prev_pc = prev.pc;
    r.prev_sp = prev.sp;
    push(ZKB_STREAM_FRAME, r);
  }
  void finish_execution_context(uint32_t cycle, bool panicked) {
    ZkbFrameRec r;
    memset(&r, 0, sizeof(r));
    r.kind = ZKB_FRAMEKIND_FINISH;
    r.panicked = panicked;
    r.cycle = cycle;
    push(ZKB_STREAM_FRAME, r);
  }
};

// ---------------------------------------------------------------------------------------------------
// SimpleMemory (src/reference_impls/memory.rs)
// ---------------------------------------------------------------------------------------------------
typedef std::shared_ptr<const std::vector<U256>> CodeRef;

struct Indirection {
  enum Kind { HEAP, AUX_HEAP, RETURNDATA_EXTENDED_LIFETIME, EMPTY } kind;
  size_t index;
};

struct SimpleMemory {
  struct StackPage {
    uint32_t page;
    std::vector<PrimitiveValue> content;  // lazily grown; reference preallocates 2^16 entries (memory.rs:176)
  };
  struct HeapPage {
    uint32_t page;
    std::vector<U256> content;
  };
  std::vector<StackPage> stack_pages;
  std::vector<std::pair<HeapPage, HeapPage>> heaps;
  std::unordered_map<uint32_t, CodeRef> code_pages;
  std::unordered_map<uint32_t, std::vector<U256>> pages_with_extended_lifetime;
  std::unordered_map<uint32_t, Indirection> page_numbers_indirections;
  std::vector<std::unordered_set<uint32_t>> indirections_to_cleanup_on_return;

  // memory.rs:209-238 SimpleMemory::new
  SimpleMemory() {
    code_pages[0] = std::make_shared<const std::vector<U256>>();  // page 0: all zeroes (memory.rs:223-225)
    pages_with_extended_lifetime[ZK_BOOTLOADER_CALLDATA_PAGE] = {};
    page_numbers_indirections[0] = Indirection{Indirection::EMPTY, 0};
    indirections_to_cleanup_on_return.emplace_back();
    heaps.push_back({HeapPage{0, {}}, HeapPage{0, {}}});
  }

  static U256 get_or_zero(const std::vector<U256>& v, uint32_t idx) { return idx < v.size() ? v[idx] : U256::zero(); }

  // memory.rs:404-528 execute_partial_query
  MemoryQuery execute_partial_query(MemoryQuery q) {
    switch (q.memory_type) {
      case ZK_MEM_STACK: {
        REF_ASSERT(!stack_pages.empty(), "no stack page");
        StackPage& sp = stack_pages.back();
        REF_ASSERT(sp.page == q.page, "stack page mismatch (memory.rs:416,428)");
        REF_ASSERT(q.index < ZK_MAX_STACK_PAGE_SIZE_IN_WORDS, "out of bounds for stack page");
        if (q.rw_flag) {
          if (sp.content.size() <= q.index) sp.content.resize(q.index + 1, PrimitiveValue::empty());
          sp.content[q.index] = PrimitiveValue{q.value, q.value_is_pointer};
        } else {
          PrimitiveValue pv = q.index < sp.content.size() ? sp.content[q.index] : PrimitiveValue::empty();
          q.value = pv.value;
          q.value_is_pointer = pv.is_pointer;
        }
        break;
      }
      case ZK_MEM_HEAP:
      case ZK_MEM_AUX_HEAP: {
        REF_ASSERT(!q.value_is_pointer, "heap value is pointer (memory.rs:440)");
        auto& cur = heaps.back();
        HeapPage& hp = q.memory_type == ZK_MEM_HEAP ? cur.first : cur.second;
        REF_ASSERT(hp.page == q.page, "heap page mismatch (memory.rs:447,451,462,466 debug_assert)");
        if (q.rw_flag) {
          if (hp.content.size() <= q.index) hp.content.resize((size_t)q.index + 1, U256::zero());
          hp.content[q.index] = q.value;
        } else {
          q.value = get_or_zero(hp.content, q.index);  // resize_to_fit + read == read-or-zero
        }
        break;
      }
      case ZK_MEM_FAT_PTR: {
        REF_ASSERT(!q.rw_flag && !q.value_is_pointer, "fat pointer write (memory.rs:476-477)");
        auto it = page_numbers_indirections.find(q.page);
        REF_ASSERT(it != page_numbers_indirections.end(), "fat pointer only points to reachable memory (memory.rs:481)");
        const Indirection& ind = it->second;
        switch (ind.kind) {
          case Indirection::HEAP:
            REF_ASSERT(heaps[ind.index].first.page == q.page, "indirection page mismatch (memory.rs:489)");
            q.value = get_or_zero(heaps[ind.index].first.content, q.index);
            break;
          case Indirection::AUX_HEAP:
            REF_ASSERT(heaps[ind.index].second.page == q.page, "indirection page mismatch (memory.rs:499)");
            q.value = get_or_zero(heaps[ind.index].second.content, q.index);
            break;
          case Indirection::RETURNDATA_EXTENDED_LIFETIME: {
            auto pit = pages_with_extended_lifetime.find(q.page);
            REF_ASSERT(pit != pages_with_extended_lifetime.end(), "indirection target must exist (memory.rs:510)");
            q.value = get_or_zero(pit->second, q.index);
            break;
          }
          case Indirection::EMPTY:
            q.value = U256::zero();
            break;
        }
        break;
      }
      default:
        throw RefPanic("code should be through specialized query (memory.rs:523)");
    }
    return q;
  }

  // memory.rs:556-569 read_code_query (pages are 2^16 zero-padded words, memory.rs:279)
  MemoryQuery read_code_query(MemoryQuery q) const {
    REF_ASSERT(q.memory_type == ZK_MEM_CODE && !q.rw_flag, "bad code query");
    auto it = code_pages.find(q.page);
    REF_ASSERT(it != code_pages.end(), "code page must exist (memory.rs:565 unwrap)");
    REF_ASSERT(q.index < ZK_MAX_CODE_PAGE_SIZE_IN_WORDS, "code index");
    q.value = get_or_zero(*it->second, q.index);
    return q;
  }

  // memory.rs:573-657 start_global_frame
  void start_global_frame(uint32_t /*current_base_page*/, uint32_t new_base_page, FatPointer calldata) {
    stack_pages.push_back(StackPage{stack_page_from_base(new_base_page), {}});
    uint32_t heap_page = heap_page_from_base(new_base_page), aux_page = aux_heap_page_from_base(new_base_page);
    uint32_t cur_heap = heaps.back().first.page, cur_aux = heaps.back().second.page;
    size_t idx_for_calldata = heaps.size() - 1;
    heaps.push_back({HeapPage{heap_page, {}}, HeapPage{aux_page, {}}});
    indirections_to_cleanup_on_return.emplace_back();
    if (calldata.memory_page == 0) {
    } else if (calldata.memory_page == cur_heap) {
      page_numbers_indirections[cur_heap] = Indirection{Indirection::HEAP, idx_for_calldata};
      indirections_to_cleanup_on_return.back().insert(cur_heap);
    } else if (calldata.memory_page == cur_aux) {
      page_numbers_indirections[cur_aux] = Indirection{Indirection::AUX_HEAP, idx_for_calldata};
      indirections_to_cleanup_on_return.back().insert(cur_aux);
    } else {
      auto it = page_numbers_indirections.find(calldata.memory_page);
      REF_ASSERT(it != page_numbers_indirections.end(), "fat pointer must only point to reachable memory (memory.rs:641)");
      REF_ASSERT(it->second.kind == Indirection::HEAP || it->second.kind == Indirection::AUX_HEAP,
                 "calldata forwarding should already have a heap/aux heap indirection (memory.rs:645)");
    }
  }

  // memory.rs:660-758 finish_global_frame
  void finish_global_frame(uint32_t base_page, FatPointer returndata) {
    REF_ASSERT(!stack_pages.empty() && stack_pages.back().page == stack_page_from_base(base_page), "stack page (memory.rs:673)");
    stack_pages.pop_back();
    uint32_t returndata_page = returndata.memory_page;
    auto cur = std::move(heaps.back());
    heaps.pop_back();
    REF_ASSERT(cur.first.page == heap_page_from_base(base_page), "heap page (memory.rs:690)");
    REF_ASSERT(cur.second.page == aux_heap_page_from_base(base_page), "aux page (memory.rs:691)");
    auto current_cleanup = std::move(indirections_to_cleanup_on_return.back());
    indirections_to_cleanup_on_return.pop_back();
    REF_ASSERT(!indirections_to_cleanup_on_return.empty(), "previous page indirections must exist");
    auto& previous_cleanup = indirections_to_cleanup_on_return.back();
    if (returndata_page == cur.first.page) {
      REF_ASSERT(!pages_with_extended_lifetime.count(cur.first.page), "existing extended page (memory.rs:707)");
      pages_with_extended_lifetime[cur.first.page] = std::move(cur.first.content);
      page_numbers_indirections[cur.first.page] = Indirection{Indirection::RETURNDATA_EXTENDED_LIFETIME, 0};
      previous_cleanup.insert(cur.first.page);
    } else if (returndata_page == cur.second.page) {
      REF_ASSERT(!pages_with_extended_lifetime.count(cur.second.page), "existing extended page (memory.rs:719)");
      pages_with_extended_lifetime[cur.second.page] = std::move(cur.second.content);
      page_numbers_indirections[cur.second.page] = Indirection{Indirection::RETURNDATA_EXTENDED_LIFETIME, 0};
      previous_cleanup.insert(cur.second.page);
    } else if (returndata_page != 0) {
      REF_ASSERT(page_numbers_indirections.count(returndata_page), "expected that indirections contain page (memory.rs:735)");
      current_cleanup.erase(returndata_page);
      previous_cleanup.insert(returndata_page);
    }
    for (uint32_t el : current_cleanup) {
      REF_ASSERT(page_numbers_indirections.erase(el) == 1, "double free in indirection (memory.rs:755)");
    }
  }
};

// ---------------------------------------------------------------------------------------------------
// InMemoryStorage (src/testing/storage.rs) and InMemoryEventSink (src/reference_impls/event_sink.rs)
// ---------------------------------------------------------------------------------------------------
struct SlotKey {
  uint8_t shard;
  Address address;
  U256 key;
  bool operator==(const SlotKey& o) const { return shard == o.shard && address == o.address && key == o.key; }
};
struct SlotKeyHash {
  size_t operator()(const SlotKey& k) const {
    uint64_t h = 0x9E3779B97F4A7C15ull * (k.shard + 1);
    for (int i = 0; i < 4; i++) h = (h ^ k.key.w[i]) * 0xff51afd7ed558ccdULL + (h >> 29);
    uint64_t a0, a1;
    uint32_t a2;
    memcpy(&a0, k.address.b, 8);
    memcpy(&a1, k.address.b + 8, 8);
    memcpy(&a2, k.address.b + 16, 4);
    h = (h ^ a0) * 0xc4ceb9fe1a85ec53ULL;
    h = (h ^ a1) * 0xff51afd7ed558ccdULL;
    h ^= a2;
    return (size_t)(h ^ (h >> 32));
  }
};

struct ApplicationData {
  std::vector<LogQuery> forward, rollbacks;
};

struct InMemoryStorage {
  std::unordered_map<SlotKey, U256, SlotKeyHash> inner;
  std::unordered_set<SlotKey, SlotKeyHash> cold_warm_markers;  // storage.rs:10: set on every executed read / write, never rolled back
  std::vector<ApplicationData> frames_stack;
  // row f-3 (refund-aware oracle): 0 = the reference's test storage (always RefundType::None, storage.rs:80-86);
  // 1..64 = RepeatedWrite refund, in pubdata bytes, for a write to an already warm slot of the rollup shard
  uint32_t warm_write_refund_bytes = 0;
  InMemoryStorage() { frames_stack.emplace_back(); }  // storage.rs:18-24

  // Storage::estimate_refunds_for_write (storage.rs:80-86 answers None; the policy above is this build's f-3 oracle).
  // Returns (refund type, pubdata bytes); called BEFORE the write executes (log.rs:99-102).
  std::pair<uint32_t, uint32_t> estimate_refunds_for_write(const LogQuery& q) const {
    if (warm_write_refund_bytes == 0 || q.shard_id != 0) return {0u, 0u};
    SlotKey k{q.shard_id, q.address, q.key};
    if (!cold_warm_markers.count(k)) return {0u, 0u};
    return {1u, warm_write_refund_bytes};
  }

  // storage.rs:88-139
  LogQuery execute_partial_query(LogQuery q) {
    REF_ASSERT(q.shard_id < 2, "shard id out of range (storage.rs:93 index)");
    REF_ASSERT(!frames_stack.empty(), "frame must be started");
    ApplicationData& frame = frames_stack.back();
    REF_ASSERT(!q.rollback, "rollback query");
    SlotKey k{q.shard_id, q.address, q.key};
    auto it = inner.find(k);
    U256 current = it == inner.end() ? U256::zero() : it->second;
    cold_warm_markers.insert(k);  // storage.rs:105-110 (writes), :126-131 (reads)
    if (q.rw_flag) {
      inner[k] = q.written_value;
      q.read_value = current;
      frame.forward.push_back(q);
      q.rollback = true;
      frame.rollbacks.push_back(q);
      q.rollback = false;
    } else {
      q.read_value = current;
      frame.forward.push_back(q);
    }
    return q;
  }
  void start_frame() { frames_stack.emplace_back(); }  // storage.rs:140-143
  // storage.rs:144-186
  void finish_frame(bool panicked) {
    REF_ASSERT(frames_stack.size() >= 2, "parent frame must exist");
    ApplicationData cur = std::move(frames_stack.back());
    frames_stack.pop_back();
    ApplicationData& parent = frames_stack.back();
    if (panicked) {
      for (auto it = cur.rollbacks.rbegin(); it != cur.rollbacks.rend(); ++it) {
        SlotKey k{it->shard_id, it->address, it->key};
        auto sit = inner.find(k);
        REF_ASSERT(sit != inner.end(), "must always exist on rollback");
        REF_ASSERT(sit->second == it->written_value, "rollback value mismatch (storage.rs:173)");
        sit->second = it->read_value;
      }
      parent.forward.insert(parent.forward.end(), cur.forward.begin(), cur.forward.end());
      parent.forward.insert(parent.forward.end(), cur.rollbacks.rbegin(), cur.rollbacks.rend());
    } else {
      parent.forward.insert(parent.forward.end(), cur.forward.begin(), cur.forward.end());
      parent.rollbacks.insert(parent.rollbacks.end(), cur.rollbacks.begin(), cur.rollbacks.end());
    }
  }
};

struct InMemoryEventSink {
  std::vector<ApplicationData> frames_stack;
  InMemoryEventSink() { frames_stack.emplace_back(); }
  // event_sink.rs:139-149
  void add_partial_query(LogQuery q) {
    REF_ASSERT(q.rw_flag && !q.rollback, "event query flags");
    REF_ASSERT(q.aux_byte == ZK_EVENT_AUX_BYTE || q.aux_byte == ZK_L1_MESSAGE_AUX_BYTE, "event aux byte");
    ApplicationData& frame = frames_stack.back();
    frame.forward.push_back(q);
    q.rollback = true;
    frame.rollbacks.push_back(q);
  }
  void start_frame() { frames_stack.emplace_back(); }
  // event_sink.rs:154-175
  void finish_frame(bool panicked) {
    REF_ASSERT(frames_stack.size() >= 2, "parent frame must exist");
    ApplicationData cur = std::move(frames_stack.back());
    frames_stack.pop_back();
    ApplicationData& parent = frames_stack.back();
    parent.forward.insert(parent.forward.end(), cur.forward.begin(), cur.forward.end());
    if (panicked)
      parent.forward.insert(parent.forward.end(), cur.rollbacks.rbegin(), cur.rollbacks.rend());
    else
      parent.rollbacks.insert(parent.rollbacks.end(), cur.rollbacks.begin(), cur.rollbacks.end());
  }
};

// ---------------------------------------------------------------------------------------------------
// SimpleDecommitter (src/reference_impls/decommitter.rs); known_hashes is shared by the whole batch
// ---------------------------------------------------------------------------------------------------
struct U256Hash {
  size_t operator()(const U256& v) const { return (size_t)(v.w[0] ^ (v.w[1] * 3) ^ (v.w[2] * 5) ^ (v.w[3] * 7)); }
};
typedef std::unordered_map<U256, CodeRef, U256Hash> KnownHashes;

struct SimpleDecommitter {
  const KnownHashes* known_hashes = nullptr;
  std::unordered_map<U256, std::pair<uint32_t, uint16_t>, U256Hash> history;
  // decommitter.rs:32-99 (B = true flavour: the tracer is always called, quirk 14)
  DecommittmentQuery decommit_into_memory(DecommittmentQuery q, SimpleMemory& memory) {
    auto h = history.find(q.hash);
    if (h != history.end()) {
      q.is_fresh = false;
      q.memory_page = h->second.first;
      q.decommitted_length = h->second.second;
      return q;
    }
    auto it = known_hashes->find(q.hash);
    if (it == known_hashes->end()) throw UnknownCodeHash("Code hash must be known (decommitter.rs:50-56)");
    q.decommitted_length = (uint16_t)it->second->size();
    q.is_fresh = true;
    history[q.hash] = {q.memory_page, q.decommitted_length};
    memory.code_pages[q.memory_page] = it->second;  // == specialized_code_query per word (decommitter.rs:81-97)
    return q;
  }
};

// ---------------------------------------------------------------------------------------------------
// DefaultPrecompilesProcessor (EXTERNAL zk_evm_abstractions::precompiles; reconstruction)
// ---------------------------------------------------------------------------------------------------
struct PrecompileCallABI {
  uint32_t input_memory_offset, input_memory_length, output_memory_offset, output_memory_length;
  uint32_t memory_page_to_read, memory_page_to_write;
  uint64_t precompile_interpreted_data;
  static PrecompileCallABI from_u256(const U256& v) {
    return PrecompileCallABI{(uint32_t)v.w[0], (uint32_t)(v.w[0] >> 32), (uint32_t)v.w[1], (uint32_t)(v.w[1] >> 32),
                             (uint32_t)v.w[2], (uint32_t)(v.w[2] >> 32), v.w[3]};
  }
  U256 to_u256() const {
    return U256{{(uint64_t)input_memory_offset | ((uint64_t)input_memory_length << 32),
                 (uint64_t)output_memory_offset | ((uint64_t)output_memory_length << 32),
                 (uint64_t)memory_page_to_read | ((uint64_t)memory_page_to_write << 32), precompile_interpreted_data}};
  }
};

struct PrecompileResult {
  bool executed = false;
  std::vector<MemoryQuery> mem_in, mem_out;
};

// keccak256 v1.4.1 memory ABI as pinned by keccak256.rs:100-139: input byte offset/length (unaligned allowed),
// padding done by the precompile, one output word at WORD index output_memory_offset.  Reads go through the
// FatPointer memory type (the test registers an indirection for the page, keccak256.rs:87-90), one query per
// distinct 32-byte word in ascending order at `timestamp`; the write is a Heap query at `timestamp + 1`.
static PrecompileResult keccak256_precompile(const LogQuery& query, SimpleMemory& memory) {
  PrecompileResult res;
  res.executed = true;
  PrecompileCallABI p = PrecompileCallABI::from_u256(query.key);
  uint32_t ts_read = query.timestamp, ts_write = query.timestamp + 1;
  std::vector<uint8_t> input;
  input.reserve(p.input_memory_length);
  if (p.input_memory_length > 0) {
    uint64_t first = p.input_memory_offset / 32;
    uint64_t last = ((uint64_t)p.input_memory_offset + p.input_memory_length - 1) / 32;
    for (uint64_t wi = first; wi <= last; wi++) {
      MemoryQuery q{ts_read, ZK_MEM_FAT_PTR, p.memory_page_to_read, (uint32_t)wi, U256::zero(), false, false};
      q = memory.execute_partial_query(q);
      res.mem_in.push_back(q);
      uint8_t buf[32];
      q.value.to_be(buf);
      uint64_t lo = wi * 32, hi = lo + 32;
      uint64_t a = std::max<uint64_t>(lo, p.input_memory_offset);
      uint64_t b = std::min<uint64_t>(hi, (uint64_t)p.input_memory_offset + p.input_memory_length);
      input.insert(input.end(), buf + (a - lo), buf + (b - lo));
    }
  }
  uint8_t digest[32];
  orc_hash::keccak_sponge256(input.data(), input.size(), 0x01, digest);
  MemoryQuery w{ts_write, ZK_MEM_HEAP, p.memory_page_to_write, p.output_memory_offset, U256::from_be(digest), false, true};
  w = memory.execute_partial_query(w);
  res.mem_out.push_back(w);
  return res;
}

// sha256 round-function precompile (reconstruction; parity unpinned): input = pre-padded 64-byte blocks,
// input_memory_offset in WORDS, number of rounds = precompile_interpreted_data, 2 Heap-type word reads per round,
// digest written to word output_memory_offset after the last round.
static PrecompileResult sha256_precompile(const LogQuery& query, SimpleMemory& memory) {
  PrecompileResult res;
  res.executed = true;
  PrecompileCallABI p = PrecompileCallABI::from_u256(query.key);
  uint32_t ts_read = query.timestamp, ts_write = query.timestamp + 1;
  uint32_t h[8];
  memcpy(h, orc_hash::SHA256_IV, sizeof(h));
  uint64_t rounds = p.precompile_interpreted_data;
  for (uint64_t r = 0; r < rounds; r++) {
    uint8_t block[64];
    for (int k = 0; k < 2; k++) {
      MemoryQuery q{ts_read, ZK_MEM_HEAP, p.memory_page_to_read, (uint32_t)(p.input_memory_offset + 2 * r + k), U256::zero(), false, false};
      q = memory.execute_partial_query(q);
      res.mem_in.push_back(q);
      q.value.to_be(block + 32 * k);
    }
    orc_hash::sha256_compress(h, block);
  }
  if (rounds > 0) {
    uint8_t digest[32];
    for (int i = 0; i < 8; i++) {
      digest[4 * i] = (uint8_t)(h[i] >> 24);
      digest[4 * i + 1] = (uint8_t)(h[i] >> 16);
      digest[4 * i + 2] = (uint8_t)(h[i] >> 8);
      digest[4 * i + 3] = (uint8_t)h[i];
    }
    MemoryQuery w{ts_write, ZK_MEM_HEAP, p.memory_page_to_write, p.output_memory_offset, U256::from_be(digest), false, true};
    w = memory.execute_partial_query(w);
    res.mem_out.push_back(w);
  }
  return res;
}

// ecrecover precompile (external; memory ABI reconstructed): four Heap-type word reads at
// input_memory_offset (digest, v in {0, 1}, r, s) at `timestamp`, two Heap-type word writes at output_memory_offset
// (success marker, then the address right-aligned in the word; both zero on failure) at `timestamp + 1`.
static bool ecrecover_address(const U256& hash, const U256& r, const U256& s, bool v_odd, U256* address) {
  U256 qx, qy;
  if (!orc_secp::recover(hash, r, s, v_odd, &qx, &qy)) return false;
  uint8_t pk[64], digest[32];
  qx.to_be(pk);
  qy.to_be(pk + 32);
  orc_hash::keccak_sponge256(pk, 64, 0x01, digest);
  memset(digest, 0, 12);
  *address = U256::from_be(digest);
  return true;
}

static PrecompileResult ecrecover_precompile(const LogQuery& query, SimpleMemory& memory) {
  PrecompileResult res;
  res.executed = true;
  PrecompileCallABI p = PrecompileCallABI::from_u256(query.key);
  uint32_t ts_read = query.timestamp, ts_write = query.timestamp + 1;
  U256 in[4];
  for (uint32_t i = 0; i < 4; i++) {
    MemoryQuery q{ts_read, ZK_MEM_HEAP, p.memory_page_to_read, p.input_memory_offset + i, U256::zero(), false, false};
    q = memory.execute_partial_query(q);
    res.mem_in.push_back(q);
    in[i] = q.value;
  }
  const U256& v = in[1];
  REF_ASSERT((v.w[1] | v.w[2] | v.w[3]) == 0 && v.w[0] <= 1, "ecrecover: v must be 0 or 1 (assert in the external precompile)");
  U256 address = U256::zero();
  bool ok = ecrecover_address(in[0], in[2], in[3], v.w[0] == 1, &address);
  if (!ok) address = U256::zero();
  MemoryQuery w0{ts_write, ZK_MEM_HEAP, p.memory_page_to_write, p.output_memory_offset, U256::from_u64(ok ? 1 : 0), false, true};
  res.mem_out.push_back(memory.execute_partial_query(w0));
  MemoryQuery w1{ts_write, ZK_MEM_HEAP, p.memory_page_to_write, p.output_memory_offset + 1, address, false, true};
  res.mem_out.push_back(memory.execute_partial_query(w1));
  return res;
}

static PrecompileResult execute_precompile(const LogQuery& query, SimpleMemory& memory) {
  uint16_t address_low = (uint16_t)(query.address.b[19] | (query.address.b[18] << 8));
  switch (address_low) {
    case ZK_KECCAK256_PRECOMPILE_ADDRESS:
      return keccak256_precompile(query, memory);
    case ZK_SHA256_PRECOMPILE_ADDRESS:
      return sha256_precompile(query, memory);
    case ZK_ECRECOVER_PRECOMPILE_ADDRESS:
      return ecrecover_precompile(query, memory);
    default:
      return PrecompileResult{};
  }
}

// ---------------------------------------------------------------------------------------------------
// decoded opcode (zkevm_opcode_defs::DecodedOpcode; src/opcodes/parsing.rs:5-7)
// ---------------------------------------------------------------------------------------------------
struct DecodedOpcode {
  uint32_t variant_idx, entry;
  uint8_t condition, src0_reg_idx, src1_reg_idx, dst0_reg_idx, dst1_reg_idx;
  uint16_t imm_0, imm_1;
  uint32_t family() const { return entry & 15; }
  uint32_t sub() const { return (entry >> ZK_E_SUB_SHIFT) & 15; }
  uint32_t src_mode() const { return (entry >> ZK_E_SRC_SHIFT) & 7; }
  uint32_t dst_mode() const { return (entry >> ZK_E_DST_SHIFT) & 3; }
  bool flag0() const { return entry & ZK_E_FLAG0; }
  bool flag1() const { return entry & ZK_E_FLAG1; }
  static DecodedOpcode parse(uint64_t raw) {
    DecodedOpcode d;
    d.variant_idx = (uint32_t)(raw & ((1u << ZK_VARIANT_BITS) - 1));
    d.entry = ZK_OPCODE_TABLE[d.variant_idx];
    d.condition = (raw >> ZK_COND_SHIFT) & 7;
    d.src0_reg_idx = (raw >> 16) & 15;
    d.src1_reg_idx = (raw >> 20) & 15;
    d.dst0_reg_idx = (raw >> 24) & 15;
    d.dst1_reg_idx = (raw >> 28) & 15;
    d.imm_0 = (uint16_t)(raw >> 32);
    d.imm_1 = (uint16_t)(raw >> 48);
    return d;
  }
  void mask_to(uint32_t idx) {
    variant_idx = idx;
    entry = ZK_OPCODE_TABLE[idx];
    condition = 0;
    src0_reg_idx = src1_reg_idx = dst0_reg_idx = dst1_reg_idx = 0;
    imm_0 = imm_1 = 0;
  }
  void mask_into_panic() { mask_to(ZK_PANIC_VARIANT_IDX); }  // parsing.rs:39-41
  void mask_into_nop() { mask_to(ZK_NOP_VARIANT_IDX); }      // parsing.rs:43-45
};

struct MemoryLocation {
  uint8_t memory_type;
  uint32_t page, index;
};

// helpers.rs:344-352
enum { ERR_INVALID_OPCODE = 1, ERR_NOT_ENOUGH_ERGS = 2, ERR_PRIVILEGED = 4, ERR_WRITE_IN_STATIC = 8, ERR_CALLSTACK_FULL = 16 };

struct BlockProperties {
  U256 default_aa_code_hash = U256::zero();
  bool zkporter_is_available = false;
};

// ---------------------------------------------------------------------------------------------------
// VmState (src/vm_state/mod.rs:54-73,157-175)
// ---------------------------------------------------------------------------------------------------
struct VmState {
  // VmLocalState
  U256 previous_code_word = U256::zero();
  uint32_t previous_code_memory_page = 0;
  PrimitiveValue registers[ZK_REGISTERS_COUNT];
  Flags flags;
  uint32_t timestamp = ZK_STARTING_TIMESTAMP;
  uint32_t monotonic_cycle_counter = 0;
  uint32_t spent_pubdata_counter = 0;
  uint32_t memory_page_counter = ZK_STARTING_BASE_PAGE;
  uint32_t absolute_execution_step = 0;
  uint32_t current_ergs_per_pubdata_byte = 0;
  uint16_t tx_number_in_block = 0;
  bool pending_exception = false;
  uint16_t previous_super_pc = 0;
  uint32_t context_u128_register[4] = {0, 0, 0, 0};
  CallStackEntry current = CallStackEntry::empty_context();
  std::vector<CallStackEntry> inner;

  BlockProperties block_properties;
  InMemoryStorage storage;
  SimpleMemory memory;
  InMemoryEventSink event_sink;
  SimpleDecommitter decommitter;
  Recorder wt;

  uint32_t status = ZKB_VM_RUNNING;

  // per-cycle row scratch
  ZkbCycleRow row;

  VmState() {
    for (auto& r : registers) r = PrimitiveValue::empty();
  }

  bool execution_has_ended() const { return inner.empty(); }                          // mod.rs:96-98
  bool callstack_is_full() const { return inner.size() == ZK_VM_MAX_STACK_DEPTH; }     // execution_stack.rs:119-121
  uint32_t ts_read() const { return timestamp; }                                       // mod.rs:220-222
  uint32_t ts_first_decommit_or_precompile_read() const { return timestamp + 1; }      // :223-225
  uint32_t ts_second_decommit_or_precompile_write() const { return timestamp + 2; }    // :226-228
  uint32_t ts_dst_write() const { return timestamp + 3; }                              // :229-231

  // helpers.rs:318-334
  PrimitiveValue select_register_value(uint8_t idx) const { return idx == 0 ? PrimitiveValue::empty() : registers[idx - 1]; }
  void update_register_value(uint8_t idx, const PrimitiveValue& v) {
    if (idx > 0) registers[idx - 1] = v;
  }
  void set_shorthand_panic() { pending_exception = true; }  // helpers.rs:336-338

  // helpers.rs:10-40
  MemoryQuery read_code(MemoryLocation loc, uint32_t ts) {
    MemoryQuery q{ts, loc.memory_type, loc.page, loc.index, U256::zero(), false, false};
    q = memory.read_code_query(q);
    wt.add_memory_query(q, ZKB_MEMORIGIN_VM);
    return q;
  }
  // helpers.rs:53-76
  MemoryQuery read_memory(MemoryLocation loc, uint32_t ts) {
    MemoryQuery q{ts, loc.memory_type, loc.page, loc.index, U256::zero(), false, false};
    q = memory.execute_partial_query(q);
    wt.add_memory_query(q, ZKB_MEMORIGIN_VM);
    return q;
  }
  // helpers.rs:87-117
  MemoryQuery write_memory(MemoryLocation loc, uint32_t ts, const PrimitiveValue& v) {
    MemoryQuery q{ts, loc.memory_type, loc.page, loc.index, v.value, v.is_pointer, true};
    q = memory.execute_partial_query(q);
    wt.add_memory_query(q, ZKB_MEMORIGIN_VM);
    return q;
  }
  // helpers.rs:119-136: ask the storage oracle, hand the answer to the tracer, return RefundType::pubdata_refund()
  uint32_t refund_for_partial_query(const LogQuery& q) {
    REF_ASSERT(q.rw_flag, "refund for read");
    auto refund = storage.estimate_refunds_for_write(q);
    wt.record_refund(refund.first, refund.second);
    return refund.second;
  }
  // helpers.rs:138-155
  LogQuery access_storage(LogQuery q) {
    q = storage.execute_partial_query(q);
    if (!q.rw_flag) q.written_value = q.read_value;
    wt.add_log_query(q);
    return q;
  }
  // helpers.rs:157-162
  void emit_event(const LogQuery& q) {
    event_sink.add_partial_query(q);
    wt.add_log_query(q);
  }
  // helpers.rs:164-194
  DecommittmentQuery decommit(const U256& hash, uint32_t candidate_page, uint32_t ts) {
    DecommittmentQuery q{hash, ts, candidate_page, 0, false};
    q = decommitter.decommit_into_memory(q, memory);
    wt.add_decommittment(q);
    return q;
  }
  // helpers.rs:196-223
  void call_precompile(const LogQuery& q) {
    wt.add_log_query(q);
    PrecompileResult r = execute_precompile(q, memory);
    if (r.executed) {
      for (auto& m : r.mem_in) wt.add_memory_query(m, ZKB_MEMORIGIN_PRECOMPILE_IN);
      for (auto& m : r.mem_out) wt.add_memory_query(m, ZKB_MEMORIGIN_PRECOMPILE_OUT);
    }
  }
  // helpers.rs:225-246
  void start_frame(const CallStackEntry& e) {
    storage.start_frame();
    event_sink.start_frame();
    wt.start_new_execution_context(monotonic_cycle_counter, current, e);
    inner.push_back(current);  // execution_stack.rs:103-109 push_entry
    current = e;
  }
  // helpers.rs:248-264
  CallStackEntry finish_frame(bool panicked) {
    storage.finish_frame(panicked);
    event_sink.finish_frame(panicked);
    wt.finish_execution_context(monotonic_cycle_counter, panicked);
    REF_ASSERT(!inner.empty(), "pop from empty callstack (execution_stack.rs:113 unwrap)");
    CallStackEntry old = current;
    current = inner.back();
    inner.pop_back();
    return old;
  }
  // helpers.rs:266-283
  void perform_dst0_update(const PrimitiveValue& v, const MemoryLocation* loc, const DecodedOpcode& op) {
    v.value.to_limbs32(row.dst0);
    row.bits |= ZKB_ROWBIT_DST0_VALID | (v.is_pointer ? ZKB_ROWBIT_DST0_PTR : 0);
    if (loc)
      write_memory(*loc, ts_dst_write(), v);
    else
      update_register_value(op.dst0_reg_idx, v);
  }
  // helpers.rs:285-287
  void perform_dst1_update(const PrimitiveValue& v, uint8_t idx) {
    v.value.to_limbs32(row.dst1);
    row.bits |= ZKB_ROWBIT_DST1_VALID | (v.is_pointer ? ZKB_ROWBIT_DST1_PTR : 0);
    update_register_value(idx, v);
  }
  // helpers.rs:289-316
  void push_bootloader_context(const CallStackEntry& boot) {
    REF_ASSERT(current.ergs_remaining >= boot.ergs_remaining, "bootloader frame with more ergs than VM has available");
    current.ergs_remaining -= boot.ergs_remaining;
    start_frame(boot);
    memory.start_global_frame(ZK_UNMAPPED_PAGE, boot.base_memory_page, FatPointer::empty());
  }

  // mem_ops.rs:14-125
  bool compute_address(uint16_t& sp, uint8_t reg_idx, uint16_t imm, bool is_write, uint32_t mode, MemoryLocation* loc, PrimitiveValue* reg_out) {
    PrimitiveValue pv = select_register_value(reg_idx);
    if (reg_out) *reg_out = pv;
    uint16_t reg_low = (uint16_t)pv.value.low_u64();
    uint16_t vaddr = (uint16_t)(reg_low + imm);
    uint32_t stack_page = stack_page_from_base(current.base_memory_page);
    if (!is_write) {
      switch (mode) {
        case ZK_SRC_STACK_POP: {
          uint16_t new_sp = (uint16_t)(sp - vaddr);
          sp = new_sp;
          *loc = MemoryLocation{ZK_MEM_STACK, stack_page, new_sp};
          return true;
        }
        case ZK_SRC_STACK_REL:
          *loc = MemoryLocation{ZK_MEM_STACK, stack_page, (uint16_t)(sp - vaddr)};
          return true;
        case ZK_SRC_STACK_ABS:
          *loc = MemoryLocation{ZK_MEM_STACK, stack_page, vaddr};
          return true;
        case ZK_SRC_CODE:
          *loc = MemoryLocation{ZK_MEM_CODE, current.code_page, vaddr};
          return true;
        default:
          return false;  // SRC_REG, SRC_IMM
      }
    } else {
      switch (mode) {
        case ZK_DST_STACK_PUSH: {
          uint16_t old_sp = sp;
          sp = (uint16_t)(sp + vaddr);
          *loc = MemoryLocation{ZK_MEM_STACK, stack_page, old_sp};
          return true;
        }
        case ZK_DST_STACK_REL:
          *loc = MemoryLocation{ZK_MEM_STACK, stack_page, (uint16_t)(sp - vaddr)};
          return true;
        case ZK_DST_STACK_ABS:
          *loc = MemoryLocation{ZK_MEM_STACK, stack_page, vaddr};
          return true;
        default:
          return false;
      }
    }
  }

  struct PreState {
    PrimitiveValue src0, src1;
    bool has_dst0_loc;
    MemoryLocation dst0_loc;
    uint16_t new_pc;
    bool is_kernel_mode;
    const MemoryLocation* dst0() const { return has_dst0_loc ? &dst0_loc : nullptr; }
  };

  void cycle();
  // handlers (src/opcodes/execution/*.rs)
  void op_nop(const DecodedOpcode&, const PreState&);
  void op_add(const DecodedOpcode&, const PreState&);
  void op_sub(const DecodedOpcode&, const PreState&);
  void op_mul(const DecodedOpcode&, const PreState&);
  void op_div(const DecodedOpcode&, const PreState&);
  void op_jump(const DecodedOpcode&, const PreState&);
  void op_context(const DecodedOpcode&, const PreState&);
  void op_shift(const DecodedOpcode&, const PreState&);
  void op_binop(const DecodedOpcode&, const PreState&);
  void op_ptr(const DecodedOpcode&, const PreState&);
  void op_near_call(const DecodedOpcode&, const PreState&);
  void op_log(const DecodedOpcode&, const PreState&);
  void op_far_call(const DecodedOpcode&, const PreState&);
  void op_ret(const DecodedOpcode&, const PreState&);
  void op_uma(const DecodedOpcode&, const PreState&);
};

// ---------------------------------------------------------------------------------------------------
// cycle.rs:19-236 read_and_decode + cycle.rs:257-429 cycle
// ---------------------------------------------------------------------------------------------------
void VmState::cycle() {
  memset(&row, 0, sizeof(row));
  uint64_t counts_before[ZKB_N_STREAMS];
  memcpy(counts_before, wt.counts, sizeof(counts_before));
  row.cycle = monotonic_cycle_counter;
  row.timestamp = timestamp;
  row.pc_before = current.pc;

  // ---- read_and_decode (delayed changes are applied right after, cycle.rs:267) ----
  bool ended = execution_has_ended();
  bool pending = pending_exception;
  uint32_t code_page = current.code_page;
  uint32_t new_prev_code_page = code_page;                         // cycle.rs:49
  bool has_new_code_word = false, has_new_super_pc = false, clear_pending = false;
  U256 new_code_word = U256::zero();
  uint16_t new_super_pc = 0;
  uint16_t pc = current.pc;
  bool code_pages_are_different = current.code_page != previous_code_memory_page;
  uint16_t super_pc = pc >> 2;
  uint8_t sub_pc = pc & 3;
  uint64_t opcode_encoding;
  if (!ended && !pending) {
    U256 word;
    if (code_pages_are_different || previous_super_pc != super_pc) {
      MemoryQuery q = read_code(MemoryLocation{ZK_MEM_CODE, code_page, super_pc}, ts_read());  // cycle.rs:61-81
      word = q.value;
      has_new_code_word = true;
      new_code_word = word;
      has_new_super_pc = true;
      new_super_pc = super_pc;
    } else {
      word = previous_code_word;  // cycle.rs:96-100
    }
    opcode_encoding = word.w[3 - sub_pc];  // E::integer_representaiton_from_u256 (cycle.rs:86-94 comment)
  } else if (pending) {
    REF_ASSERT(!ended, "pending exception at end of execution (cycle.rs:107)");
    clear_pending = true;
    has_new_super_pc = true;
    new_super_pc = super_pc;
    opcode_encoding = ZK_EXCEPTION_REVERT_ENCODING;
  } else {
    opcode_encoding = ZK_NOP_ENCODING;
  }
  bool skip_cycle = ended;
  row.raw_opcode = opcode_encoding;

  uint32_t error_flags = 0;
  DecodedOpcode op = DecodedOpcode::parse(opcode_encoding);
  uint32_t raw_variant_idx = op.variant_idx;
  if (op.entry & ZK_E_INVALID) error_flags |= ERR_INVALID_OPCODE;  // cycle.rs:142-144
  uint32_t ergs_cost = ZK_OPCODE_PRICES[raw_variant_idx];
  if (skip_cycle) ergs_cost = 0;
  uint32_t ergs_remaining;
  if (current.ergs_remaining < ergs_cost) {
    ergs_remaining = 0;
    error_flags |= ERR_NOT_ENOUGH_ERGS;
  } else {
    ergs_remaining = current.ergs_remaining - ergs_cost;
  }
  bool is_kernel = current.is_kernel_mode();
  if ((op.entry & ZK_E_KERNEL_ONLY) && !is_kernel) error_flags |= ERR_PRIVILEGED;
  if ((op.entry & ZK_E_STATIC_FORBIDDEN) && current.is_static) error_flags |= ERR_WRITE_IN_STATIC;
  if (callstack_is_full()) error_flags |= ERR_CALLSTACK_FULL;
  bool mask_into_panic = error_flags != 0;
  if (mask_into_panic) op.mask_into_panic();
  bool resolved;
  switch (op.condition) {  // cycle.rs:193-210
    case 0: resolved = true; break;
    case 1: resolved = flags.gt; break;
    case 2: resolved = flags.lt; break;
    case 3: resolved = flags.eq; break;
    case 4: resolved = flags.gt | flags.eq; break;
    case 5: resolved = flags.lt | flags.eq; break;
    case 6: resolved = !flags.eq; break;
    default: resolved = flags.gt | flags.lt; break;
  }
  if (!resolved && !mask_into_panic) op.mask_into_nop();
  row.masked_variant = (uint16_t)op.variant_idx;
  row.cond_resolved = resolved;
  row.error_flags = (uint8_t)error_flags;

  // ---- DelayedLocalStateChanges::apply (mod.rs:134-153) ----
  current.ergs_remaining = ergs_remaining;
  if (has_new_code_word) previous_code_word = new_code_word;
  if (has_new_super_pc) previous_super_pc = new_super_pc;
  if (clear_pending) pending_exception = false;
  previous_code_memory_page = new_prev_code_page;

  // ---- operand addressing (cycle.rs:275-301) ----
  uint16_t sp = current.sp;
  MemoryLocation src0_loc, dst0_loc;
  PrimitiveValue src0_reg_value;
  bool has_src0_loc = compute_address(sp, op.src0_reg_idx, op.imm_0, false, op.src_mode(), &src0_loc, &src0_reg_value);
  bool has_dst0_loc = compute_address(sp, op.dst0_reg_idx, op.imm_1, true, op.dst_mode(), &dst0_loc, nullptr);
  current.sp = sp;
  if (op.family() == ZK_OP_NOP) has_src0_loc = false;

  PrimitiveValue src0_mem_value = PrimitiveValue::empty();
  if (has_src0_loc) {
    MemoryQuery q = src0_loc.memory_type == ZK_MEM_CODE ? read_code(src0_loc, ts_read()) : read_memory(src0_loc, ts_read());
    src0_mem_value = PrimitiveValue{q.value, q.value_is_pointer};
  }
  PrimitiveValue src0;
  switch (op.src_mode()) {  // cycle.rs:327-337
    case ZK_SRC_REG: src0 = src0_reg_value; break;
    case ZK_SRC_IMM: src0 = PrimitiveValue{U256::from_u64(op.imm_0), false}; break;
    default: src0 = src0_mem_value; break;
  }
  PrimitiveValue src1 = select_register_value(op.src1_reg_idx);
  if (op.entry & ZK_E_SWAP) std::swap(src0, src1);  // cycle.rs:341-345
  uint16_t new_pc = current.pc;
  if (!skip_cycle) new_pc = (uint16_t)(new_pc + 1);
  bool is_kernel_mode = current.is_kernel_mode();
  if (src0.is_pointer) row.bits |= ZKB_ROWBIT_SRC0_PTR;
  if (src1.is_pointer) row.bits |= ZKB_ROWBIT_SRC1_PTR;
  // cycle.rs:374-396 erase_fat_pointer_metadata: zero the top 128 bits
  if (!(op.entry & ZK_E_SRC0_PTR_OK) && src0.is_pointer && !is_kernel_mode) {
    src0.value.w[2] = src0.value.w[3] = 0;
    src0.is_pointer = false;
  }
  if (!(op.entry & ZK_E_SRC1_PTR_OK) && src1.is_pointer && !is_kernel_mode) {
    src1.value.w[2] = src1.value.w[3] = 0;
    src1.is_pointer = false;
  }
  src0.value.to_limbs32(row.src0);
  src1.value.to_limbs32(row.src1);
  PreState pre{src0, src1, has_dst0_loc, dst0_loc, new_pc, is_kernel_mode};

  // ---- DecodedOpcode::apply (parsing.rs:47-79) ----
  switch (op.family()) {
    case ZK_OP_NOP: op_nop(op, pre); break;
    case ZK_OP_ADD: op_add(op, pre); break;
    case ZK_OP_SUB: op_sub(op, pre); break;
    case ZK_OP_MUL: op_mul(op, pre); break;
    case ZK_OP_DIV: op_div(op, pre); break;
    case ZK_OP_JUMP: op_jump(op, pre); break;
    case ZK_OP_CONTEXT: op_context(op, pre); break;
    case ZK_OP_SHIFT: op_shift(op, pre); break;
    case ZK_OP_BINOP: op_binop(op, pre); break;
    case ZK_OP_PTR: op_ptr(op, pre); break;
    case ZK_OP_NEAR_CALL: op_near_call(op, pre); break;
    case ZK_OP_LOG: op_log(op, pre); break;
    case ZK_OP_FAR_CALL: op_far_call(op, pre); break;
    case ZK_OP_RET: op_ret(op, pre); break;
    case ZK_OP_UMA: op_uma(op, pre); break;
    default: throw RefPanic("unreachable: Invalid opcode after masking (parsing.rs:77)");
  }
  if (!skip_cycle) timestamp += ZK_TIME_DELTA_PER_CYCLE;
  monotonic_cycle_counter += 1;

  // ---- end_execution_cycle: finish the row ----
  row.pc_after = current.pc;
  row.sp_after = current.sp;
  row.flags_after = flags.bits();
  if (pending_exception) row.bits |= ZKB_ROWBIT_PENDING;
  if (skip_cycle) row.bits |= ZKB_ROWBIT_SKIP;
  row.ergs_after = current.ergs_remaining;
  row.callstack_depth = (uint32_t)inner.size();
  row.spent_pubdata = spent_pubdata_counter;
  row.memory_page_counter = memory_page_counter;
  row.n_mem = (uint16_t)(wt.counts[ZKB_STREAM_MEM] - counts_before[ZKB_STREAM_MEM]);
  row.n_log = (uint8_t)(wt.counts[ZKB_STREAM_LOG] - counts_before[ZKB_STREAM_LOG]);
  row.n_dfr = (uint8_t)((wt.counts[ZKB_STREAM_DECOMMIT] - counts_before[ZKB_STREAM_DECOMMIT]) |
                        ((wt.counts[ZKB_STREAM_FRAME] - counts_before[ZKB_STREAM_FRAME]) << 2) |
                        ((wt.counts[ZKB_STREAM_REFUND] - counts_before[ZKB_STREAM_REFUND]) << 4));
  memcpy(row.context_u128, context_u128_register, 16);
  row.tx_number = tx_number_in_block;
  row.previous_super_pc = previous_super_pc;
  row.ergs_per_pubdata = current_ergs_per_pubdata_byte;
  row.code_page = current.code_page;
  row.base_page = current.base_memory_page;
  row.heap_bound = current.heap_bound;
  row.aux_heap_bound = current.aux_heap_bound;
  row.exception_handler = current.exception_handler_location;
  row.frame_bits = (uint8_t)((current.is_static ? ZKB_FRAMEBIT_STATIC : 0) | (current.is_local_frame ? ZKB_FRAMEBIT_LOCAL : 0) |
                             (current.is_kernel_mode() ? ZKB_FRAMEBIT_KERNEL : 0));
  wt.push(ZKB_STREAM_ROWS, row);
}

// noop.rs:4-20
void VmState::op_nop(const DecodedOpcode&, const PreState& pre) { current.pc = pre.new_pc; }

// add.rs:4-54 (flags are assigned without reset, quirk 5)
void VmState::op_add(const DecodedOpcode& op, const PreState& pre) {
  current.pc = pre.new_pc;
  bool of;
  U256 result = u256_add(pre.src0.value, pre.src1.value, &of);
  bool eq = result.is_zero();
  bool gt = !eq && !of;
  if (op.flag0()) {
    flags.lt = of;
    flags.eq = eq;
    flags.gt = gt;
  }
  perform_dst0_update(PrimitiveValue{result, false}, pre.dst0(), op);
}

// sub.rs:4-55
void VmState::op_sub(const DecodedOpcode& op, const PreState& pre) {
  current.pc = pre.new_pc;
  bool of;
  U256 result = u256_sub(pre.src0.value, pre.src1.value, &of);
  bool eq = result.is_zero();
  bool gt = !eq && !of;
  if (op.flag0()) {
    flags.reset();
    flags.lt = of;
    flags.eq = eq;
    flags.gt = gt;
  }
  perform_dst0_update(PrimitiveValue{result, false}, pre.dst0(), op);
}

// mul.rs:4-66
void VmState::op_mul(const DecodedOpcode& op, const PreState& pre) {
  current.pc = pre.new_pc;
  uint64_t tmp[8];
  u256_full_mul(pre.src0.value, pre.src1.value, tmp);
  U256 low{{tmp[0], tmp[1], tmp[2], tmp[3]}}, high{{tmp[4], tmp[5], tmp[6], tmp[7]}};
  if (op.flag0()) {
    bool of = !high.is_zero(), eq = low.is_zero();
    flags.reset();
    flags.lt = of;
    flags.eq = eq;
    flags.gt = !of & !eq;
  }
  perform_dst0_update(PrimitiveValue{low, false}, pre.dst0(), op);
  perform_dst1_update(PrimitiveValue{high, false}, op.dst1_reg_idx);
}

// div.rs:4-76 (gt means "remainder is zero", quirk 5)
void VmState::op_div(const DecodedOpcode& op, const PreState& pre) {
  current.pc = pre.new_pc;
  if (pre.src1.value.is_zero()) {
    if (op.flag0()) {
      flags.reset();
      flags.lt = true;
    }
    perform_dst0_update(PrimitiveValue::empty(), pre.dst0(), op);
    perform_dst1_update(PrimitiveValue::empty(), op.dst1_reg_idx);
  } else {
    U256 q, r;
    u256_div_mod(pre.src0.value, pre.src1.value, &q, &r);
    if (op.flag0()) {
      bool eq = q.is_zero(), gt = r.is_zero();
      flags.reset();
      flags.eq = eq;
      flags.gt = gt;
    }
    perform_dst0_update(PrimitiveValue{q, false}, pre.dst0(), op);
    perform_dst1_update(PrimitiveValue{r, false}, op.dst1_reg_idx);
  }
}

// jump.rs:4-26
void VmState::op_jump(const DecodedOpcode&, const PreState& pre) { current.pc = (uint16_t)pre.src0.value.low_u64(); }

// context.rs:6-111
void VmState::op_context(const DecodedOpcode& op, const PreState& pre) {
  current.pc = pre.new_pc;
  uint32_t sub = op.sub();
  if (sub == ZK_CTX_SET_U128) {
    uint32_t tmp[8];
    pre.src0.value.to_limbs32(tmp);
    memcpy(context_u128_register, tmp, 16);  // src0.low_u128()
    return;
  }
  if (sub == ZK_CTX_SET_ERGS_PER_PUBDATA) {
    current_ergs_per_pubdata_byte = pre.src0.value.low_u32();
    return;
  }
  if (sub == ZK_CTX_INC_TX) {
    tx_number_in_block = (uint16_t)(tx_number_in_block + 1);
    return;
  }
  U256 value;
  switch (sub) {
    case ZK_CTX_THIS: value = address_to_u256(current.this_address); break;
    case ZK_CTX_CALLER: value = address_to_u256(current.msg_sender); break;
    case ZK_CTX_CODE_ADDRESS: value = address_to_u256(current.code_address); break;
    case ZK_CTX_META:
      // VmMetaParameters::to_u256 (EXTERNAL layout, reconstruction): ergs_per_pubdata | heap sizes | shard ids
      value = U256{{(uint64_t)current_ergs_per_pubdata_byte, 0,
                    (uint64_t)current.heap_bound | ((uint64_t)current.aux_heap_bound << 32),
                    ((uint64_t)current.this_shard_id << 56) | ((uint64_t)current.caller_shard_id << 48) | ((uint64_t)current.code_shard_id << 40)}};
      break;
    case ZK_CTX_ERGS_LEFT: value = U256::from_u64(current.ergs_remaining); break;
    case ZK_CTX_SP: value = U256::from_u64(current.sp); break;
    case ZK_CTX_GET_U128: {
      uint32_t tmp[8] = {current.context_u128_value[0], current.context_u128_value[1], current.context_u128_value[2], current.context_u128_value[3], 0, 0, 0, 0};
      value = U256::from_limbs32(tmp);
      break;
    }
    default: throw RefPanic("unreachable context variant");
  }
  perform_dst0_update(PrimitiveValue{value, false}, pre.dst0(), op);
}

// shift.rs:8-79
void VmState::op_shift(const DecodedOpcode& op, const PreState& pre) {
  current.pc = pre.new_pc;
  uint32_t sub = op.sub();
  uint32_t shift_abs = (uint8_t)pre.src1.value.low_u64();
  bool is_cyclic = sub == ZK_ROL || sub == ZK_ROR;
  bool is_right = sub == ZK_SHR || sub == ZK_ROR;
  U256 result;
  if (is_right) {
    result = u256_shr(pre.src0.value, shift_abs);
    if (is_cyclic) result = u256_or(result, u256_shl(pre.src0.value, 256u - shift_abs));
  } else {
    result = u256_shl(pre.src0.value, shift_abs);
    if (is_cyclic) result = u256_or(result, u256_shr(pre.src0.value, 256u - shift_abs));
  }
  if (op.flag0()) {
    flags.reset();
    flags.eq = result.is_zero();
  }
  perform_dst0_update(PrimitiveValue{result, false}, pre.dst0(), op);
}

// binop.rs:5-62
void VmState::op_binop(const DecodedOpcode& op, const PreState& pre) {
  current.pc = pre.new_pc;
  U256 result;
  switch (op.sub()) {
    case ZK_XOR: result = u256_xor(pre.src0.value, pre.src1.value); break;
    case ZK_AND: result = u256_and(pre.src0.value, pre.src1.value); break;
    default: result = u256_or(pre.src0.value, pre.src1.value); break;
  }
  if (op.flag0()) {
    flags.reset();
    flags.eq = result.is_zero();
  }
  perform_dst0_update(PrimitiveValue{result, false}, pre.dst0(), op);
}

// ptr.rs:6-194
void VmState::op_ptr(const DecodedOpcode& op, const PreState& pre) {
  current.pc = pre.new_pc;
  uint32_t sub = op.sub();
  if (!pre.src0.is_pointer || pre.src1.is_pointer) {  // ptr.rs:35-45,98-108,142-152
    set_shorthand_panic();
    return;
  }
  const U256& src0 = pre.src0.value;
  const U256& src1 = pre.src1.value;
  if (sub == ZK_PTR_ADD || sub == ZK_PTR_SUB) {
    if (src1.w[0] >= (uint64_t)ZK_MAX_OFFSET_FOR_ADD_SUB || src1.w[1] || src1.w[2] || src1.w[3]) {  // ptr.rs:47
      set_shorthand_panic();
      return;
    }
    FatPointer fp = FatPointer::from_u256(src0);
    uint32_t offset = src1.low_u32();
    uint64_t wide = sub == ZK_PTR_ADD ? (uint64_t)fp.offset + offset : (uint64_t)fp.offset - offset;
    if (wide > 0xFFFFFFFFull) {  // overflowing_add / overflowing_sub (ptr.rs:65-74)
      set_shorthand_panic();
      return;
    }
    fp.offset = (uint32_t)wide;
    U256 p = fp.to_u256();
    perform_dst0_update(PrimitiveValue{U256{{p.w[0], p.w[1], src0.w[2], src0.w[3]}}, true}, pre.dst0(), op);
  } else if (sub == ZK_PTR_PACK) {
    if (src1.w[0] != 0 || src1.w[1] != 0) {  // ptr.rs:110
      set_shorthand_panic();
      return;
    }
    perform_dst0_update(PrimitiveValue{U256{{src0.w[0], src0.w[1], src1.w[2], src1.w[3]}}, true}, pre.dst0(), op);
  } else {  // Shrink, ptr.rs:140-192
    FatPointer fp = FatPointer::from_u256(src0);
    uint32_t offset = src1.low_u32();
    if (fp.length < offset) {
      set_shorthand_panic();
      return;
    }
    fp.length -= offset;
    U256 p = fp.to_u256();
    perform_dst0_update(PrimitiveValue{U256{{p.w[0], p.w[1], src0.w[2], src0.w[3]}}, true}, pre.dst0(), op);
  }
}

// near_call.rs:6-68
void VmState::op_near_call(const DecodedOpcode& op, const PreState& pre) {
  flags.reset();
  uint16_t dst = op.imm_0, eh = op.imm_1;
  uint32_t ergs_passed_abi = pre.src0.value.low_u32();  // NearCallABI::from_u256
  uint32_t remaining = current.ergs_remaining;
  uint32_t passed, remaining_for_this;
  if (ergs_passed_abi == 0) {
    passed = remaining;
    remaining_for_this = 0;
  } else if (remaining < ergs_passed_abi) {
    passed = remaining;
    remaining_for_this = 0;
  } else {
    passed = ergs_passed_abi;
    remaining_for_this = remaining - ergs_passed_abi;
  }
  current.ergs_remaining = remaining_for_this;
  current.pc = pre.new_pc;
  CallStackEntry ns = current;
  ns.pc = dst;
  ns.exception_handler_location = eh;
  ns.ergs_remaining = passed;
  ns.is_local_frame = true;
  wt.far_call_bound_kind = 0;  // near calls never touch the heap bounds
  start_frame(ns);
}

// log.rs:11-330
void VmState::op_log(const DecodedOpcode& op, const PreState& pre) {
  const U256& src0 = pre.src0.value;
  const U256& src1 = pre.src1.value;
  uint32_t sub = op.sub();
  current.pc = pre.new_pc;
  bool is_first_message = op.flag0();
  uint8_t shard_id = current.this_shard_id;
  uint32_t ergs_available = current.ergs_remaining;
  bool is_rollup = shard_id == 0;
  uint32_t timestamp_for_log = ts_first_decommit_or_precompile_read();
  uint16_t tx = tx_number_in_block;
  Address address = current.this_address;

  uint32_t ergs_on_pubdata = 0;
  if (sub == ZK_LOG_SSTORE) {
    LogQuery partial{timestamp_for_log, tx, ZK_STORAGE_AUX_BYTE, shard_id, address, src0, U256::zero(), src1, true, false, false};
    uint32_t pubdata_refund = refund_for_partial_query(partial);
    uint32_t net_pubdata = 0;
    if (is_rollup) {
      REF_ASSERT(ZK_INITIAL_STORAGE_WRITE_PUBDATA_BYTES >= pubdata_refund, "refund can not be more than net cost itself");
      net_pubdata = ZK_INITIAL_STORAGE_WRITE_PUBDATA_BYTES - pubdata_refund;
    }
    ergs_on_pubdata = current_ergs_per_pubdata_byte * net_pubdata;  // wrapping, as release-mode Rust
  } else if (sub == ZK_LOG_TO_L1) {
    ergs_on_pubdata = current_ergs_per_pubdata_byte * ZK_L1_MESSAGE_PUBDATA_BYTES;
  }
  uint32_t extra_cost = sub == ZK_LOG_PRECOMPILE ? src1.low_u32() : 0;
  uint32_t total_cost = extra_cost + ergs_on_pubdata;
  bool not_enough_power = ergs_available < total_cost;
  uint32_t ergs_remaining = ergs_available - total_cost;
  if (not_enough_power) {
    current.ergs_remaining = 0;
    spent_pubdata_counter += std::min(ergs_available, ergs_on_pubdata);
  } else {
    current.ergs_remaining = ergs_remaining;
    spent_pubdata_counter += ergs_on_pubdata;
  }
  switch (sub) {
    case ZK_LOG_SLOAD: {
      REF_ASSERT(!not_enough_power, "sload without power (log.rs:164)");
      LogQuery partial{timestamp_for_log, tx, ZK_STORAGE_AUX_BYTE, shard_id, address, src0, U256::zero(), U256::zero(), false, false, is_first_message};
      LogQuery q = access_storage(partial);
      perform_dst0_update(PrimitiveValue{q.read_value, false}, pre.dst0(), op);
      break;
    }
    case ZK_LOG_SSTORE: {
      if (not_enough_power) return;
      LogQuery partial{timestamp_for_log, tx, ZK_STORAGE_AUX_BYTE, shard_id, address, src0, U256::zero(), src1, true, false, is_first_message};
      access_storage(partial);
      break;
    }
    case ZK_LOG_EVENT:
    case ZK_LOG_TO_L1: {
      if (not_enough_power) {
        REF_ASSERT(sub == ZK_LOG_TO_L1, "event without power (log.rs:223)");
        return;
      }
      uint8_t aux = sub == ZK_LOG_EVENT ? ZK_EVENT_AUX_BYTE : ZK_L1_MESSAGE_AUX_BYTE;
      LogQuery q{timestamp_for_log, tx, aux, shard_id, address, src0, U256::zero(), src1, true, false, is_first_message};
      emit_event(q);
      break;
    }
    default: {  // PrecompileCall, log.rs:252-328
      if (not_enough_power) {
        perform_dst0_update(PrimitiveValue::empty(), pre.dst0(), op);
        return;
      }
      PrecompileCallABI abi = PrecompileCallABI::from_u256(src0);
      current.ergs_remaining = ergs_remaining;
      if (abi.memory_page_to_read == 0) abi.memory_page_to_read = heap_page_from_base(current.base_memory_page);
      if (abi.memory_page_to_write == 0) abi.memory_page_to_write = heap_page_from_base(current.base_memory_page);
      LogQuery q{timestamp_for_log, tx, ZK_PRECOMPILE_AUX_BYTE, shard_id, address, abi.to_u256(), U256::zero(), U256::zero(), false, false, is_first_message};
      call_precompile(q);
      perform_dst0_update(PrimitiveValue{U256::from_u64(1), false}, pre.dst0(), op);
      break;
    }
  }
}

// far_call.rs:35-613
void VmState::op_far_call(const DecodedOpcode& op, const PreState& pre) {
  enum { EX_INPUT_IS_NOT_POINTER = 1, EX_INVALID_CODE_HASH_FORMAT = 2, EX_NOT_ENOUGH_ERGS_TO_DECOMMIT = 4, EX_NOT_ENOUGH_ERGS_TO_GROW_MEMORY = 8,
         EX_MALFORMED_ABI_QUASI_POINTER = 16, EX_CALL_IN_NOW_CONSTRUCTED_SYSTEM_CONTRACT = 32, EX_NOT_ENOUGH_ERGS_FOR_EXTRA = 64 };
  uint32_t inner_variant = op.sub();
  const U256& abi_src = pre.src0.value;
  bool abi_src_is_ptr = pre.src0.is_pointer;
  const U256& call_destination_value = pre.src1.value;
  flags.reset();
  bool is_call_shard = op.flag0(), is_static_call = op.flag1();
  uint16_t exception_handler_location = op.imm_0;
  Address called_address = u256_to_address_unchecked(call_destination_value);
  U256 called_address_as_u256 = address_to_u256(called_address);  // value & U256_TO_ADDRESS_MASK
  bool dst_is_kernel = CallStackEntry::address_is_kernel(called_address);

  // FarCallABI::from_u256 (EXTERNAL layout, reconstruction)
  FatPointer abi_ptr = FatPointer::from_u256(abi_src);
  uint32_t abi_ergs_passed = (uint32_t)abi_src.w[3];
  uint8_t fwd_byte = (uint8_t)(abi_src.w[3] >> 32);
  uint8_t abi_shard_id = (uint8_t)(abi_src.w[3] >> 40);
  bool constructor_call = ((abi_src.w[3] >> 48) & 0xFF) != 0;
  bool to_system = ((abi_src.w[3] >> 56) & 0xFF) != 0;
  uint32_t forwarding_mode = fwd_byte == ZK_FWD_FORWARD_FAT_POINTER ? ZK_FWD_FORWARD_FAT_POINTER
                             : fwd_byte == ZK_FWD_USE_AUX_HEAP      ? ZK_FWD_USE_AUX_HEAP
                                                                    : ZK_FWD_USE_HEAP;
  constructor_call = constructor_call & pre.is_kernel_mode;
  to_system = to_system & dst_is_kernel;

  Address current_address = current.this_address, current_msg_sender = current.msg_sender;
  uint32_t current_base_page = current.base_memory_page;
  uint8_t caller_shard_id = current.this_shard_id;
  uint32_t remaining_ergs = current.ergs_remaining;
  uint32_t current_context_u128[4];
  memcpy(current_context_u128, current.context_u128_value, 16);
  uint32_t timestamp_for_storage_read = ts_first_decommit_or_precompile_read();
  uint16_t tx = tx_number_in_block;
  uint8_t new_code_shard_id = is_call_shard ? abi_shard_id : caller_shard_id;
  uint8_t new_this_shard_id = inner_variant == ZK_FC_DELEGATE ? caller_shard_id : new_code_shard_id;
  uint32_t new_base_memory_page = memory_page_counter;

  U256 code_hash;
  bool map_to_trivial;
  if (new_code_shard_id != 0 && !block_properties.zkporter_is_available) {
    code_hash = U256::zero();
    map_to_trivial = true;
  } else {
    Address deployer = Address::zero();
    deployer.b[18] = (uint8_t)(ZK_DEPLOYER_SYSTEM_CONTRACT_ADDRESS >> 8);
    deployer.b[19] = (uint8_t)(ZK_DEPLOYER_SYSTEM_CONTRACT_ADDRESS & 0xFF);
    LogQuery partial{timestamp_for_storage_read, tx, ZK_STORAGE_AUX_BYTE, new_code_shard_id, deployer, called_address_as_u256,
                     U256::zero(), U256::zero(), false, false, false};
    LogQuery q = access_storage(partial);
    U256 from_storage = q.read_value;
    bool mask_into_default_aa = from_storage.is_zero() && !dst_is_kernel;
    code_hash = mask_into_default_aa ? block_properties.default_aa_code_hash : from_storage;
    map_to_trivial = false;
  }
  uint32_t memory_page_candidate_for_code = map_to_trivial ? ZK_UNMAPPED_PAGE : new_base_memory_page;

  uint32_t exceptions = 0;
  uint8_t buffer[32];
  code_hash.to_be(buffer);
  uint32_t code_length_in_words = 0;
  // VersionedHashGeneric::<ContractCodeSha256>::try_create_from_raw (far_call.rs:169-252)
  if (buffer[0] == ZK_CODE_HASH_VERSION_BYTE) {
    uint8_t marker = buffer[1];
    uint32_t len_words = ((uint32_t)buffer[2] << 8) | buffer[3];
    bool at_rest = marker == ZK_CODE_AT_REST_MARKER, constructed_now = marker == ZK_YET_CONSTRUCTED_MARKER;
    if (!(at_rest || constructed_now)) {
      exceptions |= EX_INVALID_CODE_HASH_FORMAT;
      code_hash = U256::zero();
    } else {
      uint8_t stored[32];
      memcpy(stored, buffer, 32);
      stored[1] = ZK_CODE_AT_REST_MARKER;  // serialize_to_stored
      U256 code_hash_at_storage = U256::from_be(stored);
      bool can_call_at_rest = !constructor_call && at_rest;
      bool can_call_by_constructor = constructor_call && constructed_now;
      if (can_call_at_rest || can_call_by_constructor) {
        code_hash = code_hash_at_storage;
        code_length_in_words = len_words;
      } else if (!dst_is_kernel) {
        uint8_t aa[32];
        block_properties.default_aa_code_hash.to_be(aa);
        REF_ASSERT(aa[0] == ZK_CODE_HASH_VERSION_BYTE, "default AA code hash must be always valid (far_call.rs:222)");
        REF_ASSERT(aa[1] == ZK_CODE_AT_REST_MARKER, "default AA marker is always in storage format (far_call.rs:225)");
        code_hash = block_properties.default_aa_code_hash;
        code_length_in_words = ((uint32_t)aa[2] << 8) | aa[3];
      } else {
        exceptions |= EX_CALL_IN_NOW_CONSTRUCTED_SYSTEM_CONTRACT;
        code_hash = U256::zero();
      }
    }
  } else {
    exceptions |= EX_INVALID_CODE_HASH_FORMAT;
    code_hash = U256::zero();
  }
  if (forwarding_mode == ZK_FWD_FORWARD_FAT_POINTER && !abi_src_is_ptr) exceptions |= EX_INPUT_IS_NOT_POINTER;
  bool validate_as_fresh = forwarding_mode != ZK_FWD_FORWARD_FAT_POINTER;
  uint32_t pointer_validation_exceptions = abi_ptr.validate(validate_as_fresh);
  if (pointer_validation_exceptions != 0) exceptions |= EX_MALFORMED_ABI_QUASI_POINTER;
  if (!abi_ptr.validate_as_slice()) exceptions |= EX_MALFORMED_ABI_QUASI_POINTER;
  switch (forwarding_mode) {
    case ZK_FWD_FORWARD_FAT_POINTER:
      abi_ptr.start = abi_ptr.start + abi_ptr.offset;
      abi_ptr.length = abi_ptr.length - abi_ptr.offset;
      abi_ptr.offset = 0;
      break;
    case ZK_FWD_USE_HEAP: abi_ptr.memory_page = heap_page_from_base(current_base_page); break;
    default: abi_ptr.memory_page = aux_heap_page_from_base(current_base_page); break;
  }
  if (exceptions != 0) abi_ptr = FatPointer::empty();

  uint32_t memory_growth_in_bytes = 0;
  if (forwarding_mode != ZK_FWD_FORWARD_FAT_POINTER) {
    uint32_t upper_bound = abi_ptr.start + abi_ptr.length;
    if (pointer_validation_exceptions & FatPointer::DEREF_BEYOND_HEAP_RANGE) upper_bound = 0xFFFFFFFFu;
    uint32_t& bound = forwarding_mode == ZK_FWD_USE_HEAP ? current.heap_bound : current.aux_heap_bound;
    if (upper_bound >= bound) {
      memory_growth_in_bytes = upper_bound - bound;
      bound = upper_bound;
    }
    wt.far_call_bound_kind = forwarding_mode == ZK_FWD_USE_HEAP ? 1 : 2;
  }
  uint32_t cost_of_memory_growth = memory_growth_in_bytes * ZK_MEMORY_GROWTH_ERGS_PER_BYTE;
  uint32_t remaining_ergs_after_growth;
  if (remaining_ergs >= cost_of_memory_growth) {
    remaining_ergs_after_growth = remaining_ergs - cost_of_memory_growth;
  } else {
    exceptions |= EX_NOT_ENOUGH_ERGS_TO_GROW_MEMORY;
    remaining_ergs_after_growth = 0;
  }
  uint32_t msg_value_stipend = 0;  // FORCED_ERGS_FOR_MSG_VALUE_SIMULATOR == false (far_call.rs:13,387-420)
  uint32_t remaining_ergs_of_caller_frame = remaining_ergs_after_growth - msg_value_stipend;
  uint32_t cost_of_decommittment = ZK_ERGS_PER_CODE_WORD_DECOMMITTMENT * code_length_in_words;
  uint32_t remaining_ergs_after_decommittment;
  if (remaining_ergs_of_caller_frame >= cost_of_decommittment) {
    remaining_ergs_after_decommittment = remaining_ergs_of_caller_frame - cost_of_decommittment;
  } else {
    exceptions |= EX_NOT_ENOUGH_ERGS_TO_DECOMMIT;
    remaining_ergs_after_decommittment = remaining_ergs_of_caller_frame;
  }
  uint32_t mapped_code_page;
  if (exceptions != 0) {
    set_shorthand_panic();
    mapped_code_page = ZK_UNMAPPED_PAGE;
  } else {
    DecommittmentQuery dq = decommit(code_hash, memory_page_candidate_for_code, ts_first_decommit_or_precompile_read());
    if (!dq.is_fresh) remaining_ergs_after_decommittment += cost_of_decommittment;
    mapped_code_page = dq.memory_page;
  }
  uint32_t stipend_for_callee = msg_value_stipend;

  // far_call.rs:468-487: the 63/64 rule
  uint32_t remaining_ergs_to_pass = remaining_ergs_after_decommittment;
  uint32_t max_passable = (remaining_ergs_to_pass / 64) * 63;
  uint32_t leftover = remaining_ergs_to_pass - max_passable;
  uint32_t passed_ergs, remaining_for_this;
  if (max_passable < abi_ergs_passed) {
    passed_ergs = max_passable;
    remaining_for_this = leftover;
  } else {
    passed_ergs = abi_ergs_passed;
    remaining_for_this = leftover + (max_passable - abi_ergs_passed);
  }
  passed_ergs += stipend_for_callee;
  current.ergs_remaining = remaining_for_this;
  current.pc = pre.new_pc;
  bool new_context_is_static = current.is_static | is_static_call;
  memory_page_counter += ZK_NEW_MEMORY_PAGES_PER_FAR_CALL;
  Address address_from_implicit_reg = u256_to_address_unchecked(registers[ZK_CALL_IMPLICIT_PARAMETER_REG_IDX].value);
  Address address_for_next, msg_sender_for_next;
  switch (inner_variant) {
    case ZK_FC_NORMAL: address_for_next = called_address; msg_sender_for_next = current_address; break;
    case ZK_FC_DELEGATE: address_for_next = current_address; msg_sender_for_next = current_msg_sender; break;
    default: address_for_next = called_address; msg_sender_for_next = address_from_implicit_reg; break;
  }
  CallStackEntry ns;
  memset(&ns, 0, sizeof(ns));
  ns.this_address = address_for_next;
  ns.msg_sender = msg_sender_for_next;
  ns.code_address = called_address;
  ns.base_memory_page = new_base_memory_page;
  ns.code_page = mapped_code_page;
  ns.sp = ZK_INITIAL_SP_ON_FAR_CALL;
  ns.pc = 0;
  ns.exception_handler_location = exception_handler_location;
  ns.ergs_remaining = passed_ergs;
  ns.this_shard_id = new_this_shard_id;
  ns.caller_shard_id = caller_shard_id;
  ns.code_shard_id = new_code_shard_id;
  ns.is_static = new_context_is_static;
  ns.is_local_frame = false;
  if (inner_variant == ZK_FC_DELEGATE)
    memcpy(ns.context_u128_value, current_context_u128, 16);
  else
    memcpy(ns.context_u128_value, context_u128_register, 16);
  ns.heap_bound = ZK_NEW_FRAME_MEMORY_STIPEND;
  ns.aux_heap_bound = ZK_NEW_FRAME_MEMORY_STIPEND;
  memset(context_u128_register, 0, 16);
  start_frame(ns);
  memory.start_global_frame(current_base_page, new_base_memory_page, abi_ptr);

  // far_call.rs:573-610 register ABI
  PrimitiveValue r1{abi_ptr.to_u256(), true};
  registers[ZK_CALL_IMPLICIT_CALLDATA_FAT_PTR_REGISTER] = r1;
  U256 r2 = U256::zero();
  if (constructor_call) r2.w[0] |= 1;
  if (to_system) r2.w[0] |= 2;
  registers[ZK_CALL_IMPLICIT_CONSTRUCTOR_MARKER_REGISTER] = PrimitiveValue{r2, false};
  for (uint32_t i = ZK_CALL_SYSTEM_ABI_REGISTERS_LO; i < ZK_CALL_SYSTEM_ABI_REGISTERS_HI; i++) {
    if (!to_system)
      registers[i] = PrimitiveValue::empty();
    else
      registers[i].is_pointer = false;
  }
  for (uint32_t i = ZK_CALL_RESERVED_RANGE_LO; i < ZK_CALL_RESERVED_RANGE_HI; i++) registers[i] = PrimitiveValue::empty();
  registers[ZK_CALL_IMPLICIT_PARAMETER_REG_IDX] = PrimitiveValue::empty();
  // row: the implicit register writes
  r1.value.to_limbs32(row.dst0);
  r2.to_limbs32(row.dst1);
  row.bits |= ZKB_ROWBIT_DST0_VALID | ZKB_ROWBIT_DST0_PTR | ZKB_ROWBIT_DST1_VALID;
}

// ret.rs:9-265
void VmState::op_ret(const DecodedOpcode& op, const PreState& pre) {
  uint32_t inner_variant = op.sub();
  flags.reset();
  U256 src0 = pre.src0.value;
  bool src0_is_ptr = pre.src0.is_pointer;
  if (inner_variant == ZK_RET_PANIC) {
    src0 = U256::zero();
    src0_is_ptr = false;
  }
  // RetABI::from_u256
  FatPointer ptr = FatPointer::from_u256(src0);
  uint8_t fwd_byte = (uint8_t)(src0.w[3] >> 32);
  uint32_t page_forwarding_mode = fwd_byte == ZK_FWD_FORWARD_FAT_POINTER ? ZK_FWD_FORWARD_FAT_POINTER
                                  : fwd_byte == ZK_FWD_USE_AUX_HEAP      ? ZK_FWD_USE_AUX_HEAP
                                                                         : ZK_FWD_USE_HEAP;
  bool is_to_label = op.flag0();
  uint16_t label_pc = op.imm_0;
  uint32_t pointer_validation_exceptions = 0;
  if (!current.is_local_frame) {
    if (page_forwarding_mode == ZK_FWD_FORWARD_FAT_POINTER) {
      if (!src0_is_ptr) inner_variant = ZK_RET_PANIC;
      if (ptr.memory_page < current.base_memory_page) inner_variant = ZK_RET_PANIC;
    }
    bool validate_as_fresh = page_forwarding_mode != ZK_FWD_FORWARD_FAT_POINTER;
    pointer_validation_exceptions = ptr.validate(validate_as_fresh);
    if (pointer_validation_exceptions != 0) inner_variant = ZK_RET_PANIC;
    if (!ptr.validate_as_slice()) inner_variant = ZK_RET_PANIC;
    if (inner_variant == ZK_RET_PANIC) ptr = FatPointer::empty();
  }
  uint32_t ergs_remaining = current.ergs_remaining;
  bool has_returndata_ptr = !current.is_local_frame;
  if (has_returndata_ptr) {
    if (inner_variant == ZK_RET_OK || inner_variant == ZK_RET_REVERT) {
      switch (page_forwarding_mode) {
        case ZK_FWD_FORWARD_FAT_POINTER:
          ptr.start = ptr.start + ptr.offset;
          ptr.length = ptr.length - ptr.offset;
          ptr.offset = 0;
          break;
        case ZK_FWD_USE_HEAP: ptr.memory_page = heap_page_from_base(current.base_memory_page); break;
        default: ptr.memory_page = aux_heap_page_from_base(current.base_memory_page); break;
      }
    }
    uint32_t memory_growth_in_bytes = 0;
    if (page_forwarding_mode != ZK_FWD_FORWARD_FAT_POINTER) {
      uint32_t upper_bound = ptr.start + ptr.length;
      if (pointer_validation_exceptions & FatPointer::DEREF_BEYOND_HEAP_RANGE) upper_bound = 0xFFFFFFFFu;
      uint32_t bound = page_forwarding_mode == ZK_FWD_USE_HEAP ? current.heap_bound : current.aux_heap_bound;
      if (upper_bound >= bound) memory_growth_in_bytes = upper_bound - bound;
    }
    uint32_t cost = memory_growth_in_bytes * ZK_MEMORY_GROWTH_ERGS_PER_BYTE;
    if (ergs_remaining >= cost) {
      ergs_remaining -= cost;
    } else {
      ergs_remaining = 0;
      inner_variant = ZK_RET_PANIC;
      ptr = FatPointer::empty();
    }
  }
  bool panicked = inner_variant == ZK_RET_REVERT || inner_variant == ZK_RET_PANIC;
  CallStackEntry finished = finish_frame(panicked);
  is_to_label = is_to_label & finished.is_local_frame;
  if (!finished.is_local_frame) {
    memory.finish_global_frame(finished.base_memory_page, ptr);
    PrimitiveValue r1{ptr.to_u256(), true};
    registers[ZK_RET_IMPLICIT_RETURNDATA_PARAMS_REGISTER] = r1;
    for (uint32_t i = 1; i < ZK_REGISTERS_COUNT; i++) registers[i] = PrimitiveValue::empty();  // ret.rs:218-233
    memset(context_u128_register, 0, 16);
    r1.value.to_limbs32(row.dst0);
    row.bits |= ZKB_ROWBIT_DST0_VALID | ZKB_ROWBIT_DST0_PTR;
  }
  current.ergs_remaining += ergs_remaining;  // ret.rs:243
  if (is_to_label)
    current.pc = label_pc;
  else if (panicked)
    current.pc = finished.exception_handler_location;
  if (finished.is_local_frame) {
    REF_ASSERT(finished.heap_bound >= current.heap_bound && finished.aux_heap_bound >= current.aux_heap_bound, "heap bound monotonicity (ret.rs:255-256)");
    current.heap_bound = finished.heap_bound;
    current.aux_heap_bound = finished.aux_heap_bound;
  }
  if (inner_variant == ZK_RET_PANIC) flags.lt = true;
}

// uma.rs:26-425
void VmState::op_uma(const DecodedOpcode& op, const PreState& pre) {
  enum { EX_INPUT_IS_NOT_POINTER = 1, EX_DEREF_BEYOND_HEAP_RANGE = 2, EX_OVERFLOW_ON_INCREMENT = 4, EX_NOT_ENOUGH_ERGS_TO_GROW_MEMORY = 8 };
  uint32_t variant = op.sub();
  current.pc = pre.new_pc;
  bool increment_offset = op.flag0();
  const U256& src0_value = pre.src0.value;
  bool src0_is_ptr = pre.src0.is_pointer;
  const U256& src1 = pre.src1.value;
  FatPointer fat_ptr = FatPointer::from_u256(src0_value);
  uint32_t exceptions = 0;
  bool skip_fat_ptr_oob = false, skip_deref = false;
  bool is_ptr_read = variant == ZK_UMA_PTR_READ;
  if (is_ptr_read && !src0_is_ptr) exceptions |= EX_INPUT_IS_NOT_POINTER;
  uint8_t memory_type;
  bool is_heap = variant == ZK_UMA_HEAP_READ || variant == ZK_UMA_HEAP_WRITE;
  bool is_aux = variant == ZK_UMA_AUX_READ || variant == ZK_UMA_AUX_WRITE;
  if (is_heap) {
    fat_ptr.memory_page = heap_page_from_base(current.base_memory_page);
    memory_type = ZK_MEM_HEAP;
  } else if (is_aux) {
    fat_ptr.memory_page = aux_heap_page_from_base(current.base_memory_page);
    memory_type = ZK_MEM_AUX_HEAP;
  } else {
    memory_type = ZK_MEM_FAT_PTR;
  }
  uint32_t src_offset;
  if (is_ptr_read) {
    if (!fat_ptr.validate_in_bounds()) skip_fat_ptr_oob = true;
    src_offset = fat_ptr.start + fat_ptr.offset;
  } else {
    // src0.value > MAX_OFFSET_TO_DEREF (uma.rs:127)
    if (src0_value.w[1] || src0_value.w[2] || src0_value.w[3] || src0_value.w[0] > (uint64_t)ZK_MAX_OFFSET_TO_DEREF) {
      exceptions |= EX_DEREF_BEYOND_HEAP_RANGE;
      skip_deref = true;
    }
    src_offset = fat_ptr.offset;
  }
  uint32_t incremented_offset = fat_ptr.offset + 32;
  bool increment_offset_of = incremented_offset < fat_ptr.offset;
  if (increment_offset_of) {
    exceptions |= EX_OVERFLOW_ON_INCREMENT;
    if (!is_ptr_read) REF_ASSERT(exceptions & EX_DEREF_BEYOND_HEAP_RANGE, "uma.rs:145 sanity");
  }
  uint32_t memory_growth_in_bytes = 0;
  if (!is_ptr_read) {
    uint32_t& bound = is_heap ? current.heap_bound : current.aux_heap_bound;
    uint32_t upper_bound = incremented_offset;
    if (upper_bound >= bound) {
      memory_growth_in_bytes = upper_bound - bound;
      bound = upper_bound;
    }
  }
  uint32_t cost_of_memory_growth = memory_growth_in_bytes * ZK_MEMORY_GROWTH_ERGS_PER_BYTE;
  if (exceptions & EX_DEREF_BEYOND_HEAP_RANGE) cost_of_memory_growth = 0xFFFFFFFFu;
  if (current.ergs_remaining < cost_of_memory_growth) {
    current.ergs_remaining = 0;
    exceptions |= EX_NOT_ENOUGH_ERGS_TO_GROW_MEMORY;
  } else {
    current.ergs_remaining -= cost_of_memory_growth;
  }
  bool set_panic = exceptions != 0;
  bool skip_memory_access = skip_fat_ptr_oob || skip_deref || set_panic;
  uint32_t word_0 = src_offset / 32;
  uint32_t word_1 = word_0 + 1;
  uint32_t unalignment = src_offset % 32;
  uint32_t word_0_lowest_bytes = 32 - unalignment;
  uint32_t word_1_highest_bytes = unalignment;
  bool is_unaligned = unalignment != 0;
  MemoryLocation loc0{memory_type, fat_ptr.memory_page, word_0}, loc1{memory_type, fat_ptr.memory_page, word_1};
  uint32_t t_read = ts_read(), t_write = ts_dst_write();
  U256 w0 = U256::zero(), w1 = U256::zero();
  if (!skip_memory_access) w0 = read_memory(loc0, t_read).value;
  if (is_unaligned && !skip_memory_access) w1 = read_memory(loc1, t_read).value;

  if (variant == ZK_UMA_HEAP_READ || variant == ZK_UMA_AUX_READ || variant == ZK_UMA_PTR_READ) {
    U256 result = u256_or(u256_shl(w0, unalignment * 8), u256_shr(w1, (32 - unalignment) * 8));
    if (variant == ZK_UMA_PTR_READ) {
      uint32_t bytes_beyond = incremented_offset - fat_ptr.length;
      bool uf = incremented_offset < fat_ptr.length;
      if (uf || skip_memory_access) bytes_beyond = 0;
      bytes_beyond %= 32;
      result = u256_shl(u256_shr(result, bytes_beyond * 8), bytes_beyond * 8);
    }
    if (!set_panic) {
      perform_dst0_update(PrimitiveValue{result, false}, pre.dst0(), op);
      if (increment_offset) {
        U256 updated = src0_value;
        updated.w[0] = (updated.w[0] & 0xFFFFFFFF00000000ull) + (uint64_t)incremented_offset;
        perform_dst1_update(PrimitiveValue{updated, src0_is_ptr}, op.dst1_reg_idx);
      }
    } else {
      set_shorthand_panic();
    }
  } else {
    U256 new_w0 = u256_shl(u256_shr(w0, word_0_lowest_bytes * 8), word_0_lowest_bytes * 8);
    new_w0 = u256_or(new_w0, u256_shr(src1, unalignment * 8));
    U256 new_w1 = u256_shr(u256_shl(w1, word_1_highest_bytes * 8), word_1_highest_bytes * 8);
    new_w1 = u256_or(new_w1, u256_shl(src1, (32 - word_1_highest_bytes) * 8));
    if (!skip_memory_access) write_memory(loc0, t_write, PrimitiveValue{new_w0, false});
    if (is_unaligned && !skip_memory_access) write_memory(loc1, t_write, PrimitiveValue{new_w1, false});
    if (!set_panic) {
      if (increment_offset) {
        U256 updated = src0_value;
        updated.w[0] = (updated.w[0] & 0xFFFFFFFF00000000ull) + (uint64_t)incremented_offset;
        perform_dst0_update(PrimitiveValue{updated, false}, pre.dst0(), op);
      }
    } else {
      set_shorthand_panic();
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// batch object + C API (mirrors include/zkb.h one-to-one with the prefix orc_)
// ---------------------------------------------------------------------------------------------------
struct OrcBatch {
  ZkbConfig cfg;
  KnownHashes known_hashes;
  BlockProperties block_properties;
  std::vector<std::unique_ptr<VmState>> vms;
  double last_run_ms = 0;
  int last_threads = 1;
};

static thread_local std::string g_err;

static void run_vm(VmState& vm, uint32_t max_cycles) {
  if (vm.status != ZKB_VM_RUNNING) return;
  uint32_t n = 0;
  try {
    while (!vm.execution_has_ended()) {
      if (max_cycles && n >= max_cycles) return;
      vm.cycle();
      n++;
    }
    vm.status = ZKB_VM_ENDED;
  } catch (const UnknownCodeHash&) {
    vm.status = ZKB_VM_UNKNOWN_CODE_HASH;
  } catch (const Unsupported&) {
    vm.status = ZKB_VM_UNSUPPORTED;
  } catch (const RefPanic& e) {
    vm.status = ZKB_VM_REFERENCE_PANIC;
    if (getenv("ORC_DEBUG")) fprintf(stderr, "oracle: reference panic: %s\n", e.what());
  }
}

static VmState* new_vm(OrcBatch* b) {
  VmState* vm = new VmState();
  vm->decommitter.known_hashes = &b->known_hashes;
  vm->block_properties = b->block_properties;
  vm->wt.enabled = b->cfg.witness_mode != 0;
  vm->storage.warm_write_refund_bytes = b->cfg.reserved[1];
  return vm;
}

static Address addr_from(const uint8_t* p) {
  Address a;
  memcpy(a.b, p, 20);
  return a;
}

}  // namespace

extern "C" {

int32_t orc_create(const ZkbConfig* cfg, OrcBatch** out) {
  if (!cfg || !out || cfg->n_vms == 0) return ZKB_ERR_INVALID_ARGUMENT;
  if (cfg->reserved[1] > ZK_INITIAL_STORAGE_WRITE_PUBDATA_BYTES) return ZKB_ERR_INVALID_ARGUMENT;  // log.rs:110
  OrcBatch* b = new OrcBatch();
  b->cfg = *cfg;
  b->vms.resize(cfg->n_vms);
  for (auto& v : b->vms) v.reset(new_vm(b));
  *out = b;
  return ZKB_OK;
}

int32_t orc_destroy(OrcBatch* b) {
  delete b;
  return ZKB_OK;
}

const char* orc_last_error(void) { return g_err.c_str(); }

int32_t orc_reset(OrcBatch* b) {
  for (auto& v : b->vms) v.reset(new_vm(b));
  return ZKB_OK;
}

int32_t orc_load_bytecode(OrcBatch* b, const uint8_t hash_be[32], const uint8_t* words_be, uint32_t n_words) {
  U256 h = U256::from_be(hash_be);
  if (b->known_hashes.count(h)) return ZKB_ERR_INVALID_ARGUMENT;  // decommitter.rs:25 assert
  auto v = std::make_shared<std::vector<U256>>(n_words);
  for (uint32_t i = 0; i < n_words; i++) (*v)[i] = U256::from_be(words_be + 32 * i);
  b->known_hashes[h] = v;
  return ZKB_OK;
}

// SimpleDecommitter.known_hashes lookup (decommitter.rs:10-13): the words a fresh decommit hands to the tracer (:81-97)
int32_t orc_read_bytecode(OrcBatch* b, const uint8_t hash_be[32], uint8_t* words_be_out, uint32_t max_words, uint32_t* n_words_out) {
  if (!b || !hash_be) return ZKB_ERR_INVALID_ARGUMENT;
  auto it = b->known_hashes.find(U256::from_be(hash_be));
  if (it == b->known_hashes.end()) return ZKB_ERR_UNKNOWN_BYTECODE;
  const std::vector<U256>& w = *it->second;
  if (n_words_out) *n_words_out = (uint32_t)w.size();
  for (uint32_t i = 0; i < std::min<uint32_t>(max_words, (uint32_t)w.size()); i++) w[i].to_be(words_be_out + 32 * (size_t)i);
  return ZKB_OK;
}

// SimpleMemory::polulate_bootloaders_calldata (memory.rs:293-298)
int32_t orc_set_calldata(OrcBatch* b, uint32_t vm_lo, uint32_t vm_hi, const uint8_t* words_be, uint32_t n_words, uint32_t per_vm) {
  if (!b || vm_lo > vm_hi || vm_hi > b->cfg.n_vms || (!words_be && n_words)) return ZKB_ERR_INVALID_ARGUMENT;
  for (uint32_t v = vm_lo; v < vm_hi; v++) {
    const uint8_t* src = per_vm ? words_be + (size_t)(v - vm_lo) * n_words * 32 : words_be;
    std::vector<U256> values(n_words);
    for (uint32_t i = 0; i < n_words; i++) values[i] = U256::from_be(src + 32 * (size_t)i);
    b->vms[v]->memory.pages_with_extended_lifetime.at(ZK_BOOTLOADER_CALLDATA_PAGE) = std::move(values);
  }
  return ZKB_OK;
}

// dump_page_content(BOOTLOADER_CALLDATA_PAGE, range) (memory.rs:300-344)
int32_t orc_read_calldata(OrcBatch* b, uint32_t vm, uint32_t word_lo, uint32_t n_words, uint8_t* words_be_out) {
  if (!b || vm >= b->cfg.n_vms || (!words_be_out && n_words)) return ZKB_ERR_INVALID_ARGUMENT;
  const std::vector<U256>& page = b->vms[vm]->memory.pages_with_extended_lifetime.at(ZK_BOOTLOADER_CALLDATA_PAGE);
  for (uint32_t i = 0; i < n_words; i++) {
    const size_t at = (size_t)word_lo + i;
    (at < page.size() ? page[at] : U256()).to_be(words_be_out + 32 * (size_t)i);
  }
  return ZKB_OK;
}

int32_t orc_set_block_properties(OrcBatch* b, const uint8_t default_aa_code_hash_be[32], uint8_t zkporter_is_available) {
  b->block_properties.default_aa_code_hash = U256::from_be(default_aa_code_hash_be);
  b->block_properties.zkporter_is_available = zkporter_is_available != 0;
  for (auto& v : b->vms) v->block_properties = b->block_properties;
  return ZKB_OK;
}

int32_t orc_populate_storage(OrcBatch* b, uint32_t vm_lo, uint32_t vm_hi, const ZkbStorageInit* entries, uint32_t n, uint32_t per_vm) {
  if (vm_hi > b->cfg.n_vms || vm_lo > vm_hi) return ZKB_ERR_INVALID_ARGUMENT;
  for (uint32_t v = vm_lo; v < vm_hi; v++) {
    const ZkbStorageInit* e = per_vm ? entries + (size_t)(v - vm_lo) * n : entries;
    for (uint32_t i = 0; i < n; i++) {
      SlotKey k{e[i].shard_id, addr_from(e[i].address), U256::from_be(e[i].key_be)};
      b->vms[v]->storage.inner[k] = U256::from_be(e[i].value_be);
    }
  }
  return ZKB_OK;
}

int32_t orc_populate_code(OrcBatch* b, uint32_t vm_lo, uint32_t vm_hi, uint32_t page, const uint8_t hash_be[32]) {
  if (vm_hi > b->cfg.n_vms || vm_lo > vm_hi) return ZKB_ERR_INVALID_ARGUMENT;
  auto it = b->known_hashes.find(U256::from_be(hash_be));
  if (it == b->known_hashes.end()) return ZKB_ERR_UNKNOWN_BYTECODE;
  for (uint32_t v = vm_lo; v < vm_hi; v++) b->vms[v]->memory.code_pages[page] = it->second;
  return ZKB_OK;
}

int32_t orc_push_bootloader_context(OrcBatch* b, uint32_t vm_lo, uint32_t vm_hi, const ZkbFrame* f) {
  if (vm_hi > b->cfg.n_vms || vm_lo > vm_hi) return ZKB_ERR_INVALID_ARGUMENT;
  CallStackEntry e;
  memset(&e, 0, sizeof(e));
  e.this_address = addr_from(f->this_address);
  e.msg_sender = addr_from(f->msg_sender);
  e.code_address = addr_from(f->code_address);
  e.base_memory_page = f->base_memory_page;
  e.code_page = f->code_page;
  e.sp = f->sp;
  e.pc = f->pc;
  e.exception_handler_location = f->exception_handler_location;
  e.ergs_remaining = f->ergs_remaining;
  e.this_shard_id = f->this_shard_id;
  e.caller_shard_id = f->caller_shard_id;
  e.code_shard_id = f->code_shard_id;
  e.is_static = f->is_static;
  e.is_local_frame = f->is_local_frame;
  memcpy(e.context_u128_value, f->context_u128_value, 16);
  e.heap_bound = f->heap_bound;
  e.aux_heap_bound = f->aux_heap_bound;
  for (uint32_t v = vm_lo; v < vm_hi; v++) {
    try {
      b->vms[v]->push_bootloader_context(e);
    } catch (const RefPanic&) {
      return ZKB_ERR_INVALID_ARGUMENT;
    }
  }
  return ZKB_OK;
}

int32_t orc_populate_heap(OrcBatch* b, uint32_t vm_lo, uint32_t vm_hi, const uint8_t* bytes, uint32_t n_bytes, uint32_t per_vm) {
  if (vm_hi > b->cfg.n_vms || vm_lo > vm_hi) return ZKB_ERR_INVALID_ARGUMENT;
  uint32_t n_words = (n_bytes + 31) / 32;
  for (uint32_t v = vm_lo; v < vm_hi; v++) {
    const uint8_t* src = per_vm ? bytes + (size_t)(v - vm_lo) * n_bytes : bytes;
    std::vector<U256> words(n_words);
    for (uint32_t i = 0; i < n_words; i++) {
      uint8_t buf[32];
      memset(buf, 0, 32);
      uint32_t take = std::min<uint32_t>(32, n_bytes - 32 * i);
      memcpy(buf, src + 32 * i, take);
      words[i] = U256::from_be(buf);
    }
    b->vms[v]->memory.heaps.back().first.content = std::move(words);  // memory.rs:287-291
  }
  return ZKB_OK;
}

int32_t orc_set_register(OrcBatch* b, uint32_t vm_lo, uint32_t vm_hi, uint32_t reg, const uint8_t* value_be, uint8_t is_pointer, uint32_t per_vm) {
  if (vm_hi > b->cfg.n_vms || vm_lo > vm_hi || reg >= ZK_REGISTERS_COUNT) return ZKB_ERR_INVALID_ARGUMENT;
  for (uint32_t v = vm_lo; v < vm_hi; v++) {
    const uint8_t* src = per_vm ? value_be + (size_t)(v - vm_lo) * 32 : value_be;
    b->vms[v]->registers[reg] = PrimitiveValue{U256::from_be(src), is_pointer != 0};
  }
  return ZKB_OK;
}

int32_t orc_set_local_field(OrcBatch* b, uint32_t vm_lo, uint32_t vm_hi, uint32_t field, uint32_t value) {
  if (vm_hi > b->cfg.n_vms || vm_lo > vm_hi) return ZKB_ERR_INVALID_ARGUMENT;
  for (uint32_t v = vm_lo; v < vm_hi; v++) {
    VmState& vm = *b->vms[v];
    switch (field) {
      case ZKB_FIELD_MEMORY_PAGE_COUNTER: vm.memory_page_counter = value; break;
      case ZKB_FIELD_ERGS_PER_PUBDATA: vm.current_ergs_per_pubdata_byte = value; break;
      case ZKB_FIELD_TX_NUMBER: vm.tx_number_in_block = (uint16_t)value; break;
      case ZKB_FIELD_TIMESTAMP: vm.timestamp = value; break;
      default: return ZKB_ERR_INVALID_ARGUMENT;
    }
  }
  return ZKB_OK;
}

// n_threads == 0 => std::thread::hardware_concurrency()
int32_t orc_run_threads(OrcBatch* b, uint32_t max_cycles_per_vm, uint32_t n_threads) {
  if (n_threads == 0) n_threads = std::max(1u, std::thread::hardware_concurrency());
  n_threads = std::min<uint32_t>(n_threads, b->cfg.n_vms);
  b->last_threads = (int)n_threads;
  auto t0 = std::chrono::steady_clock::now();
  if (n_threads <= 1) {
    for (auto& v : b->vms) run_vm(*v, max_cycles_per_vm);
  } else {
    std::atomic<uint32_t> next{0};
    std::vector<std::thread> ts;
    for (uint32_t t = 0; t < n_threads; t++)
      ts.emplace_back([&]() {
        while (true) {
          uint32_t lo = next.fetch_add(16);
          if (lo >= b->cfg.n_vms) break;
          uint32_t hi = std::min(b->cfg.n_vms, lo + 16);
          for (uint32_t v = lo; v < hi; v++) run_vm(*b->vms[v], max_cycles_per_vm);
        }
      });
    for (auto& t : ts) t.join();
  }
  b->last_run_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  return ZKB_OK;
}

int32_t orc_run(OrcBatch* b, uint32_t max_cycles_per_vm, void* /*stream*/) { return orc_run_threads(b, max_cycles_per_vm, 1); }
int32_t orc_sync(OrcBatch*) { return ZKB_OK; }

int32_t orc_last_run_ms(OrcBatch* b, float* ms, uint32_t* n_launches) {
  if (ms) *ms = (float)b->last_run_ms;
  if (n_launches) *n_launches = 0;
  return ZKB_OK;
}

int32_t orc_vm_status(OrcBatch* b, uint32_t vm_lo, uint32_t vm_hi, ZkbVmStatus* out) {
  if (vm_hi > b->cfg.n_vms || vm_lo > vm_hi) return ZKB_ERR_INVALID_ARGUMENT;
  for (uint32_t v = vm_lo; v < vm_hi; v++) {
    out[v - vm_lo].code = b->vms[v]->status;
    out[v - vm_lo].cycles = b->vms[v]->monotonic_cycle_counter;
  }
  return ZKB_OK;
}

int32_t orc_read_local_state(OrcBatch* b, uint32_t vm, ZkbLocalState* out) {
  if (vm >= b->cfg.n_vms) return ZKB_ERR_INVALID_ARGUMENT;
  const VmState& s = *b->vms[vm];
  memset(out, 0, sizeof(*out));
  s.previous_code_word.to_limbs32(out->previous_code_word);
  out->previous_code_memory_page = s.previous_code_memory_page;
  for (int i = 0; i < ZK_REGISTERS_COUNT; i++) {
    s.registers[i].value.to_limbs32(out->registers[i]);
    if (s.registers[i].is_pointer) out->register_is_pointer |= (uint16_t)(1u << i);
  }
  out->flags = s.flags.bits();
  out->pending_exception = s.pending_exception;
  out->timestamp = s.timestamp;
  out->monotonic_cycle_counter = s.monotonic_cycle_counter;
  out->spent_pubdata_counter = s.spent_pubdata_counter;
  out->memory_page_counter = s.memory_page_counter;
  out->absolute_execution_step = s.absolute_execution_step;
  out->current_ergs_per_pubdata_byte = s.current_ergs_per_pubdata_byte;
  out->tx_number_in_block = s.tx_number_in_block;
  out->previous_super_pc = s.previous_super_pc;
  memcpy(out->context_u128_register, s.context_u128_register, 16);
  out->callstack_depth = (uint32_t)s.inner.size();
  const CallStackEntry& e = s.current;
  ZkbFrame& f = out->current_frame;
  memcpy(f.this_address, e.this_address.b, 20);
  memcpy(f.msg_sender, e.msg_sender.b, 20);
  memcpy(f.code_address, e.code_address.b, 20);
  f.base_memory_page = e.base_memory_page;
  f.code_page = e.code_page;
  f.sp = e.sp;
  f.pc = e.pc;
  f.exception_handler_location = e.exception_handler_location;
  f.ergs_remaining = e.ergs_remaining;
  f.this_shard_id = e.this_shard_id;
  f.caller_shard_id = e.caller_shard_id;
  f.code_shard_id = e.code_shard_id;
  f.is_static = e.is_static;
  f.is_local_frame = e.is_local_frame;
  memcpy(f.context_u128_value, e.context_u128_value, 16);
  f.heap_bound = e.heap_bound;
  f.aux_heap_bound = e.aux_heap_bound;
  return ZKB_OK;
}

static const uint32_t REC_BYTES[ZKB_N_STREAMS] = {ZKB_ROW_BYTES, ZKB_MEM_BYTES, ZKB_LOG_BYTES, ZKB_DECOMMIT_BYTES, ZKB_FRAME_BYTES, ZKB_REFUND_BYTES};

int32_t orc_stream_counts(OrcBatch* b, uint32_t kind, uint32_t vm_lo, uint32_t vm_hi, uint32_t* counts_out) {
  if (vm_hi > b->cfg.n_vms || vm_lo > vm_hi || kind >= ZKB_N_STREAMS) return ZKB_ERR_INVALID_ARGUMENT;
  for (uint32_t v = vm_lo; v < vm_hi; v++) counts_out[v - vm_lo] = (uint32_t)b->vms[v]->wt.counts[kind];
  return ZKB_OK;
}

int32_t orc_totals(OrcBatch* b, uint64_t* total_cycles, uint64_t stream_bytes[ZKB_N_STREAMS]) {
  uint64_t cyc = 0;
  for (int k = 0; k < ZKB_N_STREAMS; k++) stream_bytes[k] = 0;
  for (auto& v : b->vms) {
    cyc += v->monotonic_cycle_counter;
    for (int k = 0; k < ZKB_N_STREAMS; k++) stream_bytes[k] += v->wt.counts[k] * REC_BYTES[k];
  }
  if (total_cycles) *total_cycles = cyc;
  return ZKB_OK;
}

int32_t orc_read_stream(OrcBatch* b, uint32_t vm, uint32_t kind, void* dst, uint64_t max_bytes, uint64_t* n_bytes) {
  if (vm >= b->cfg.n_vms || kind >= ZKB_N_STREAMS) return ZKB_ERR_INVALID_ARGUMENT;
  const auto& s = b->vms[vm]->wt.s[kind];
  uint64_t n = std::min<uint64_t>(max_bytes, s.size());
  if (dst && n) memcpy(dst, s.data(), n);
  if (n_bytes) *n_bytes = s.size();
  return ZKB_OK;
}

// ---- transport codec (include/zkb_codec.h): the scalar encoder over the oracle's streams = the checker of the CUDA
// encoder's blob (bit-identical layout), plus the same decoder entry points under the orc_ prefix -----------------------
static void build_encoded_blob(OrcBatch* b, std::vector<uint8_t>& blob) {
  const size_t n = b->cfg.n_vms;
  ZkbEncodedHeader h;
  memset(&h, 0, sizeof(h));
  h.magic = ZKB_CODEC_MAGIC;
  h.version = ZKB_CODEC_VERSION;
  h.n_vms = (uint32_t)n;
  h.counts_offset = sizeof(ZkbEncodedHeader);
  h.offsets_offset = h.counts_offset + n * 32;
  std::vector<uint64_t> offsets((n + 1) * ZKB_N_STREAMS, 0);
  static const uint32_t rec_bytes[ZKB_N_STREAMS] = {ZKB_ROW_BYTES, ZKB_MEM_BYTES, ZKB_LOG_BYTES, ZKB_DECOMMIT_BYTES, ZKB_FRAME_BYTES, ZKB_REFUND_BYTES};
  // sizes: rows + memory queries of a VM are coded jointly (format v2), the other streams one by one
  std::vector<uint64_t> sz((size_t)n * ZKB_N_STREAMS, 0);
  for (size_t v = 0; v < n; v++) {
    const auto& w = b->vms[v]->wt;
    zkb_codec::encode_joint(w.s[0].data(), w.s[0].size() / ZKB_ROW_BYTES, w.s[1].data(), w.s[1].size() / ZKB_MEM_BYTES, nullptr, nullptr,
                            &sz[v * ZKB_N_STREAMS + 0], &sz[v * ZKB_N_STREAMS + 1]);
    for (int k = 2; k < ZKB_N_STREAMS; k++) sz[v * ZKB_N_STREAMS + k] = zkb_codec::encode_records(k, w.s[k].data(), w.s[k].size() / rec_bytes[k], nullptr);
    for (int k = 0; k < ZKB_N_STREAMS; k++) h.raw_bytes += w.s[k].size();
  }
  for (int k = 0; k < ZKB_N_STREAMS; k++) {
    uint64_t run = 0;
    for (size_t v = 0; v < n; v++) {
      offsets[(size_t)k * (n + 1) + v] = run;
      run += sz[v * ZKB_N_STREAMS + k];
    }
    offsets[(size_t)k * (n + 1) + n] = run;
    h.payload_bytes[k] = run;
  }
  uint64_t at = h.offsets_offset + (n + 1) * ZKB_N_STREAMS * 8;
  for (int k = 0; k < ZKB_N_STREAMS; k++) {
    at = (at + 15) / 16 * 16;
    h.payload_offset[k] = at;
    at += h.payload_bytes[k];
  }
  h.total_bytes = (at + 15) / 16 * 16;
  blob.assign(h.total_bytes, 0);
  memcpy(blob.data(), &h, sizeof(h));
  uint32_t* counts = (uint32_t*)(blob.data() + h.counts_offset);
  for (size_t v = 0; v < n; v++) {
    for (int k = 0; k < ZKB_N_STREAMS; k++) counts[v * 8 + k] = (uint32_t)b->vms[v]->wt.counts[k];
    counts[v * 8 + 6] = b->vms[v]->status;
    counts[v * 8 + 7] = b->vms[v]->monotonic_cycle_counter;
  }
  memcpy(blob.data() + h.offsets_offset, offsets.data(), offsets.size() * 8);
  for (size_t v = 0; v < n; v++) {
    const auto& w = b->vms[v]->wt;
    uint64_t rb = 0, mb = 0;
    zkb_codec::encode_joint(w.s[0].data(), w.s[0].size() / ZKB_ROW_BYTES, w.s[1].data(), w.s[1].size() / ZKB_MEM_BYTES,
                            blob.data() + h.payload_offset[0] + offsets[v], blob.data() + h.payload_offset[1] + offsets[(n + 1) + v], &rb, &mb);
    for (int k = 2; k < ZKB_N_STREAMS; k++)
      zkb_codec::encode_records(k, w.s[k].data(), w.s[k].size() / rec_bytes[k], blob.data() + h.payload_offset[k] + offsets[(size_t)k * (n + 1) + v]);
  }
}

int32_t orc_fetch_encoded(OrcBatch* b, void* host_dst, uint64_t host_capacity, uint64_t* n_bytes) {
  if (!b) return ZKB_ERR_INVALID_ARGUMENT;
  std::vector<uint8_t> blob;
  build_encoded_blob(b, blob);
  if (n_bytes) *n_bytes = blob.size();
  if (!host_dst) return ZKB_OK;
  if (blob.size() > host_capacity) return ZKB_ERR_INVALID_ARGUMENT;
  memcpy(host_dst, blob.data(), blob.size());
  return ZKB_OK;
}
int32_t orc_fetch_encoded_async(OrcBatch* b, void* host_dst, uint64_t host_capacity, uint64_t* n_bytes, void*) {
  return orc_fetch_encoded(b, host_dst, host_capacity, n_bytes);
}
int32_t orc_decode_stream(const void* blob, uint64_t blob_bytes, uint32_t vm, uint32_t kind, void* dst, uint64_t max_bytes, uint64_t* n_bytes) {
  zkb_codec::EncodedView v;
  if (!v.open(blob, blob_bytes)) return ZKB_ERR_INVALID_ARGUMENT;
  const uint64_t n = v.decode(vm, kind, dst, max_bytes, /*reference_decoder=*/true);   // the word-by-word inverse: checks the library's fast decoder
  if (n == UINT64_MAX) return ZKB_ERR_INVALID_ARGUMENT;
  if (n_bytes) *n_bytes = n;
  return ZKB_OK;
}
int32_t orc_decode_counts(const void* blob, uint64_t blob_bytes, uint32_t vm, uint32_t counts_out[8]) {
  zkb_codec::EncodedView v;
  if (!v.open(blob, blob_bytes) || vm >= v.n_vms() || !counts_out) return ZKB_ERR_INVALID_ARGUMENT;
  for (int i = 0; i < 8; i++) counts_out[i] = v.counts(vm)[i];
  return ZKB_OK;
}

int32_t orc_decode_all(const void* blob, uint64_t blob_bytes, uint32_t kind, void* dst, uint64_t capacity, uint64_t* offsets_out, uint32_t) {
  zkb_codec::EncodedView v;
  if (!v.open(blob, blob_bytes) || kind >= ZKB_N_STREAMS || !offsets_out) return ZKB_ERR_INVALID_ARGUMENT;
  const uint64_t rec = (uint64_t)ZKB_CODEC_REC_WORDS[kind] * 4;
  offsets_out[0] = 0;
  for (uint32_t vm = 0; vm < v.n_vms(); vm++) offsets_out[vm + 1] = offsets_out[vm] + (uint64_t)v.counts(vm)[kind] * rec;
  if (!dst) return ZKB_OK;
  if (offsets_out[v.n_vms()] > capacity) return ZKB_ERR_INVALID_ARGUMENT;
  for (uint32_t vm = 0; vm < v.n_vms(); vm++) {
    const uint64_t len = offsets_out[vm + 1] - offsets_out[vm];
    if (v.decode(vm, kind, (uint8_t*)dst + offsets_out[vm], len, true) != len) return ZKB_ERR_INVALID_ARGUMENT;
  }
  return ZKB_OK;
}

// ---- flattened backend histories (the reference's own post-processing, restated) --------------------------------
static ZkbLogQueryRec to_rec(const LogQuery& q) {
  ZkbLogQueryRec r;
  memset(&r, 0, sizeof(r));
  r.timestamp = q.timestamp;
  r.tx_number_in_block = q.tx_number_in_block;
  r.aux_byte = q.aux_byte;
  r.shard_id = q.shard_id;
  memcpy(r.address, q.address.b, 20);
  r.rw_flag = q.rw_flag;
  r.rollback = q.rollback;
  r.is_service = q.is_service;
  q.key.to_limbs32(r.key);
  q.read_value.to_limbs32(r.read_value);
  q.written_value.to_limbs32(r.written_value);
  return r;
}

// kind 0: InMemoryStorage::flatten_and_net_history().0 (storage.rs:34-76); kind 1: InMemoryEventSink::flatten().0;
// kinds 2 / 3: the net events / L1 messages of flatten() (event_sink.rs:66-131: insert on forward, remove on rollback,
// sort by timestamp), as the LogQuery each EventMessage is projected from.  false = the VM has not ended
// (the reference asserts frames_stack.len() == 1).
static bool flat_of(const VmState& vm, uint32_t kind, std::vector<ZkbLogQueryRec>* out) {
  out->clear();
  if (vm.storage.frames_stack.size() != 1 || vm.event_sink.frames_stack.size() != 1) return false;
  if (kind == 0) {
    for (const LogQuery& q : vm.storage.frames_stack[0].forward) out->push_back(to_rec(q));
    return true;
  }
  const std::vector<LogQuery>& forward = vm.event_sink.frames_stack[0].forward;
  if (kind == 1) {
    for (const LogQuery& q : forward) out->push_back(to_rec(q));
    return true;
  }
  std::map<uint32_t, LogQuery> tmp;
  for (const LogQuery& q : forward) {
    auto it = tmp.find(q.timestamp);
    if (it != tmp.end()) {
      REF_ASSERT(q.rollback, "event flatten: second query with the same timestamp must be a rollback (event_sink.rs:87)");
      tmp.erase(it);
    } else {
      REF_ASSERT(!q.rollback, "event flatten: rollback without a forward query (event_sink.rs:90)");
      tmp[q.timestamp] = q;
    }
  }
  for (const auto& kv : tmp) {
    bool is_event = kv.second.aux_byte == ZK_EVENT_AUX_BYTE;
    if ((kind == 2) == is_event) out->push_back(to_rec(kv.second));
  }
  return true;
}

int32_t orc_flatten_logs(OrcBatch*, void*) { return ZKB_OK; }  // the backends already hold their histories

int32_t orc_flat_counts(OrcBatch* b, uint32_t kind, uint32_t vm_lo, uint32_t vm_hi, uint32_t* counts_out, uint32_t* status_out) {
  if (vm_lo > vm_hi || vm_hi > b->cfg.n_vms || kind >= 4) return ZKB_ERR_INVALID_ARGUMENT;
  std::vector<ZkbLogQueryRec> v;
  for (uint32_t i = vm_lo; i < vm_hi; i++) {
    bool ok = flat_of(*b->vms[i], kind, &v);
    if (counts_out) counts_out[i - vm_lo] = (uint32_t)v.size();
    if (status_out) status_out[i - vm_lo] = ok ? 0 : 1;
  }
  return ZKB_OK;
}

int32_t orc_read_flat(OrcBatch* b, uint32_t vm, uint32_t kind, void* dst, uint64_t max_bytes, uint64_t* n_bytes) {
  if (vm >= b->cfg.n_vms || kind >= 4) return ZKB_ERR_INVALID_ARGUMENT;
  std::vector<ZkbLogQueryRec> v;
  flat_of(*b->vms[vm], kind, &v);
  uint64_t total = v.size() * sizeof(ZkbLogQueryRec);
  uint64_t n = std::min<uint64_t>(max_bytes, total);
  if (dst && n) memcpy(dst, v.data(), n);
  if (n_bytes) *n_bytes = total;
  return ZKB_OK;
}

// flatten_and_net_history().1 (storage.rs:50-73) of every VM: the storage history filed per slot, history order kept
// inside a slot.  The reference returns a HashMap (no slot order); the wire order here is the one zkb_net_storage_history
// documents -- slots by the 44-bit slot hash, colliding slots by first appearance -- so that the GPU output can be
// compared byte for byte.  The hash is restated (era_zk_evm_b200/csrc/logsort.cuh slot_hash64).
static uint64_t orc_slot_hash64(const ZkbLogQueryRec& r) {
  uint32_t aw[5];
  memcpy(aw, r.address, 20);
  uint64_t h = 0x9E3779B97F4A7C15ull ^ r.shard_id;
  for (int i = 0; i < 5; i++) {
    h = (h ^ aw[i]) * 0xFF51AFD7ED558CCDull;
    h ^= h >> 32;
  }
  for (int i = 0; i < 8; i++) {
    h = (h ^ r.key[i]) * 0xFF51AFD7ED558CCDull;
    h ^= h >> 32;
  }
  return h;
}
static bool same_slot(const ZkbLogQueryRec& a, const ZkbLogQueryRec& b) {
  return a.shard_id == b.shard_id && memcmp(a.address, b.address, 20) == 0 && memcmp(a.key, b.key, 32) == 0;
}
int32_t orc_net_storage_history(OrcBatch* b, void* host_sorted_out, uint64_t host_capacity, uint8_t* host_boundary_out, uint64_t* offsets_out,
                                uint64_t* n_slots_out, void*) {
  if (!b || !offsets_out) return ZKB_ERR_INVALID_ARGUMENT;
  const size_t n = b->cfg.n_vms;
  std::vector<std::vector<ZkbLogQueryRec>> per_vm(n);
  offsets_out[0] = 0;
  for (size_t v = 0; v < n; v++) {
    if (b->vms[v]->status == ZKB_VM_ENDED) flat_of(*b->vms[v], 0, &per_vm[v]);
    offsets_out[v + 1] = offsets_out[v] + per_vm[v].size();
  }
  if (n_slots_out) *n_slots_out = 0;
  if (!host_sorted_out || offsets_out[n] == 0) return ZKB_OK;
  if (offsets_out[n] * sizeof(ZkbLogQueryRec) > host_capacity || !host_boundary_out) return ZKB_ERR_INVALID_ARGUMENT;
  ZkbLogQueryRec* out = (ZkbLogQueryRec*)host_sorted_out;
  uint64_t slots = 0;
  for (size_t v = 0; v < n; v++) {
    const auto& h = per_vm[v];
    std::vector<size_t> order(h.size());
    for (size_t i = 0; i < h.size(); i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](size_t x, size_t y) { return (orc_slot_hash64(h[x]) >> 20) < (orc_slot_hash64(h[y]) >> 20); });
    for (size_t i = 0; i < h.size(); i++) {
      const uint64_t at = offsets_out[v] + i;
      out[at] = h[order[i]];
      const bool first = i == 0 || (orc_slot_hash64(h[order[i]]) >> 20) != (orc_slot_hash64(h[order[i - 1]]) >> 20) || !same_slot(h[order[i]], h[order[i - 1]]);
      host_boundary_out[at] = first ? 1 : 0;
      slots += first ? 1 : 0;
    }
  }
  if (n_slots_out) *n_slots_out = slots;
  return ZKB_OK;
}

int32_t orc_read_storage(OrcBatch* b, uint32_t vm, uint8_t shard_id, const uint8_t address[20], const uint8_t key_be[32], uint8_t value_be_out[32]) {
  if (vm >= b->cfg.n_vms) return ZKB_ERR_INVALID_ARGUMENT;
  SlotKey k{shard_id, addr_from(address), U256::from_be(key_be)};
  auto it = b->vms[vm]->storage.inner.find(k);
  U256 v = it == b->vms[vm]->storage.inner.end() ? U256::zero() : it->second;
  v.to_be(value_be_out);
  return ZKB_OK;
}

int32_t orc_read_heap(OrcBatch* b, uint32_t vm, uint32_t byte_offset, uint32_t n_bytes, uint8_t* out) {
  if (vm >= b->cfg.n_vms) return ZKB_ERR_INVALID_ARGUMENT;
  const auto& content = b->vms[vm]->memory.heaps.back().first.content;
  for (uint32_t i = 0; i < n_bytes; i++) {
    uint32_t a = byte_offset + i;
    uint8_t buf[32];
    SimpleMemory::get_or_zero(content, a / 32).to_be(buf);
    out[i] = buf[a % 32];
  }
  return ZKB_OK;
}

// ---- standalone primitives exposed for the oracle's own pinning tests (tests/test_oracle_*.py) ----
// versioned code hash (row f-4): ContractCodeSha256 layout as parsed at /root/reference/src/opcodes/execution/far_call.rs:169-252
// (version byte, marker byte, big-endian u16 length in words, sha256(code)[4..32]) = the key of
// SimpleDecommitter::populate (/root/reference/src/reference_impls/decommitter.rs:23-28)
int32_t orc_hash_bytecodes(int32_t /*device*/, const uint8_t* words_be, const uint64_t* offsets_words, uint32_t n, uint8_t marker,
                           uint8_t* hashes_be_out) {
  if (!offsets_words || (n && !hashes_be_out)) return ZKB_ERR_INVALID_ARGUMENT;
  for (uint32_t i = 0; i < n; i++) {
    if (offsets_words[i + 1] < offsets_words[i]) return ZKB_ERR_INVALID_ARGUMENT;
    const uint64_t n_words = offsets_words[i + 1] - offsets_words[i];
    if (n_words > 0xFFFFu) return ZKB_ERR_INVALID_ARGUMENT;
    uint8_t* out = hashes_be_out + 32 * (size_t)i;
    orc_hash::sha256(words_be + 32 * offsets_words[i], n_words * 32, out);
    out[0] = (uint8_t)ZK_CODE_HASH_VERSION_BYTE;
    out[1] = marker;
    out[2] = (uint8_t)(n_words >> 8);
    out[3] = (uint8_t)n_words;
  }
  return ZKB_OK;
}
int32_t orc_ingest_bytecodes(OrcBatch* b, const uint8_t* words_be, const uint64_t* offsets_words, uint32_t n, uint8_t* hashes_be_out) {
  if (!b) return ZKB_ERR_INVALID_ARGUMENT;
  int32_t rc = orc_hash_bytecodes(0, words_be, offsets_words, n, (uint8_t)ZK_CODE_AT_REST_MARKER, hashes_be_out);
  for (uint32_t i = 0; i < n && rc == ZKB_OK; i++)
    rc = orc_load_bytecode(b, hashes_be_out + 32 * (size_t)i, words_be + 32 * offsets_words[i], (uint32_t)(offsets_words[i + 1] - offsets_words[i]));
  return rc;
}

void orc_keccak256(const uint8_t* data, uint64_t len, uint8_t out[32]) { orc_hash::keccak_sponge256(data, len, 0x01, out); }
void orc_sha3_256(const uint8_t* data, uint64_t len, uint8_t out[32]) { orc_hash::keccak_sponge256(data, len, 0x06, out); }
void orc_sha256(const uint8_t* data, uint64_t len, uint8_t out[32]) { orc_hash::sha256(data, len, out); }
int32_t orc_ecrecover(const uint8_t hash_be[32], const uint8_t r_be[32], const uint8_t s_be[32], uint32_t v_odd, uint8_t address_out[20]) {
  U256 a;
  bool ok = ecrecover_address(U256::from_be(hash_be), U256::from_be(r_be), U256::from_be(s_be), v_odd != 0, &a);
  uint8_t buf[32];
  a.to_be(buf);
  memcpy(address_out, buf + 12, 20);
  if (!ok) memset(address_out, 0, 20);
  return ok ? 1 : 0;
}

// op: 0 add, 1 sub, 2 mul (out = low||high), 3 div (out = q||r), 4 shl, 5 shr ; operands as 32 BE bytes
void orc_u256_op(uint32_t op, const uint8_t a_be[32], const uint8_t b_be[32], uint8_t out0_be[32], uint8_t out1_be[32], uint8_t* flag) {
  U256 a = U256::from_be(a_be), b = U256::from_be(b_be), r0 = U256::zero(), r1 = U256::zero();
  bool of = false;
  switch (op) {
    case 0: r0 = u256_add(a, b, &of); break;
    case 1: r0 = u256_sub(a, b, &of); break;
    case 2: {
      uint64_t t[8];
      u256_full_mul(a, b, t);
      r0 = U256{{t[0], t[1], t[2], t[3]}};
      r1 = U256{{t[4], t[5], t[6], t[7]}};
      break;
    }
    case 3:
      if (!b.is_zero()) u256_div_mod(a, b, &r0, &r1);
      break;
    case 4: r0 = u256_shl(a, b.low_u32()); break;
    case 5: r0 = u256_shr(a, b.low_u32()); break;
  }
  r0.to_be(out0_be);
  r1.to_be(out1_be);
  *flag = of;
}

// the reference's live keccak test harness (src/testing/tests/precompiles/keccak256.rs:74-142) restated:
// SimpleMemory with heap page 4 + Heap(1) indirection, input written as BE words prefixed by `unalignment`
// bytes of 0xff, precompile called with byte offset/length in and a word index out; returns the output word.
int32_t orc_keccak_precompile_harness(const uint8_t* input, uint32_t len, uint32_t unalignment, uint8_t out[32], uint32_t* n_reads) {
  SimpleMemory memory;
  const uint32_t page = 4;
  memory.heaps.push_back({SimpleMemory::HeapPage{page, {}}, SimpleMemory::HeapPage{0, {}}});
  memory.page_numbers_indirections[page] = Indirection{Indirection::HEAP, 1};
  std::vector<uint8_t> padded(unalignment, 0xff);
  padded.insert(padded.end(), input, input + len);
  uint32_t n_words = (uint32_t)((padded.size() + 31) / 32);
  for (uint32_t i = 0; i < n_words; i++) {
    uint8_t buf[32];
    memset(buf, 0, 32);
    size_t take = std::min<size_t>(32, padded.size() - 32 * (size_t)i);
    memcpy(buf, padded.data() + 32 * (size_t)i, take);
    MemoryQuery q{0, ZK_MEM_HEAP, page, i, U256::from_be(buf), false, true};
    memory.execute_partial_query(q);
  }
  PrecompileCallABI abi{unalignment, len, n_words, 0, page, page, 0};
  LogQuery query;
  memset(&query, 0, sizeof(query));
  query.timestamp = 1;
  query.aux_byte = ZK_PRECOMPILE_AUX_BYTE;
  query.address.b[18] = (uint8_t)(ZK_KECCAK256_PRECOMPILE_ADDRESS >> 8);
  query.address.b[19] = (uint8_t)(ZK_KECCAK256_PRECOMPILE_ADDRESS & 0xFF);
  query.key = abi.to_u256();
  try {
    PrecompileResult r = execute_precompile(query, memory);
    if (n_reads) *n_reads = (uint32_t)r.mem_in.size();
  } catch (const std::exception&) {
    return ZKB_ERR_INVALID_ARGUMENT;
  }
  SimpleMemory::get_or_zero(memory.heaps.back().first.content, n_words).to_be(out);
  return ZKB_OK;
}

}  // extern "C"
