// zkb_host.hpp — C++ host-side mirror of the reference's interface for the hot path, over the C ABI of zkb.h.
//
// The reference is a Rust crate and there is no Rust toolchain in this image, so the host layer that a downstream
// user programs against is provided in C++ (header-only) with the reference's own names, argument meaning and error
// behaviour (INTEGRATION.md shows the equivalent Rust `-sys` binding + shim):
//
//   reference (zk_evm 1.4.1)                                              here
//   ------------------------------------------------------------------   --------------------------------------------
//   VmState::empty_state(..) x n            src/vm_state/mod.rs:188-207   GpuVmBatch::empty_state(cfg, block_properties)
//   SimpleDecommitter::populate             reference_impls/decommitter.rs:23   GpuVmBatch::populate_decommitter
//   SimpleMemory::populate_code / _heap     reference_impls/memory.rs:271-291   populate_code / populate_heap
//   InMemoryStorage::populate               src/testing/storage.rs:26-32        populate_storage
//   VmState::push_bootloader_context        src/vm_state/helpers.rs:289-316     push_bootloader_context
//   while !vm.execution_has_ended() { vm.cycle(&mut tracer)? }   cycle.rs:257   run(max_cycles) + replay(vm, tracer)
//   VmWitnessTracer (10 callbacks)          src/witness_trace/mod.rs:11-72      struct VmWitnessTracer (same 10 virtuals)
//   anyhow::Error "unknown code hash"       decommitter.rs:50-56                VmError{ZKB_VM_UNKNOWN_CODE_HASH}
//
// `replay` decodes one VM's packed streams (include/zkb_records.h) and calls the tracer in the reference's program
// order, handing it the same values the Rust callbacks receive — including the full `VmLocalState` before and after
// every cycle, which is reconstructed from the cycle rows + frame records (registers are tracked from the dst0/dst1
// write-backs and the deterministic register ABI of far_call / ret, far_call.rs:573-610, ret.rs:213-236).
#pragma once
#include <array>
#include <cstdint>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "zkb.h"
#include "zkb_codec.h"

#ifndef ZK_TABLE_QUALIFIER
#define ZK_TABLE_QUALIFIER static const
#endif
#include "../era_zk_evm_b200/csrc/isa_tables.inc"

#ifndef ZKB_HOST_PREFIX
#define ZKB_HOST_PREFIX zkb_
#endif
#define ZKB_HOST_CAT2(a, b) a##b
#define ZKB_HOST_CAT(a, b) ZKB_HOST_CAT2(a, b)
#define ZKB_FN(name) ZKB_HOST_CAT(ZKB_HOST_PREFIX, name)

// the prefixed entry points (identical to zkb.h when the prefix is zkb_; tests bind the CPU oracle with orc_)
extern "C" {
int32_t ZKB_FN(create)(const ZkbConfig*, ZkbBatch**);
int32_t ZKB_FN(destroy)(ZkbBatch*);
const char* ZKB_FN(last_error)(void);
int32_t ZKB_FN(load_bytecode)(ZkbBatch*, const uint8_t*, const uint8_t*, uint32_t);
int32_t ZKB_FN(set_block_properties)(ZkbBatch*, const uint8_t*, uint8_t);
int32_t ZKB_FN(populate_storage)(ZkbBatch*, uint32_t, uint32_t, const ZkbStorageInit*, uint32_t, uint32_t);
int32_t ZKB_FN(populate_code)(ZkbBatch*, uint32_t, uint32_t, uint32_t, const uint8_t*);
int32_t ZKB_FN(push_bootloader_context)(ZkbBatch*, uint32_t, uint32_t, const ZkbFrame*);
int32_t ZKB_FN(populate_heap)(ZkbBatch*, uint32_t, uint32_t, const uint8_t*, uint32_t, uint32_t);
int32_t ZKB_FN(set_register)(ZkbBatch*, uint32_t, uint32_t, uint32_t, const uint8_t*, uint8_t, uint32_t);
int32_t ZKB_FN(run)(ZkbBatch*, uint32_t, void*);
int32_t ZKB_FN(sync)(ZkbBatch*);
int32_t ZKB_FN(vm_status)(ZkbBatch*, uint32_t, uint32_t, ZkbVmStatus*);
int32_t ZKB_FN(read_local_state)(ZkbBatch*, uint32_t, ZkbLocalState*);
int32_t ZKB_FN(read_stream)(ZkbBatch*, uint32_t, uint32_t, void*, uint64_t, uint64_t*);
int32_t ZKB_FN(read_bytecode)(ZkbBatch*, const uint8_t*, uint8_t*, uint32_t, uint32_t*);
}

namespace zkb_host {

// ---- value types (zk_evm_abstractions::queries / aux, as used by src/witness_trace/mod.rs) -------------------------
struct U256 {
  std::array<uint64_t, 4> limbs{};  // little-endian, == ethereum_types::U256.0
  static U256 from_limbs32(const uint32_t* l) {
    U256 v;
    for (int i = 0; i < 4; i++) v.limbs[i] = (uint64_t)l[2 * i] | (uint64_t)l[2 * i + 1] << 32;
    return v;
  }
  void to_limbs32(uint32_t* l) const {
    for (int i = 0; i < 4; i++) {
      l[2 * i] = (uint32_t)limbs[i];
      l[2 * i + 1] = (uint32_t)(limbs[i] >> 32);
    }
  }
  static U256 from_be(const uint8_t* b) {
    U256 v;
    for (int i = 0; i < 4; i++)
      for (int k = 0; k < 8; k++) v.limbs[3 - i] = v.limbs[3 - i] << 8 | b[8 * i + k];
    return v;
  }
  void to_be(uint8_t* b) const {
    for (int i = 0; i < 4; i++)
      for (int k = 0; k < 8; k++) b[8 * i + k] = (uint8_t)(limbs[3 - i] >> (56 - 8 * k));
  }
  bool operator==(const U256& o) const { return limbs == o.limbs; }
  bool operator<(const U256& o) const {
    for (int i = 3; i >= 0; i--)
      if (limbs[i] != o.limbs[i]) return limbs[i] < o.limbs[i];
    return false;
  }
  bool operator!=(const U256& o) const { return !(*this == o); }
  uint32_t low_u32() const { return (uint32_t)limbs[0]; }
};

enum class MemoryType : uint8_t { Stack = 0, Heap = 1, AuxHeap = 2, FatPointer = 3, Code = 4 };
struct MemoryLocation {
  MemoryType memory_type;
  uint32_t page, index;
};
struct MemoryQuery {
  uint32_t timestamp;
  MemoryLocation location;
  U256 value;
  bool value_is_pointer, rw_flag;
};
struct LogQuery {
  uint32_t timestamp;
  uint16_t tx_number_in_block;
  uint8_t aux_byte, shard_id;
  std::array<uint8_t, 20> address;
  U256 key, read_value, written_value;
  bool rw_flag, rollback, is_service;
};
struct DecommittmentQuery {
  U256 hash;
  uint32_t timestamp, memory_page;
  uint16_t decommitted_length;
  bool is_fresh;
};
enum class RefundKind : uint32_t { None = 0, RepeatedWrite = 1 };
struct RefundType {
  RefundKind kind;
  uint32_t value;
};
// PrecompileCyclesWitness (zk_evm_abstractions, third output of execute_precompile, helpers.rs:211-221): per round of the
// precompile's state machine, the request that started it (first round only), the memory reads the round consumed and
// the writes it produced.  The rounds are a pure regrouping of the call's memory witness, so the replay rebuilds them
// on the host from the precompile-origin MemoryQueryRecs: keccak256 -- one round per 136-byte block, a round reads the
// input words it is the first to touch (<= 6), the digest write belongs to the last round; sha256 -- two reads per
// round, digest write in the last; ecrecover -- a single round of four reads and two writes.  (The external enum's
// exact field layout is not in /root/reference; this mirrors its content, see INTEGRATION.md.)
enum class PrecompileKind : uint8_t { Keccak256 = 0, Sha256 = 1, ECRecover = 2 };
struct PrecompileRoundWitness {
  bool has_new_request = false;
  LogQuery new_request{};
  std::vector<MemoryQuery> reads, writes;
};
struct PrecompileCyclesWitness {
  PrecompileKind kind = PrecompileKind::Keccak256;
  std::vector<PrecompileRoundWitness> rounds;
};
struct PrimitiveValue {
  U256 value;
  bool is_pointer = false;
};
struct Flags {
  bool overflow_or_less_than_flag = false, equality_flag = false, greater_than_flag = false;
};
// execution_stack.rs:6-24
struct CallStackEntry {
  std::array<uint8_t, 20> this_address{}, msg_sender{}, code_address{};
  uint32_t base_memory_page = 0, code_page = 0;
  uint16_t sp = 0, pc = 0, exception_handler_location = 0;
  uint32_t ergs_remaining = 0;
  uint8_t this_shard_id = 0, caller_shard_id = 0, code_shard_id = 0;
  bool is_static = false, is_local_frame = false;
  std::array<uint32_t, 4> context_u128_value{};
  uint32_t heap_bound = 0, aux_heap_bound = 0;
};
// vm_state/mod.rs:54-73 (callstack = current entry + the saved entries below it)
struct VmLocalState {
  U256 previous_code_word;
  uint32_t previous_code_memory_page = 0;
  std::array<PrimitiveValue, 15> registers{};
  Flags flags;
  uint32_t timestamp = 0, monotonic_cycle_counter = 0, spent_pubdata_counter = 0, memory_page_counter = 0;
  uint32_t absolute_execution_step = 0, current_ergs_per_pubdata_byte = 0;
  uint16_t tx_number_in_block = 0, previous_super_pc = 0;
  bool pending_exception = false;
  std::array<uint32_t, 4> context_u128_register{};
  CallStackEntry current;
  std::vector<CallStackEntry> inner;  // Callstack.inner (execution_stack.rs:27-30); depth() == inner.size()
};

// ---- the output contract: src/witness_trace/mod.rs:11-72 (defaults = DummyTracer, :75-77) -------------------------
struct VmWitnessTracer {
  virtual ~VmWitnessTracer() = default;
  virtual void start_new_execution_cycle(const VmLocalState&) {}
  virtual void end_execution_cycle(const VmLocalState&) {}
  virtual void add_memory_query(uint32_t /*monotonic_cycle_counter*/, const MemoryQuery&) {}
  virtual void record_refund_for_query(uint32_t, const LogQuery&, RefundType) {}
  virtual void add_log_query(uint32_t, const LogQuery&) {}
  virtual void add_decommittment(uint32_t, const DecommittmentQuery&, const std::vector<U256>& /*code words if fresh*/) {}
  virtual void add_precompile_call_result(uint32_t, const LogQuery&, const std::vector<MemoryQuery>& /*mem_witness_in*/,
                                          const std::vector<MemoryQuery>& /*memory_witness_out*/, const PrecompileCyclesWitness&) {}
  virtual void add_revertable_precompile_call(uint32_t, const LogQuery&) {}  // never called by the reference either
  virtual void start_new_execution_context(uint32_t, const CallStackEntry& /*previous*/, const CallStackEntry& /*new*/) {}
  virtual void finish_execution_context(uint32_t, bool /*panicked*/) {}
};

struct VmError : std::runtime_error {
  uint32_t code;
  VmError(uint32_t c, const std::string& m) : std::runtime_error(m), code(c) {}
};

struct BlockProperties {  // src/block_properties/mod.rs:4-7
  std::array<uint8_t, 32> default_aa_code_hash{};
  bool zkporter_is_available = false;
};

// ---- the batch: n x VmState + the caller's cycle loop -----------------------------------------------------------------
class GpuVmBatch {
 public:
  // = VmState::empty_state for cfg.n_vms instances (mod.rs:188-207); owns the device batch
  static GpuVmBatch empty_state(const ZkbConfig& cfg, const BlockProperties& bp) {
    GpuVmBatch b;
    b.cfg_ = cfg;
    b.check(ZKB_FN(create)(&cfg, &b.h_), "create");
    b.owned_ = true;
    b.check(ZKB_FN(set_block_properties)(b.h_, bp.default_aa_code_hash.data(), bp.zkporter_is_available), "set_block_properties");
    return b;
  }
  // wrap a batch created elsewhere (non-owning), e.g. through another language binding of the same C ABI
  static GpuVmBatch wrap(ZkbBatch* handle, uint32_t n_vms) {
    GpuVmBatch b;
    b.h_ = handle;
    b.cfg_.n_vms = n_vms;
    return b;
  }
  GpuVmBatch(GpuVmBatch&& o) noexcept { *this = std::move(o); }
  GpuVmBatch& operator=(GpuVmBatch&& o) noexcept {
    std::swap(h_, o.h_);
    std::swap(owned_, o.owned_);
    std::swap(code_cache_, o.code_cache_);
    cfg_ = o.cfg_;
    return *this;
  }
  ~GpuVmBatch() {
    if (owned_ && h_) ZKB_FN(destroy)(h_);
  }

  uint32_t n_vms() const { return cfg_.n_vms; }
  ZkbBatch* handle() const { return h_; }

  void populate_decommitter(const uint8_t hash_be[32], const uint8_t* code_words_be, uint32_t n_words) {
    check(ZKB_FN(load_bytecode)(h_, hash_be, code_words_be, n_words), "load_bytecode");
  }
  void populate_code(uint32_t page, const uint8_t hash_be[32]) { check(ZKB_FN(populate_code)(h_, 0, cfg_.n_vms, page, hash_be), "populate_code"); }
  void populate_storage(const std::vector<ZkbStorageInit>& entries, bool per_vm = false) {
    uint32_t n = per_vm ? (uint32_t)(entries.size() / cfg_.n_vms) : (uint32_t)entries.size();
    check(ZKB_FN(populate_storage)(h_, 0, cfg_.n_vms, entries.data(), n, per_vm), "populate_storage");
  }
  void populate_heap(const std::vector<uint8_t>& bytes, bool per_vm = false) {
    uint32_t n = per_vm ? (uint32_t)(bytes.size() / cfg_.n_vms) : (uint32_t)bytes.size();
    check(ZKB_FN(populate_heap)(h_, 0, cfg_.n_vms, bytes.data(), n, per_vm), "populate_heap");
  }
  void set_register(uint32_t reg, const uint8_t value_be[32], bool is_pointer = false) {
    check(ZKB_FN(set_register)(h_, 0, cfg_.n_vms, reg, value_be, is_pointer, 0), "set_register");
  }
  void push_bootloader_context(const ZkbFrame& f) { check(ZKB_FN(push_bootloader_context)(h_, 0, cfg_.n_vms, &f), "push_bootloader_context"); }

  // the caller loop: every VM advances by at most max_cycles cycles (0 = until execution_has_ended)
  void run(uint32_t max_cycles = 0, void* cuda_stream = nullptr) {
    check(ZKB_FN(run)(h_, max_cycles, cuda_stream), "run");
    check(ZKB_FN(sync)(h_), "sync");
  }
  // VmState::execution_has_ended (mod.rs:214-216) for one VM; throws the reference's only Err for that VM
  bool execution_has_ended(uint32_t vm) {
    ZkbVmStatus st{};
    check(ZKB_FN(vm_status)(h_, vm, vm + 1, &st), "vm_status");
    if (st.code == ZKB_VM_UNKNOWN_CODE_HASH) throw VmError(st.code, "Trying to decommit unknown hash (decommitter.rs:50-56)");
    if (st.code == ZKB_VM_REFERENCE_PANIC) throw VmError(st.code, "reference assert!/expect would have fired");
    if (st.code >= ZKB_VM_CAP_STREAM) throw VmError(st.code, "device capacity exceeded (ZkbConfig)");
    return st.code == ZKB_VM_ENDED;
  }

  // SimpleDecommitter.known_hashes[hash] (decommitter.rs:10-13): the words a fresh decommit hands to the tracer
  const std::vector<U256>& code_words(const U256& hash) {
    auto it = code_cache_.find(hash);
    if (it != code_cache_.end()) return it->second;
    uint8_t hb[32];
    hash.to_be(hb);
    uint32_t n = 0;
    check(ZKB_FN(read_bytecode)(h_, hb, nullptr, 0, &n), "read_bytecode");
    std::vector<uint8_t> raw((size_t)n * 32);
    if (n) check(ZKB_FN(read_bytecode)(h_, hb, raw.data(), n, &n), "read_bytecode");
    std::vector<U256> words(n);
    for (uint32_t i = 0; i < n; i++) words[i] = U256::from_be(raw.data() + 32 * (size_t)i);
    return code_cache_.emplace(hash, std::move(words)).first->second;
  }

  template <class Rec>
  std::vector<Rec> read_stream(uint32_t vm, uint32_t kind) {
    uint64_t n = 0;
    check(ZKB_FN(read_stream)(h_, vm, kind, nullptr, 0, &n), "read_stream");
    std::vector<Rec> v(n / sizeof(Rec));
    if (n) check(ZKB_FN(read_stream)(h_, vm, kind, v.data(), n, &n), "read_stream");
    return v;
  }

  // Replays VM `vm`'s recorded witness into `wt`, callback by callback, in the reference's program order
  // (SURVEY.md §8b): per cycle  start_new_execution_cycle -> memory reads (code word, src0, UMA/precompile-free reads)
  // -> refund -> log query -> precompile result -> decommitment -> frame start/finish -> memory writes (UMA stores,
  // dst0 on the stack) -> end_execution_cycle.  `initial` is the state passed to the first cycle (the state right
  // after push_bootloader_context).  Returns the final tracked VmLocalState.
  VmLocalState replay(uint32_t vm, VmWitnessTracer& wt, const VmLocalState& initial);
  // The same replay straight from an ENCODED blob (zkb_fetch_encoded*, include/zkb_codec.h): what a PCIe-bound host loop
  // receives.  Only this VM's slices of the blob are decoded (a few hundred KB), nothing is expanded up front.
  VmLocalState replay_encoded(const zkb_codec::EncodedView& blob, uint32_t vm, VmWitnessTracer& wt, const VmLocalState& initial);
  struct VmStreams {
    std::vector<ZkbCycleRow> rows;
    std::vector<ZkbMemoryQueryRec> mems;
    std::vector<ZkbLogQueryRec> logs;
    std::vector<ZkbDecommitRec> decs;
    std::vector<ZkbFrameRec> frames;
    std::vector<ZkbRefundRec> refunds;
  };
  VmLocalState replay_records(const VmStreams& s, VmWitnessTracer& wt, const VmLocalState& initial);

 private:
  GpuVmBatch() { std::memset(&cfg_, 0, sizeof(cfg_)); }
  void check(int32_t rc, const char* what) {
    if (rc != ZKB_OK) throw std::runtime_error(std::string("zkb ") + what + " failed (" + std::to_string(rc) + "): " + ZKB_FN(last_error)());
  }
  ZkbBatch* h_ = nullptr;
  bool owned_ = false;
  ZkbConfig cfg_;
  std::map<U256, std::vector<U256>> code_cache_;
};

// ---- record decoding ------------------------------------------------------------------------------------------------
inline MemoryQuery decode(const ZkbMemoryQueryRec& r) {
  return MemoryQuery{r.timestamp, MemoryLocation{(MemoryType)r.memory_type, r.page, r.index}, U256::from_limbs32(r.value), r.value_is_pointer != 0,
                     r.rw_flag != 0};
}
inline LogQuery decode(const ZkbLogQueryRec& r) {
  LogQuery q{};
  q.timestamp = r.timestamp;
  q.tx_number_in_block = r.tx_number_in_block;
  q.aux_byte = r.aux_byte;
  q.shard_id = r.shard_id;
  std::memcpy(q.address.data(), r.address, 20);
  q.key = U256::from_limbs32(r.key);
  q.read_value = U256::from_limbs32(r.read_value);
  q.written_value = U256::from_limbs32(r.written_value);
  q.rw_flag = r.rw_flag != 0;
  q.rollback = r.rollback != 0;
  q.is_service = r.is_service != 0;
  return q;
}
inline CallStackEntry decode_new_frame(const ZkbFrameRec& r) {
  CallStackEntry e;
  std::memcpy(e.this_address.data(), r.this_address, 20);
  std::memcpy(e.msg_sender.data(), r.msg_sender, 20);
  std::memcpy(e.code_address.data(), r.code_address, 20);
  e.base_memory_page = r.base_memory_page;
  e.code_page = r.code_page;
  e.sp = r.sp;
  e.pc = r.pc;
  e.exception_handler_location = r.exception_handler_location;
  e.ergs_remaining = r.ergs_remaining;
  e.this_shard_id = r.this_shard_id;
  e.caller_shard_id = r.caller_shard_id;
  e.code_shard_id = r.code_shard_id;
  e.is_static = r.is_static != 0;
  e.is_local_frame = r.is_local_frame != 0;
  std::memcpy(e.context_u128_value.data(), r.context_u128_value, 16);
  e.heap_bound = r.heap_bound;
  e.aux_heap_bound = r.aux_heap_bound;
  return e;
}
inline CallStackEntry from_frame(const ZkbFrame& f) {
  CallStackEntry e;
  std::memcpy(e.this_address.data(), f.this_address, 20);
  std::memcpy(e.msg_sender.data(), f.msg_sender, 20);
  std::memcpy(e.code_address.data(), f.code_address, 20);
  e.base_memory_page = f.base_memory_page;
  e.code_page = f.code_page;
  e.sp = f.sp;
  e.pc = f.pc;
  e.exception_handler_location = f.exception_handler_location;
  e.ergs_remaining = f.ergs_remaining;
  e.this_shard_id = f.this_shard_id;
  e.caller_shard_id = f.caller_shard_id;
  e.code_shard_id = f.code_shard_id;
  e.is_static = f.is_static != 0;
  e.is_local_frame = f.is_local_frame != 0;
  std::memcpy(e.context_u128_value.data(), f.context_u128_value, 16);
  e.heap_bound = f.heap_bound;
  e.aux_heap_bound = f.aux_heap_bound;
  return e;
}

// regroups one precompile call's memory witness into its rounds (see PrecompileCyclesWitness above)
inline PrecompileCyclesWitness precompile_rounds(const LogQuery& request, const std::vector<MemoryQuery>& in, const std::vector<MemoryQuery>& out) {
  PrecompileCyclesWitness w;
  const uint32_t addr_low = (uint32_t)request.address[18] << 8 | request.address[19];
  size_t k = 0;
  if (addr_low == ZK_ECRECOVER_PRECOMPILE_ADDRESS) {
    w.kind = PrecompileKind::ECRecover;
    w.rounds.emplace_back();
    w.rounds[0].reads = in;
    k = in.size();
  } else if (addr_low == ZK_SHA256_PRECOMPILE_ADDRESS) {
    w.kind = PrecompileKind::Sha256;
    for (; k + 2 <= in.size(); k += 2) {
      w.rounds.emplace_back();
      w.rounds.back().reads.assign(in.begin() + k, in.begin() + k + 2);
    }
  } else {
    w.kind = PrecompileKind::Keccak256;
    // PrecompileCallABI: input byte offset [0,32), byte length [32,64) of the query key (keccak256.rs:100-111)
    const uint64_t in_off = (uint32_t)request.key.limbs[0], in_len = (uint32_t)(request.key.limbs[0] >> 32), end = in_off + in_len;
    uint64_t next_word = in_off / 32;
    for (uint64_t blk = 0; blk < in_len / 136 + 1; blk++) {
      w.rounds.emplace_back();
      const uint64_t a0 = in_off + blk * 136, nb = std::min<uint64_t>(136, end - a0);
      if (nb == 0) continue;
      const uint64_t last = (a0 + nb - 1) / 32;
      for (; next_word <= last && k < in.size(); next_word++, k++) w.rounds.back().reads.push_back(in[k]);
    }
  }
  if (w.rounds.empty()) w.rounds.emplace_back();
  if (k != in.size()) throw std::runtime_error("replay: precompile read witness does not match its request");
  w.rounds.front().has_new_request = true;
  w.rounds.front().new_request = request;
  w.rounds.back().writes = out;
  return w;
}

inline VmLocalState GpuVmBatch::replay(uint32_t vm, VmWitnessTracer& wt, const VmLocalState& initial) {
  VmStreams s;
  s.rows = read_stream<ZkbCycleRow>(vm, ZKB_STREAM_ROWS);
  s.mems = read_stream<ZkbMemoryQueryRec>(vm, ZKB_STREAM_MEM);
  s.logs = read_stream<ZkbLogQueryRec>(vm, ZKB_STREAM_LOG);
  s.decs = read_stream<ZkbDecommitRec>(vm, ZKB_STREAM_DECOMMIT);
  s.frames = read_stream<ZkbFrameRec>(vm, ZKB_STREAM_FRAME);
  s.refunds = read_stream<ZkbRefundRec>(vm, ZKB_STREAM_REFUND);
  return replay_records(s, wt, initial);
}

inline VmLocalState GpuVmBatch::replay_encoded(const zkb_codec::EncodedView& blob, uint32_t vm, VmWitnessTracer& wt, const VmLocalState& initial) {
  VmStreams s;
  auto take = [&](uint32_t kind, auto& vec) {
    using Rec = typename std::remove_reference<decltype(vec)>::type::value_type;
    const uint64_t n = blob.decode(vm, kind, nullptr, 0);
    if (n == UINT64_MAX) throw std::runtime_error("replay_encoded: malformed blob");
    vec.resize(n / sizeof(Rec));
    if (n && blob.decode(vm, kind, vec.data(), n) != n) throw std::runtime_error("replay_encoded: malformed blob");
  };
  // cycle rows and memory queries are coded jointly (format v2): one walk decodes both
  s.rows.resize(blob.counts(vm)[ZKB_STREAM_ROWS]);
  s.mems.resize(blob.counts(vm)[ZKB_STREAM_MEM]);
  if (!blob.decode_rows_mem(vm, s.rows.data(), s.rows.size() * sizeof(ZkbCycleRow), s.mems.data(), s.mems.size() * sizeof(ZkbMemoryQueryRec)))
    throw std::runtime_error("replay_encoded: malformed blob");
  take(ZKB_STREAM_LOG, s.logs);
  take(ZKB_STREAM_DECOMMIT, s.decs);
  take(ZKB_STREAM_FRAME, s.frames);
  take(ZKB_STREAM_REFUND, s.refunds);
  return replay_records(s, wt, initial);
}

inline VmLocalState GpuVmBatch::replay_records(const VmStreams& streams, VmWitnessTracer& wt, const VmLocalState& initial) {
  const auto &rows = streams.rows;
  const auto &mems = streams.mems;
  const auto &logs = streams.logs;
  const auto &decs = streams.decs;
  const auto &frames = streams.frames;
  const auto &refunds = streams.refunds;
  size_t im = 0, il = 0, id = 0, ifr = 0, ir = 0;
  VmLocalState st = initial;
  // the bootloader push (helpers.rs:289-316) is the VM's first frame record and belongs to no cycle; `initial`
  // already reflects it
  size_t frames_in_rows = 0;
  for (const ZkbCycleRow& row : rows) frames_in_rows += (row.n_dfr >> 2) & 3u;
  if (frames.size() == frames_in_rows + 1 && frames[0].kind == ZKB_FRAMEKIND_START) ifr = 1;
  auto fail = [&](const char* what) { throw std::runtime_error(std::string("replay: inconsistent streams: ") + what); };

  for (const ZkbCycleRow& row : rows) {
    if (row.cycle != st.monotonic_cycle_counter) fail("cycle counter gap");
    wt.start_new_execution_cycle(st);
    const uint32_t cyc = row.cycle;
    const uint32_t n_mem = row.n_mem, n_log = row.n_log, n_dec = row.n_dfr & 3u, n_frame = (row.n_dfr >> 2) & 3u, n_refund = (row.n_dfr >> 4) & 3u;
    if (im + n_mem > mems.size() || il + n_log > logs.size() || id + n_dec > decs.size() || ifr + n_frame > frames.size() ||
        ir + n_refund > refunds.size())
      fail("per-cycle record counts exceed the streams");
    const uint32_t entry = ZK_OPCODE_TABLE[row.masked_variant];
    const uint32_t family = entry & 15u, sub = (entry >> ZK_E_SUB_SHIFT) & 15u;
    const bool masked = row.error_flags != 0 || !row.cond_resolved;
    const uint64_t ops = masked ? 0ull : row.raw_opcode;  // mask_into_panic / mask_into_nop zero the operands
    const uint32_t dst0_reg = (uint32_t)(ops >> 24) & 15u, dst1_reg = (uint32_t)(ops >> 28) & 15u;
    const uint32_t dst_mode = (entry >> ZK_E_DST_SHIFT) & 3u;

    // previous_code_word follows the instruction fetch (cycle.rs:59-100): a fetch happened iff the page or super-pc moved
    const bool was_pending = st.pending_exception;
    const uint16_t super_pc = (uint16_t)(row.pc_before >> 2);
    size_t m_end = im + n_mem;
    if (!was_pending && !(row.bits & ZKB_ROWBIT_SKIP) &&
        (st.current.code_page != st.previous_code_memory_page || st.previous_super_pc != super_pc)) {
      if (im >= m_end || mems[im].memory_type != (uint8_t)MemoryType::Code) fail("missing instruction-fetch query");
      st.previous_code_word = U256::from_limbs32(mems[im].value);
    }
    st.previous_code_memory_page = st.current.code_page;

    // 1. memory reads issued by the VM itself
    std::vector<MemoryQuery> pre_in, pre_out;
    size_t k = im;
    for (; k < m_end && mems[k].origin == ZKB_MEMORIGIN_VM && !mems[k].rw_flag; k++) wt.add_memory_query(cyc, decode(mems[k]));
    // 2. refund, then the log query (log.rs:99-102,218-219)
    LogQuery last_log{};
    for (uint32_t i = 0; i < std::max(n_refund, n_log); i++) {
      if (i < n_log) last_log = decode(logs[il + i]);
      if (i < n_refund) {
        // the refund callback receives the PARTIAL query of the SSTORE it prices: key / written value from the operands,
        // read_value = 0, is_service = false (log.rs:84-96, helpers.rs:119-136) -- also when the SSTORE then runs out of
        // ergs and its query is never executed (log.rs:136-144,196-199)
        LogQuery partial{};
        partial.timestamp = row.timestamp + 1;
        partial.tx_number_in_block = st.tx_number_in_block;
        partial.aux_byte = ZK_STORAGE_AUX_BYTE;
        partial.shard_id = st.current.this_shard_id;
        partial.address = st.current.this_address;
        partial.key = U256::from_limbs32(row.src0);
        partial.written_value = U256::from_limbs32(row.src1);
        partial.rw_flag = true;
        wt.record_refund_for_query(cyc, partial, RefundType{(RefundKind)refunds[ir + i].refund_type, refunds[ir + i].refund_value});
      }
      if (i < n_log) wt.add_log_query(cyc, last_log);
    }
    // 3. precompile memory witness (helpers.rs:208-222)
    for (; k < m_end && mems[k].origin != ZKB_MEMORIGIN_VM; k++)
      (mems[k].origin == ZKB_MEMORIGIN_PRECOMPILE_IN ? pre_in : pre_out).push_back(decode(mems[k]));
    if (family == ZK_OP_LOG && sub == ZK_LOG_PRECOMPILE && n_log)
      wt.add_precompile_call_result(cyc, last_log, pre_in, pre_out, precompile_rounds(last_log, pre_in, pre_out));
    // 4. decommitment (helpers.rs:164-194)
    for (uint32_t i = 0; i < n_dec; i++) {
      const ZkbDecommitRec& d = decs[id + i];
      // B = true decommitter (decommitter.rs:43-47,81-97): the code words on a fresh decommit, an empty vector on a repeat
      const U256 hash = U256::from_limbs32(d.hash);
      static const std::vector<U256> no_words;
      wt.add_decommittment(cyc, DecommittmentQuery{hash, d.timestamp, d.memory_page, d.decommitted_length, d.is_fresh != 0},
                           d.is_fresh ? code_words(hash) : no_words);
    }
    // 5. frames (helpers.rs:225-264)
    for (uint32_t i = 0; i < n_frame; i++) {
      const ZkbFrameRec& f = frames[ifr + i];
      if (f.kind == ZKB_FRAMEKIND_START) {
        CallStackEntry prev = st.current;
        prev.ergs_remaining = f.prev_ergs_remaining;
        prev.pc = f.prev_pc;
        prev.sp = f.prev_sp;
        if (f.prev_bound_kind == 1) prev.heap_bound = f.prev_bound_value;  // far_call grew the caller's bound (far_call.rs:330-385)
        if (f.prev_bound_kind == 2) prev.aux_heap_bound = f.prev_bound_value;
        CallStackEntry next = decode_new_frame(f);
        wt.start_new_execution_context(cyc, prev, next);
        st.inner.push_back(prev);
        st.current = next;
      } else {
        wt.finish_execution_context(cyc, f.panicked != 0);
        if (st.inner.empty()) fail("frame finish without a parent");
        st.current = st.inner.back();
        st.inner.pop_back();
      }
    }
    // 6. memory writes issued by the VM (UMA stores, dst0 on the stack)
    for (; k < m_end; k++) {
      if (mems[k].origin != ZKB_MEMORIGIN_VM) fail("precompile witness after VM writes");
      wt.add_memory_query(cyc, decode(mems[k]));
    }
    im = m_end;
    il += n_log;
    id += n_dec;
    ifr += n_frame;
    ir += n_refund;

    // ---- state after the cycle ----
    const bool far_call_done = family == ZK_OP_FAR_CALL && n_frame == 1;
    const bool far_ret_done = family == ZK_OP_RET && n_frame == 1 && (row.bits & ZKB_ROWBIT_DST0_VALID);
    if (far_call_done) {  // far_call.rs:573-610
      PrimitiveValue r1{U256::from_limbs32(row.dst0), true}, r2{U256::from_limbs32(row.dst1), false};
      const bool to_system = (r2.value.low_u32() & 2u) != 0;
      st.registers[0] = r1;
      st.registers[1] = r2;
      for (int i = 2; i < 12; i++) {
        if (to_system)
          st.registers[i].is_pointer = false;
        else
          st.registers[i] = PrimitiveValue{};
      }
      for (int i = 12; i < 15; i++) st.registers[i] = PrimitiveValue{};
    } else if (far_ret_done) {  // ret.rs:213-236
      st.registers[0] = PrimitiveValue{U256::from_limbs32(row.dst0), true};
      for (int i = 1; i < 15; i++) st.registers[i] = PrimitiveValue{};
    } else {
      if ((row.bits & ZKB_ROWBIT_DST0_VALID) && dst_mode == ZK_DST_REG && dst0_reg != 0)
        st.registers[dst0_reg - 1] = PrimitiveValue{U256::from_limbs32(row.dst0), (row.bits & ZKB_ROWBIT_DST0_PTR) != 0};
      if ((row.bits & ZKB_ROWBIT_DST1_VALID) && dst1_reg != 0)
        st.registers[dst1_reg - 1] = PrimitiveValue{U256::from_limbs32(row.dst1), (row.bits & ZKB_ROWBIT_DST1_PTR) != 0};
    }
    st.flags.overflow_or_less_than_flag = row.flags_after & 1u;
    st.flags.equality_flag = row.flags_after & 2u;
    st.flags.greater_than_flag = row.flags_after & 4u;
    st.pending_exception = (row.bits & ZKB_ROWBIT_PENDING) != 0;
    if (!(row.bits & ZKB_ROWBIT_SKIP)) st.timestamp = row.timestamp + ZK_TIME_DELTA_PER_CYCLE;
    st.monotonic_cycle_counter = row.cycle + 1;
    st.spent_pubdata_counter = row.spent_pubdata;
    st.memory_page_counter = row.memory_page_counter;
    st.current_ergs_per_pubdata_byte = row.ergs_per_pubdata;
    st.tx_number_in_block = row.tx_number;
    st.previous_super_pc = row.previous_super_pc;
    std::memcpy(st.context_u128_register.data(), row.context_u128, 16);
    st.current.pc = row.pc_after;
    st.current.sp = row.sp_after;
    st.current.ergs_remaining = row.ergs_after;
    st.current.heap_bound = row.heap_bound;
    st.current.aux_heap_bound = row.aux_heap_bound;
    if (st.inner.size() != row.callstack_depth) fail("callstack depth mismatch");
    if (st.current.code_page != row.code_page || st.current.base_memory_page != row.base_page ||
        st.current.exception_handler_location != row.exception_handler)
      fail("current frame mismatch");
    wt.end_execution_cycle(st);
  }
  if (im != mems.size() || il != logs.size() || id != decs.size() || ifr != frames.size() || ir != refunds.size())
    fail("records left over after the last cycle");
  return st;
}

}  // namespace zkb_host
