// Test shim for include/zkb_host.hpp: replays one VM of an existing batch through a recording VmWitnessTracer and
// compares the VmLocalState tracked by the replay with the batch's own final local state.  Built by
// tests/test_host_replay.py against the CPU oracle (-DZKB_HOST_PREFIX=orc_) and, on a GPU box, against libzkb.so.
#include <cstdio>
#include <cstring>

#include "../include/zkb_host.hpp"

using namespace zkb_host;

struct Recorder : VmWitnessTracer {
  uint64_t n[10] = {0};
  uint32_t cur_ts = 0, cur_cycle = 0;
  std::string problem;
  const VmLocalState* last_end = nullptr;
  VmLocalState end_copy;
  bool have_end = false;
  void note(const std::string& p) {
    if (problem.empty()) problem = p;
  }
  void start_new_execution_cycle(const VmLocalState& s) override {
    n[0]++;
    cur_ts = s.timestamp;
    cur_cycle = s.monotonic_cycle_counter;
    if (have_end && (s.monotonic_cycle_counter != end_copy.monotonic_cycle_counter || s.timestamp != end_copy.timestamp))
      note("cycle start state differs from the previous cycle's end state");
  }
  void end_execution_cycle(const VmLocalState& s) override {
    n[1]++;
    end_copy = s;
    have_end = true;
    if (s.monotonic_cycle_counter != cur_cycle + 1) note("monotonic_cycle_counter did not advance by one");
  }
  void check_cycle(uint32_t c) {
    if (c != cur_cycle) note("callback tagged with the wrong cycle");
  }
  void add_memory_query(uint32_t c, const MemoryQuery& q) override {
    n[2]++;
    check_cycle(c);
    if (q.timestamp < cur_ts || q.timestamp > cur_ts + 3) note("memory query timestamp outside [t, t+3]");
    if (q.rw_flag && q.timestamp != cur_ts + 3) note("VM memory write not at t+3 (mod.rs:228-231)");
    if (!q.rw_flag && q.timestamp != cur_ts) note("VM memory read not at t+0 (mod.rs:220-223)");
  }
  void record_refund_for_query(uint32_t c, const LogQuery& q, RefundType) override {
    n[3]++;
    check_cycle(c);
    if (!q.rw_flag) note("refund for a read");
  }
  void add_log_query(uint32_t c, const LogQuery& q) override {
    n[4]++;
    check_cycle(c);
    if (q.timestamp != cur_ts + 1) note("log query not at t+1 (mod.rs:224-227)");
  }
  uint64_t fresh_code_words = 0, precompile_rounds_seen = 0;
  void add_decommittment(uint32_t c, const DecommittmentQuery& q, const std::vector<U256>& words) override {
    n[5]++;
    check_cycle(c);
    if (q.timestamp != cur_ts + 1) note("decommit not at t+1");
    // decommitter.rs:43-47,81-97: the code words on a fresh decommit, nothing on a repeat
    if (q.is_fresh && words.size() != q.decommitted_length) note("fresh decommit without its code words");
    if (!q.is_fresh && !words.empty()) note("repeated decommit carries code words");
    if (q.is_fresh) {
      fresh_code_words += words.size();
      // versioned hash: bytes 2..3 = length in words (far_call.rs:169-252)
      if (((q.hash.limbs[3] >> 32) & 0xFFFFu) != words.size()) note("code words do not match the length in the versioned hash");
    }
  }
  void add_precompile_call_result(uint32_t c, const LogQuery& req, const std::vector<MemoryQuery>& in, const std::vector<MemoryQuery>& out,
                                  const PrecompileCyclesWitness& w) override {
    n[6]++;
    check_cycle(c);
    size_t reads = 0, writes = 0;
    for (auto& r : w.rounds) {
      reads += r.reads.size();
      writes += r.writes.size();
      if (r.reads.size() > 6) note("precompile round with more than 6 reads");
    }
    if (reads != in.size() || writes != out.size()) note("precompile rounds do not partition the memory witness");
    if (w.rounds.empty() || !w.rounds.front().has_new_request || w.rounds.front().new_request.timestamp != req.timestamp)
      note("first precompile round without its request");
    if (w.kind == PrecompileKind::Keccak256 && w.rounds.size() != ((uint32_t)(req.key.limbs[0] >> 32)) / 136 + 1) note("keccak round count");
    precompile_rounds_seen += w.rounds.size();
    for (auto& q : in)
      if (q.rw_flag || q.timestamp != cur_ts + 1) note("precompile read witness malformed");
    for (auto& q : out)
      if (!q.rw_flag || q.timestamp != cur_ts + 2) note("precompile write witness malformed");
    n[8] += in.size() + out.size();
  }
  void start_new_execution_context(uint32_t c, const CallStackEntry&, const CallStackEntry&) override {
    n[7]++;
    check_cycle(c);
  }
  void finish_execution_context(uint32_t c, bool) override {
    n[9]++;
    check_cycle(c);
  }
};

extern "C" int host_replay_check(void* handle, uint32_t n_vms, uint32_t vm, const ZkbFrame* boot, const uint8_t* init_regs_be,
                                 uint32_t init_ptr_mask, uint32_t init_page_counter, uint32_t init_epp, uint64_t counts[10], char* err,
                                 int errlen, const void* encoded_blob, uint64_t encoded_bytes) {
  try {
    GpuVmBatch b = GpuVmBatch::wrap((ZkbBatch*)handle, n_vms);
    VmLocalState init;
    init.timestamp = ZK_STARTING_TIMESTAMP;
    init.memory_page_counter = init_page_counter;
    init.current_ergs_per_pubdata_byte = init_epp;
    for (int r = 0; r < 15; r++) {
      uint32_t limbs[8];
      for (int l = 0; l < 8; l++) {
        const uint8_t* p = init_regs_be + 32 * r + 28 - 4 * l;
        limbs[l] = (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3];
      }
      init.registers[r] = PrimitiveValue{U256::from_limbs32(limbs), ((init_ptr_mask >> r) & 1u) != 0};
    }
    CallStackEntry root;  // CallStackEntry::empty_context (execution_stack.rs:35-55) minus the bootloader's ergs
    root.sp = ZK_INITIAL_SP_ON_FAR_CALL;
    root.ergs_remaining = ZK_VM_INITIAL_FRAME_ERGS - boot->ergs_remaining;
    init.inner.push_back(root);
    init.current = from_frame(*boot);
    Recorder rec;
    VmLocalState fin = b.replay(vm, rec, init);
    if (encoded_blob) {   // the same replay straight from the transport blob must make the same calls and end in the same state
      zkb_codec::EncodedView view;
      if (!view.open(encoded_blob, encoded_bytes)) throw std::runtime_error("encoded blob rejected");
      Recorder rec2;
      VmLocalState fin2 = b.replay_encoded(view, vm, rec2, init);
      if (std::memcmp(rec.n, rec2.n, sizeof(rec.n)) || !rec2.problem.empty()) throw std::runtime_error("replay_encoded: callback counts differ from replay");
      if (fin2.monotonic_cycle_counter != fin.monotonic_cycle_counter || fin2.timestamp != fin.timestamp || fin2.current.pc != fin.current.pc ||
          fin2.current.ergs_remaining != fin.current.ergs_remaining)
        throw std::runtime_error("replay_encoded: final state differs from replay");
      for (int r = 0; r < 15; r++)
        if (fin2.registers[r].value != fin.registers[r].value) throw std::runtime_error("replay_encoded: registers differ from replay");
    }
    std::memcpy(counts, rec.n, sizeof(rec.n));
    if (!rec.problem.empty()) throw std::runtime_error(rec.problem);
    // the tracked state must equal the batch's own final VmLocalState
    ZkbLocalState ls;
    if (ZKB_FN(read_local_state)((ZkbBatch*)handle, vm, &ls) != ZKB_OK) throw std::runtime_error("read_local_state failed");
    auto mismatch = [&](const char* what) { throw std::runtime_error(std::string("tracked state differs from the batch: ") + what); };
    for (int r = 0; r < 15; r++) {
      if (fin.registers[r].value != U256::from_limbs32(ls.registers[r])) mismatch(("register r" + std::to_string(r + 1)).c_str());
      if (fin.registers[r].is_pointer != (((ls.register_is_pointer >> r) & 1u) != 0)) mismatch(("pointer bit of r" + std::to_string(r + 1)).c_str());
    }
    uint8_t fl = (fin.flags.overflow_or_less_than_flag ? 1 : 0) | (fin.flags.equality_flag ? 2 : 0) | (fin.flags.greater_than_flag ? 4 : 0);
    if (fl != ls.flags) mismatch("flags");
    if (fin.pending_exception != (ls.pending_exception != 0)) mismatch("pending_exception");
    if (fin.timestamp != ls.timestamp) mismatch("timestamp");
    if (fin.monotonic_cycle_counter != ls.monotonic_cycle_counter) mismatch("monotonic_cycle_counter");
    if (fin.spent_pubdata_counter != ls.spent_pubdata_counter) mismatch("spent_pubdata_counter");
    if (fin.memory_page_counter != ls.memory_page_counter) mismatch("memory_page_counter");
    if (fin.current_ergs_per_pubdata_byte != ls.current_ergs_per_pubdata_byte) mismatch("ergs_per_pubdata");
    if (fin.tx_number_in_block != ls.tx_number_in_block) mismatch("tx_number_in_block");
    if (fin.previous_super_pc != ls.previous_super_pc) mismatch("previous_super_pc");
    if (fin.previous_code_memory_page != ls.previous_code_memory_page) mismatch("previous_code_memory_page");
    if (fin.previous_code_word != U256::from_limbs32(ls.previous_code_word)) mismatch("previous_code_word");
    if (std::memcmp(fin.context_u128_register.data(), ls.context_u128_register, 16)) mismatch("context_u128_register");
    if (fin.inner.size() != ls.callstack_depth) mismatch("callstack depth");
    const ZkbFrame& f = ls.current_frame;
    const CallStackEntry& c = fin.current;
    if (std::memcmp(c.this_address.data(), f.this_address, 20) || std::memcmp(c.msg_sender.data(), f.msg_sender, 20) ||
        std::memcmp(c.code_address.data(), f.code_address, 20))
      mismatch("current frame addresses");
    if (c.pc != f.pc || c.sp != f.sp || c.ergs_remaining != f.ergs_remaining || c.code_page != f.code_page ||
        c.base_memory_page != f.base_memory_page || c.heap_bound != f.heap_bound || c.aux_heap_bound != f.aux_heap_bound ||
        c.exception_handler_location != f.exception_handler_location || c.is_static != (f.is_static != 0) ||
        c.is_local_frame != (f.is_local_frame != 0) || c.this_shard_id != f.this_shard_id || c.caller_shard_id != f.caller_shard_id ||
        c.code_shard_id != f.code_shard_id || std::memcmp(c.context_u128_value.data(), f.context_u128_value, 16))
      mismatch("current frame scalars");
    return 0;
  } catch (const std::exception& e) {
    std::snprintf(err, errlen, "%s", e.what());
    return 1;
  }
}


// ---- expected output of the device-side consumer (zkb_consume): the VmLocalState the reference hands to
// start_new_execution_cycle every `period` cycles and after the last cycle, serialised as ZkbLocalState ----------------
static void serialise(const VmLocalState& s, ZkbLocalState* o) {
  std::memset(o, 0, sizeof(*o));
  s.previous_code_word.to_limbs32(o->previous_code_word);
  o->previous_code_memory_page = s.previous_code_memory_page;
  for (int r = 0; r < 15; r++) {
    s.registers[r].value.to_limbs32(o->registers[r]);
    if (s.registers[r].is_pointer) o->register_is_pointer |= (uint16_t)(1u << r);
  }
  o->flags = (s.flags.overflow_or_less_than_flag ? 1 : 0) | (s.flags.equality_flag ? 2 : 0) | (s.flags.greater_than_flag ? 4 : 0);
  o->pending_exception = s.pending_exception ? 1 : 0;
  o->timestamp = s.timestamp;
  o->monotonic_cycle_counter = s.monotonic_cycle_counter;
  o->spent_pubdata_counter = s.spent_pubdata_counter;
  o->memory_page_counter = s.memory_page_counter;
  o->absolute_execution_step = 0;
  o->current_ergs_per_pubdata_byte = s.current_ergs_per_pubdata_byte;
  o->tx_number_in_block = s.tx_number_in_block;
  o->previous_super_pc = s.previous_super_pc;
  std::memcpy(o->context_u128_register, s.context_u128_register.data(), 16);
  o->callstack_depth = (uint32_t)s.inner.size();
  const CallStackEntry& c = s.current;
  ZkbFrame& f = o->current_frame;
  std::memcpy(f.this_address, c.this_address.data(), 20);
  std::memcpy(f.msg_sender, c.msg_sender.data(), 20);
  std::memcpy(f.code_address, c.code_address.data(), 20);
  f.base_memory_page = c.base_memory_page;
  f.code_page = c.code_page;
  f.sp = c.sp;
  f.pc = c.pc;
  f.exception_handler_location = c.exception_handler_location;
  f.ergs_remaining = c.ergs_remaining;
  f.this_shard_id = c.this_shard_id;
  f.caller_shard_id = c.caller_shard_id;
  f.code_shard_id = c.code_shard_id;
  f.is_static = c.is_static;
  f.is_local_frame = c.is_local_frame;
  std::memcpy(f.context_u128_value, c.context_u128_value.data(), 16);
  f.heap_bound = c.heap_bound;
  f.aux_heap_bound = c.aux_heap_bound;
}

struct Snapshotter : VmWitnessTracer {
  uint32_t period;
  std::vector<ZkbLocalState> out;
  VmLocalState last;
  bool any = false;
  void start_new_execution_cycle(const VmLocalState& s) override {
    if (s.monotonic_cycle_counter != 0 && s.monotonic_cycle_counter % period == 0) {
      out.emplace_back();
      serialise(s, &out.back());
    }
  }
  void end_execution_cycle(const VmLocalState& s) override {
    last = s;
    any = true;
  }
};

extern "C" int host_expected_snapshots(void* handle, uint32_t n_vms, uint32_t vm, const ZkbFrame* boot, const uint8_t* init_regs_be, uint32_t init_ptr_mask,
                                       uint32_t init_page_counter, uint32_t init_epp, uint32_t period, ZkbLocalState* out, uint32_t cap,
                                       uint32_t* n_out, char* err, int errlen) {
  try {
    GpuVmBatch b = GpuVmBatch::wrap((ZkbBatch*)handle, n_vms);
    VmLocalState init;
    init.timestamp = ZK_STARTING_TIMESTAMP;
    init.memory_page_counter = init_page_counter;
    init.current_ergs_per_pubdata_byte = init_epp;
    for (int r = 0; r < 15; r++) {
      uint32_t limbs[8];
      for (int l = 0; l < 8; l++) {
        const uint8_t* p = init_regs_be + 32 * r + 28 - 4 * l;
        limbs[l] = (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3];
      }
      init.registers[r] = PrimitiveValue{U256::from_limbs32(limbs), ((init_ptr_mask >> r) & 1u) != 0};
    }
    CallStackEntry root;
    root.sp = ZK_INITIAL_SP_ON_FAR_CALL;
    root.ergs_remaining = ZK_VM_INITIAL_FRAME_ERGS - boot->ergs_remaining;
    init.inner.push_back(root);
    init.current = from_frame(*boot);
    Snapshotter sn;
    sn.period = period;
    b.replay(vm, sn, init);
    if (sn.any) {   // the state after the last cycle (a boundary that coincides with it is reported once, here)
      sn.out.emplace_back();
      serialise(sn.last, &sn.out.back());
    }
    *n_out = (uint32_t)sn.out.size();
    if (sn.out.size() > cap) throw std::runtime_error("snapshot buffer too small");
    std::memcpy(out, sn.out.data(), sn.out.size() * sizeof(ZkbLocalState));
    return 0;
  } catch (const std::exception& e) {
    std::snprintf(err, errlen, "%s", e.what());
    return 1;
  }
}
