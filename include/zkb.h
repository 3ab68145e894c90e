/* zkb — C ABI of the B200 batched out-of-circuit EraVM witness generator.
 *
 * The reference (matter-labs/era-zk_evm, crate zk_evm 1.4.1) has no FFI boundary: its seam is the generic
 * `VmState<S, M, EV, PP, DP, WT>` (/root/reference/src/vm_state/mod.rs:157-207) driven by a caller loop
 * `while !vm.execution_has_ended() { vm.cycle(&mut tracer)? }` (cycle.rs:257, mod.rs:214).  This header is what a
 * Rust `-sys` crate would bind so that a `Vec<VmState<InMemoryStorage, SimpleMemory, InMemoryEventSink,
 * DefaultPrecompilesProcessor, SimpleDecommitter, WT>>` plus that loop can be replaced by ONE batch object that
 * lives on one B200 (see INTEGRATION.md for the Rust shim that replays the recorded streams into a
 * `VmWitnessTracer` in the reference's callback order).
 *
 * Conventions: every function returns an int32 status (ZKB_OK == 0); no exceptions or callbacks cross the
 * boundary; pointers are plain host pointers unless named `dptr`; U256 values cross as 32 big-endian bytes
 * (the reference's own byte convention, src/utils.rs:12-15,36-48); a ZkbBatch is single-owner and
 * non-reentrant, distinct batches may be driven from distinct threads.
 */
#ifndef ZKB_H
#define ZKB_H
#include <stdint.h>
#include "zkb_records.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes -------------------------------------------------------------------------------- */
enum ZkbStatus {
  ZKB_OK = 0,
  ZKB_ERR_INVALID_ARGUMENT = 1,
  ZKB_ERR_CUDA = 2,
  ZKB_ERR_OUT_OF_MEMORY = 3,
  ZKB_ERR_UNKNOWN_BYTECODE = 4,
  ZKB_ERR_NO_DEVICE = 5
};

/* per-VM status (ZkbVmStatus.code) */
enum ZkbVmCode {
  ZKB_VM_RUNNING = 0,               /* not finished: call zkb_run again to continue                               */
  ZKB_VM_ENDED = 1,                 /* execution_has_ended() (mod.rs:214)                                         */
  ZKB_VM_UNKNOWN_CODE_HASH = 2,     /* the crate's only Err: decommitter.rs:50-56 via far_call.rs:448             */
  ZKB_VM_REFERENCE_PANIC = 3,       /* a reference assert!/expect would have fired (e.g. memory.rs:415-437,483)   */
  ZKB_VM_CAP_STREAM = 16,           /* a per-VM stream slab is full                                               */
  ZKB_VM_CAP_STACK = 17,            /* stack index beyond ZkbConfig.stack_words                                   */
  ZKB_VM_CAP_HEAP = 18,             /* heap access beyond ZkbConfig.heap_bytes, or no free heap slab              */
  ZKB_VM_CAP_DEPTH = 19,            /* callstack deeper than max_depth / max_far_depth                            */
  ZKB_VM_CAP_STORAGE = 20,          /* per-VM storage table or rollback journal full                              */
  ZKB_VM_CAP_PAGES = 21,            /* page-indirection / decommit-history table full                             */
  ZKB_VM_UNSUPPORTED = 22           /* reached a precompile that is not implemented in this build                 */
};

/* ---- configuration ------------------------------------------------------------------------------- */
enum ZkbSchedule {
  ZKB_SCHED_AUTO = 0,      /* library default (lockstep) */
  ZKB_SCHED_FREE = 1,      /* every warp pulls the next VM from a queue and runs it to the end on its own   */
  ZKB_SCHED_LOCKSTEP = 2   /* the warps of a CTA step consecutive VMs cycle by cycle (shared I-cache lines): */
                           /* best when neighbouring VMs execute the same contracts                          */
};

typedef struct ZkbConfig {
  uint32_t n_vms;
  int32_t device;            /* CUDA device ordinal */
  uint32_t witness_mode;     /* 1 = record all streams (recording tracer); 0 = DummyTracer (witness_trace/mod.rs:75) */
  uint32_t cap_records[ZKB_N_STREAMS]; /* per-VM capacity of each stream, in records */
  uint32_t stack_words;      /* per far frame; bounded analogue of MAX_STACK_PAGE_SIZE_IN_WORDS (memory.rs:176) */
  uint32_t heap_bytes;       /* per heap page slab, multiple of 32 (reference heaps grow on demand, memory.rs:194) */
  uint32_t n_heap_slabs;     /* heap + aux heap of live far frames + extended-lifetime returndata pages, <= 32 */
  uint32_t max_far_depth;    /* far-call frames incl. bootloader */
  uint32_t max_depth;        /* all frames (near + far) */
  uint32_t storage_slots;    /* per-VM open-addressed storage table, power of two */
  uint32_t journal_entries;  /* per-VM storage rollback journal (storage.rs:98-120) */
  uint32_t host_mirror;      /* 1 = allocate pinned host mirrors for zkb_fetch_streams */
  uint32_t schedule;         /* ZkbSchedule: how warps are assigned to VMs (results are identical either way) */
  uint32_t reserved[3];      /* reserved[0] = SMs left free by the persistent interpreter grid (0 = none): room for the
                                NCCL send/recv kernels of the multi-GPU stream concat to run underneath the next launch.
                                reserved[1] = warm_write_refund_bytes (the refund-aware storage oracle):
                                0 = the reference's InMemoryStorage, estimate_refunds_for_write() == RefundType::None
                                (storage.rs:80-86); 1..64 = a storage write to a slot whose cold/warm marker
                                (storage.rs:10,105-110,126-131) is already set is answered with
                                RefundType::RepeatedWrite{pubdata_bytes = this value} on the rollup shard, which
                                log.rs:99-119 subtracts from INITIAL_STORAGE_WRITE_PUBDATA_BYTES.
                                reserved[2] = unused. */
} ZkbConfig;

/* mirror of CallStackEntry (execution_stack.rs:6-24) */
typedef struct ZkbFrame {
  uint8_t this_address[20];
  uint8_t msg_sender[20];
  uint8_t code_address[20];
  uint32_t base_memory_page;
  uint32_t code_page;
  uint16_t sp;
  uint16_t pc;
  uint16_t exception_handler_location;
  uint16_t reserved0;
  uint32_t ergs_remaining;
  uint8_t this_shard_id;
  uint8_t caller_shard_id;
  uint8_t code_shard_id;
  uint8_t is_static;
  uint8_t is_local_frame;
  uint8_t reserved1[3];
  uint32_t context_u128_value[4];
  uint32_t heap_bound;
  uint32_t aux_heap_bound;
} ZkbFrame;

/* mirror of VmLocalState (vm_state/mod.rs:54-73) with the callstack flattened to its current entry + depth */
typedef struct ZkbLocalState {
  uint32_t previous_code_word[8];
  uint32_t previous_code_memory_page;
  uint32_t registers[15][8];
  uint16_t register_is_pointer;   /* bit i = registers[i].is_pointer */
  uint8_t flags;                  /* bit0 LT/OF, bit1 EQ, bit2 GT */
  uint8_t pending_exception;
  uint32_t timestamp;
  uint32_t monotonic_cycle_counter;
  uint32_t spent_pubdata_counter;
  uint32_t memory_page_counter;
  uint32_t absolute_execution_step;
  uint32_t current_ergs_per_pubdata_byte;
  uint16_t tx_number_in_block;
  uint16_t previous_super_pc;
  uint32_t context_u128_register[4];
  uint32_t callstack_depth;
  ZkbFrame current_frame;
} ZkbLocalState;

/* One per-circuit-batch snapshot of the device-side consumer (zkb_consume): the VmLocalState handed to
 * start_new_execution_cycle of cycle `cycle` (src/witness_trace/mod.rs:11-72), how many records of the memory / log /
 * decommitment queues precede that cycle, and the queues' running commitments at that point (chaining values of the sha256
 * chain over the records, each zero-padded to 64-byte blocks; csrc/consume.cuh says why sha256). */
typedef struct ZkbSnapshot {
  ZkbLocalState state;
  uint32_t cycle;
  uint32_t n_mem, n_log, n_decommit;
  uint32_t queue_state[3][8];     /* memory, log, decommitment */
} ZkbSnapshot;

typedef struct ZkbVmStatus {
  uint32_t code;     /* ZkbVmCode */
  uint32_t cycles;   /* cycles executed so far (== monotonic_cycle_counter) */
} ZkbVmStatus;

typedef struct ZkbStorageInit {
  uint8_t shard_id;
  uint8_t reserved[3];
  uint8_t address[20];
  uint8_t key_be[32];
  uint8_t value_be[32];
} ZkbStorageInit;

enum ZkbLocalField {
  ZKB_FIELD_MEMORY_PAGE_COUNTER = 0,
  ZKB_FIELD_ERGS_PER_PUBDATA = 1,
  ZKB_FIELD_TX_NUMBER = 2,
  ZKB_FIELD_TIMESTAMP = 3
};

typedef struct ZkbBatch ZkbBatch;

/* ---- lifecycle: replaces VmState::empty_state for n_vms instances (mod.rs:188-207) --------------- */
int32_t zkb_create(const ZkbConfig* cfg, ZkbBatch** out);
int32_t zkb_destroy(ZkbBatch* b);
const char* zkb_last_error(void);
/* re-initialise every VM to empty_state, keeping bytecodes / block properties (new batch, same allocation) */
int32_t zkb_reset(ZkbBatch* b);

/* ---- population ---------------------------------------------------------------------------------- */
/* = SimpleDecommitter::populate (decommitter.rs:23-28); words are 32-byte big-endian code words */
int32_t zkb_load_bytecode(ZkbBatch* b, const uint8_t hash_be[32], const uint8_t* words_be, uint32_t n_words);
/* the code words filed under `hash_be` by zkb_load_bytecode (SimpleDecommitter.known_hashes, decommitter.rs:10-13): what
 * decommit_into_memory hands to the tracer on a FRESH decommit (decommitter.rs:81-97 -> helpers.rs:185-191).  Copies at
 * most max_words 32-byte big-endian words; *n_words_out = the code's length.  Unknown hash: ZKB_ERR_UNKNOWN_BYTECODE. */
int32_t zkb_read_bytecode(ZkbBatch* b, const uint8_t hash_be[32], uint8_t* words_be_out, uint32_t max_words, uint32_t* n_words_out);
/* = SimpleMemory::polulate_bootloaders_calldata (memory.rs:293-298): replaces the content of the always-present
 * extended-lifetime page BOOTLOADER_CALLDATA_PAGE (memory.rs:230-231) of VMs [vm_lo, vm_hi).  per_vm as in
 * zkb_populate_storage.  As in the reference the page has NO entry in page_numbers_indirections (memory.rs:232 registers
 * only page 0), so a fat-pointer read of it hits `expect("fat pointer only points to reachable memory")`
 * (memory.rs:478-481) = ZKB_VM_REFERENCE_PANIC; its content is visible through zkb_read_calldata only. */
int32_t zkb_set_calldata(ZkbBatch* b, uint32_t vm_lo, uint32_t vm_hi, const uint8_t* words_be, uint32_t n_words, uint32_t per_vm);
/* = dump_page_content(BOOTLOADER_CALLDATA_PAGE, word_lo .. word_lo + n_words) (memory.rs:300-344): words beyond the
 * populated length read as zero */
int32_t zkb_read_calldata(ZkbBatch* b, uint32_t vm, uint32_t word_lo, uint32_t n_words, uint8_t* words_be_out);
/* = BlockProperties (block_properties/mod.rs:4-7) */
int32_t zkb_set_block_properties(ZkbBatch* b, const uint8_t default_aa_code_hash_be[32], uint8_t zkporter_is_available);
/* = InMemoryStorage::populate (storage.rs:26-32).  per_vm == 0: the n entries are applied to every VM in
 * [vm_lo, vm_hi); per_vm != 0: `entries` holds (vm_hi - vm_lo) * n records, n per VM. */
int32_t zkb_populate_storage(ZkbBatch* b, uint32_t vm_lo, uint32_t vm_hi, const ZkbStorageInit* entries, uint32_t n,
                             uint32_t per_vm);
/* = SimpleMemory::populate_code (memory.rs:271-284): bind a code page number to a loaded bytecode */
int32_t zkb_populate_code(ZkbBatch* b, uint32_t vm_lo, uint32_t vm_hi, uint32_t page, const uint8_t hash_be[32]);
/* = VmState::push_bootloader_context (helpers.rs:289-316) */
int32_t zkb_push_bootloader_context(ZkbBatch* b, uint32_t vm_lo, uint32_t vm_hi, const ZkbFrame* frame);
/* = SimpleMemory::populate_heap (memory.rs:287-291) on the current (bootloader) frame's heap; bytes are the
 * big-endian memory image.  per_vm as in zkb_populate_storage (n_bytes per VM). */
int32_t zkb_populate_heap(ZkbBatch* b, uint32_t vm_lo, uint32_t vm_hi, const uint8_t* bytes, uint32_t n_bytes,
                          uint32_t per_vm);
/* local_state.registers[reg] = value (reg 0-based, 0..14) */
int32_t zkb_set_register(ZkbBatch* b, uint32_t vm_lo, uint32_t vm_hi, uint32_t reg, const uint8_t value_be[32],
                         uint8_t is_pointer, uint32_t per_vm);
int32_t zkb_set_local_field(ZkbBatch* b, uint32_t vm_lo, uint32_t vm_hi, uint32_t field, uint32_t value);

/* ---- execution: replaces the caller's `while !ended { cycle() }` loop ----------------------------- */
/* run every VM for at most max_cycles_per_vm further cycles (0 = until ended). Asynchronous on `cuda_stream`
 * (a cudaStream_t, may be NULL); resumable. */
int32_t zkb_run(ZkbBatch* b, uint32_t max_cycles_per_vm, void* cuda_stream);
int32_t zkb_sync(ZkbBatch* b);
/* elapsed device time (ms, CUDA events on the launch stream) of the last zkb_run, valid after zkb_sync */
int32_t zkb_last_run_ms(ZkbBatch* b, float* ms, uint32_t* n_kernel_launches);

/* ---- results ------------------------------------------------------------------------------------- */
int32_t zkb_vm_status(ZkbBatch* b, uint32_t vm_lo, uint32_t vm_hi, ZkbVmStatus* out);
int32_t zkb_read_local_state(ZkbBatch* b, uint32_t vm, ZkbLocalState* out);
/* per-VM record counts of one stream */
int32_t zkb_stream_counts(ZkbBatch* b, uint32_t kind, uint32_t vm_lo, uint32_t vm_hi, uint32_t* counts_out);
/* total bytes over all streams and VMs (the "algorithmic bytes" of SURVEY.md §8d) and total cycles */
int32_t zkb_totals(ZkbBatch* b, uint64_t* total_cycles, uint64_t stream_bytes[ZKB_N_STREAMS]);
/* device view: VM v's records of `kind` start at dptr + v * stride_bytes */
int32_t zkb_stream_device_view(ZkbBatch* b, uint32_t kind, void** dptr, uint64_t* stride_bytes);
/* copy one VM's stream to host memory (at most max_bytes); returns the byte length in *n_bytes */
int32_t zkb_read_stream(ZkbBatch* b, uint32_t vm, uint32_t kind, void* dst, uint64_t max_bytes, uint64_t* n_bytes);
/* pack every VM's stream `kind` contiguously (VM order) on the device and copy it to `host_dst`
 * (must hold the stream's total bytes); offsets_out[n_vms + 1] receives the byte offset of each VM. */
int32_t zkb_fetch_stream_packed(ZkbBatch* b, uint32_t kind, void* host_dst, uint64_t host_capacity,
                                uint64_t* offsets_out);
/* asynchronous variant: waits (on the host) only for THIS batch's last run, then enqueues pack + D2H on `cuda_stream`
 * and returns; host_dst should be pinned. The copy engine then overlaps with other batches' kernels: this is how a
 * host loop pipelines sub-batches (compute of k+1 under the witness download of k). Completion = stream sync. */
int32_t zkb_fetch_stream_packed_async(ZkbBatch* b, uint32_t kind, void* host_dst, uint64_t host_capacity,
                                      uint64_t* offsets_out, void* cuda_stream);
/* device-only variant: returns the packed device buffer (valid until the next run/fetch/destroy) */
int32_t zkb_pack_stream_device(ZkbBatch* b, uint32_t kind, void** dptr, uint64_t* n_bytes, void* cuda_stream);
/* as above without the final stream synchronisation: the pack kernel is only enqueued on `cuda_stream` (work queued
 * later on that stream, or a collective that waits on it, sees the packed buffer; used to overlap the multi-GPU
 * concatenation of step k with the interpreter launch of step k+1) */
int32_t zkb_pack_stream_device_async(ZkbBatch* b, uint32_t kind, void** dptr, uint64_t* n_bytes, void* cuda_stream);
/* ---- encoded transport (include/zkb_codec.h): the six streams of every VM as ONE lossless, self-describing blob ------
 * The encoder runs on the device (format v2: XOR with a per-word prediction + presence bitmaps, cycle rows and memory
 * queries coded jointly; 16.5 % of the canonical bytes on the ERC-20 workload), so a PCIe-bound host loop moves 6x fewer
 * bytes; zkb_decode_stream (host, no GPU needed) and zkb_codec::EncodedView give the canonical records back byte for byte.
 * _async: waits (on the host) for THIS batch's last run and for the encoding pass (which also yields the blob's size), then
 * enqueues the compaction + ONE D2H copy on `cuda_stream`; host_dst should be pinned; completion = stream sync.
 * host_dst == NULL: only *n_bytes (the blob size for the current streams) is returned.  A subset blob carries ROWS and MEM
 * together or not at all. */
int32_t zkb_fetch_encoded_async(ZkbBatch* b, void* host_dst, uint64_t host_capacity, uint64_t* n_bytes, void* cuda_stream);
int32_t zkb_fetch_encoded(ZkbBatch* b, void* host_dst, uint64_t host_capacity, uint64_t* n_bytes);
/* the same for a subset of the streams (bit k of kinds_mask = ZkbStreamKind k): what a host that runs zkb_consume on the
 * device still needs -- the query logs, not the rows.  Streams outside the mask decode as empty; counts are still reported. */
int32_t zkb_fetch_encoded_kinds_async(ZkbBatch* b, uint32_t kinds_mask, void* host_dst, uint64_t host_capacity, uint64_t* n_bytes, void* cuda_stream);
/* device-only variant: the blob in device memory (valid until the next encode / destroy), enqueued on cuda_stream */
int32_t zkb_encode_streams_device(ZkbBatch* b, void** dptr, uint64_t* n_bytes, void* cuda_stream);
/* host-side decoder: canonical records of VM `vm`'s stream `kind` out of a blob; dst == NULL returns the length only */
int32_t zkb_decode_stream(const void* blob, uint64_t blob_bytes, uint32_t vm, uint32_t kind, void* dst, uint64_t max_bytes, uint64_t* n_bytes);
/* bulk host-side decode: stream `kind` of EVERY VM as canonical records, VM-major, back to back (the layout of
 * zkb_fetch_stream_packed) on n_threads host threads (0 = all); offsets_out[n_vms + 1] = each VM's byte offset in dst.
 * dst == NULL: offsets only (offsets_out[n_vms] = bytes needed). */
int32_t zkb_decode_all(const void* blob, uint64_t blob_bytes, uint32_t kind, void* dst, uint64_t capacity, uint64_t* offsets_out,
                       uint32_t n_threads);
/* per-VM summary carried by a blob: 6 record counts, ZkbVmCode, cycles */
int32_t zkb_decode_counts(const void* blob, uint64_t blob_bytes, uint32_t vm, uint32_t counts_out[8]);
/* final storage value of one slot (== InMemoryStorage.inner lookup, storage.rs:9) */
int32_t zkb_read_storage(ZkbBatch* b, uint32_t vm, uint8_t shard_id, const uint8_t address[20],
                         const uint8_t key_be[32], uint8_t value_be_out[32]);
/* read back memory of the current frame's heap (debug / tests; = dump_page_content, memory.rs:300-313) */
int32_t zkb_read_heap(ZkbBatch* b, uint32_t vm, uint32_t byte_offset, uint32_t n_bytes, uint8_t* out);

/* ---- the device-side consumer: what a downstream-shaped VmWitnessTracer adapter builds, built on the GPU
 * zkb_consume walks every VM's streams once: a ZkbSnapshot every cycles_per_snapshot cycles (cycle K, 2K, ...; cycle 0 is the
 * state the host populated) and one after the last cycle, plus the final sha256 of each queue
 * (== hashlib.sha256 over the queue's records, each zero-padded to a multiple of 64 bytes).  The batch must have been
 * populated before its first run (the consumer starts from the pre-run state).  With it a host loop moves snapshots +
 * digests + the query logs across PCIe instead of every row. */
int32_t zkb_consume(ZkbBatch* b, uint32_t cycles_per_snapshot, void* cuda_stream);
int32_t zkb_snapshot_counts(ZkbBatch* b, uint32_t vm_lo, uint32_t vm_hi, uint32_t* counts_out);
int32_t zkb_read_snapshots(ZkbBatch* b, uint32_t vm, void* dst, uint64_t max_bytes, uint64_t* n_bytes);
/* digests_out: (vm_hi - vm_lo) x 3 x 32 bytes: sha256 of the memory, log and decommitment queue of every VM */
int32_t zkb_read_queue_digests(ZkbBatch* b, uint32_t vm_lo, uint32_t vm_hi, uint8_t* digests_out);
/* everything zkb_consume produced, packed, into (pinned) host memory on cuda_stream: u64 offsets[n_vms + 1] (first
 * snapshot of every VM), the snapshots VM-major, then n_vms x 3 x 8 u32 final digest words.  host_dst == NULL: size only. */
int32_t zkb_fetch_consumed_async(ZkbBatch* b, void* host_dst, uint64_t host_capacity, uint64_t* n_bytes, void* cuda_stream);

/* ---- post-processing (the step right after the path) ---------------------------------- */
/* Rebuilds on the device, per VM, what the reference's backends hold after the run:
 *   ZKB_FLAT_STORAGE_HISTORY  InMemoryStorage::flatten_and_net_history().0   (src/testing/storage.rs:34-76)
 *   ZKB_FLAT_EVENT_HISTORY    InMemoryEventSink::flatten().0                 (src/reference_impls/event_sink.rs:66-131)
 *   ZKB_FLAT_NET_EVENTS / ZKB_FLAT_NET_L1_MESSAGES   .1 / .2 of the same (as the LogQuery each EventMessage projects)
 * i.e. the chronological query logs including the rollback queries finish_frame(panicked) appends in reverse
 * (storage.rs:156-180), and the never-rolled-back events in timestamp order. Records are ZkbLogQueryRec.
 * Only ended VMs are flattened (the reference asserts frames_stack.len() == 1); others report status 1. */
enum ZkbFlatKind { ZKB_FLAT_STORAGE_HISTORY = 0, ZKB_FLAT_EVENT_HISTORY = 1, ZKB_FLAT_NET_EVENTS = 2, ZKB_FLAT_NET_L1_MESSAGES = 3 };
int32_t zkb_flatten_logs(ZkbBatch* b, void* cuda_stream);
int32_t zkb_flat_counts(ZkbBatch* b, uint32_t kind, uint32_t vm_lo, uint32_t vm_hi, uint32_t* counts_out, uint32_t* status_out);
int32_t zkb_read_flat(ZkbBatch* b, uint32_t vm, uint32_t kind, void* dst, uint64_t max_bytes, uint64_t* n_bytes);

/* Per-slot grouping of log queries = the second half of flatten_and_net_history, `.1`: every query filed under its slot
 * (shard_id, address, key), history order kept inside a slot (src/testing/storage.rs:50-73).  One stable radix sort on the
 * device (K11, csrc/logsort.cuh) by key = group << 44 | hash44(slot), then a gather of the 128-byte records:
 *   d_recs          n_records ZkbLogQueryRec in device memory (any origin: a batch's LOG stream, an exchange's output)
 *   d_group_of      optional u32 per record (20 bits used): records are grouped by it FIRST (e.g. the VM index, so that
 *                   every VM keeps its own slot map as the reference's per-VM InMemoryStorage does); NULL = one global map
 *   d_sorted_out    the records, ordered by (group, slot hash, input position)
 *   d_boundary_out  one byte per output record: 1 = first query of its slot (slots of equal hash are told apart by their
 *                   full identity); *n_groups_out = number of slots.
 * Slots come out in hash order (the reference's HashMap has no order at all); tests compare slot by slot. */
int32_t zkb_sort_log_queries(int32_t device, const void* d_recs, uint64_t n_records, const uint32_t* d_group_of, void* d_sorted_out,
                             uint8_t* d_boundary_out, uint64_t* n_groups_out, void* cuda_stream);
/* flatten_and_net_history().1 of EVERY VM of the batch: the storage histories (ZKB_FLAT_STORAGE_HISTORY, incl. rollback
 * queries) grouped per VM and slot.  offsets_out[n_vms + 1] = first RECORD of each VM in the output; host_sorted_out ==
 * NULL only fills offsets_out (offsets_out[n_vms] records are needed).  Runs zkb_flatten_logs first if it has not run. */
int32_t zkb_net_storage_history(ZkbBatch* b, void* host_sorted_out, uint64_t host_capacity, uint8_t* host_boundary_out, uint64_t* offsets_out,
                                uint64_t* n_slots_out, void* cuda_stream);

/* ---- measurement helpers -------------------------------------------------------------------------------- */
/* K14 -- integer-pipe micro-benchmarks (row j3; BASELINE north_star: "integer-pipe utilisation for the U256 ALU against
 * sm_100a peak").  op: 0 IMAD (mad.lo.u32), 1 IADD3 (add.u32), 2 LOP3 -- the MEASURED peaks, thread-level instructions
 * per second at 64 warps per SM, 8 independent chains per thread; 3 u256 add, 4 sub, 5 full 256x256->512 mul, 6 div_mod
 * (256-bit by 128-bit), 7 shl -- the octet-distributed primitives of csrc/u256.cuh that replace ethereum_types::U256 in
 * /root/reference/src/opcodes/execution/{add,sub,mul,div,shift}.rs, 256-bit operations per second.  Best of three timed
 * launches (CUDA events).  tools/alu_microbench.py prints the table and the utilisation figures. */
int32_t zkb_alu_microbench(int32_t device, uint32_t op, uint32_t iters, double* ops_per_second, float* kernel_ms);

/* ---- bytecode ingestion: the step right BEFORE the path ----------------------------- */
/* Versioned code hashes of n bytecodes, computed on the GPU: byte 0 = version (1), byte 1 = marker (0 at rest,
 * 1 being constructed), bytes 2..3 = length in 32-byte words (big-endian u16), bytes 4..31 = sha256(code)[4..32] --
 * the ContractCodeSha256 layout far_call parses (src/opcodes/execution/far_call.rs:169-252) and the key under which
 * SimpleDecommitter::populate files the code (src/reference_impls/decommitter.rs:23-28).
 * words_be: all bytecodes back to back, 32-byte big-endian words; offsets_words[n + 1]: word offset of each bytecode
 * (offsets_words[n] = total); hashes_be_out: n x 32 bytes.  Lengths above 65 535 words are rejected. */
int32_t zkb_hash_bytecodes(int32_t device, const uint8_t* words_be, const uint64_t* offsets_words, uint32_t n, uint8_t marker,
                           uint8_t* hashes_be_out);
/* hash (at-rest marker) + zkb_load_bytecode of every bytecode: populate() for callers that hold code, not hashes */
int32_t zkb_ingest_bytecodes(ZkbBatch* b, const uint8_t* words_be, const uint64_t* offsets_words, uint32_t n, uint8_t* hashes_be_out);

/* ---- multi-GPU concat over peer memory ------------------------------------------------------ */
/* One-sided push of a packed stream into another GPU's memory over NVLink: a grid-stride 16-byte copy kernel of
 * n_ctas small CTAs (128 threads, no shared memory) whose stores land in `dst_peer`, a device pointer of ANOTHER GPU
 * that the caller has mapped into this process (CUDA IPC) -- peer access from `device` to the owner of dst_peer is
 * enabled on first use.  The CTAs are small enough to co-reside with the persistent interpreter CTAs, so the push of
 * pass k runs underneath the launch of pass k + 1 and the receiving GPU spends no SM on it.  src / dst must be 8-byte
 * aligned (16-byte aligned pointers get 16-byte transfers); a tail shorter than one vector is copied bytewise. */
int32_t zkb_peer_push_async(int32_t device, int32_t peer_device, const void* src, void* dst_peer, uint64_t n_bytes, uint32_t n_ctas,
                            void* cuda_stream);
/* The sink of such pushes: device memory on `device` with a 64-byte CUDA IPC handle that the other ranks (one process per
 * GPU) open with THEIR device current, which is what maps it for peer access from their kernels. */
int32_t zkb_peer_sink_create(int32_t device, uint64_t n_bytes, void** dptr_out, uint8_t ipc_handle_out[64]);
int32_t zkb_peer_sink_open(int32_t device, const uint8_t ipc_handle[64], void** dptr_out);
int32_t zkb_peer_sink_close(int32_t device, void* dptr, uint32_t owner);   /* owner != 0: cudaFree, else cudaIpcCloseMemHandle */

/* ---- multi-GPU stream exchange over NCCL, from C (SURVEY §8b `zkb_gather`, §8e) ---------------------------------------
 * One process per GPU.  Rank 0 calls zkb_comm_unique_id, the host passes the 128 bytes to every rank over its own channel
 * (MPI, a socket, torch.distributed ...), every rank calls zkb_comm_create (collective).  NCCL is dlopen'ed on first use
 * (libnccl.so.2): no link-time dependency.  world <= 8 (one NVSwitch node). */
typedef struct ZkbComm ZkbComm;
int32_t zkb_comm_unique_id(uint8_t id_out[128]);
int32_t zkb_comm_create(int32_t device, int32_t rank, int32_t world, const uint8_t id[128], ZkbComm** out);
int32_t zkb_comm_destroy(ZkbComm* c);
/* Collective: concatenates, on rank dst_rank, every stream kind in kinds_mask (bit k = ZkbStreamKind k) of every rank's
 * batch, in rank order (VM-major inside a rank).  Sizes by one all-gather, payload by one grouped batch of
 * ncclSend / ncclRecv on cuda_stream.  *dptr_out (dst_rank only, else NULL): the concat buffer, valid until the next
 * collective on this comm; offsets_out[kind * (world + 1) + r] = byte offset of rank r's share of `kind` in it (entries of
 * kinds outside the mask are left alone).  The sink may change from call to call (dst_rank = step % world). */
int32_t zkb_gather_streams(ZkbBatch* b, ZkbComm* c, uint32_t kinds_mask, int32_t dst_rank, void** dptr_out, uint64_t* offsets_out, void* cuda_stream);
/* Collective: balanced all-to-all of the LOG streams.  Every LogQueryRec goes to rank (slot_hash64(shard, address, key)
 * >> 20) % world (csrc/logsort.cuh), so all queries of a storage slot meet on one GPU and every GPU receives ~1/world of
 * every rank's log.  *dptr_out: this rank's share (device memory, valid until the next collective on this comm), ordered
 * by (source rank, VM, position in the VM's log); src_offsets_out[world + 1]: first RECORD of every source rank. */
int32_t zkb_exchange_logs(ZkbBatch* b, ZkbComm* c, void** dptr_out, uint64_t* n_records_out, uint64_t* src_offsets_out, void* cuda_stream);
/* Both of the above in ONE collective step: one size all-gather and one host synchronisation in front of all transfers, which
 * are then left in flight on cuda_stream.  What a multi-GPU host loop calls once per pass. */
int32_t zkb_exchange_step(ZkbBatch* b, ZkbComm* c, uint32_t gather_kinds_mask, int32_t dst_rank, void** share_out, uint64_t* n_share_out,
                          uint64_t* src_offsets_out, void** concat_out, uint64_t* concat_offsets_out, void* cuda_stream);
/* The same step ONE-SIDED over NVLink peer memory (receive buffers of all ranks mapped through CUDA IPC at the first call,
 * which is collective): the LOG partition and the concat of the small streams (gather_kinds_mask within DECOMMIT | FRAME |
 * REFUND) are written straight into the destination GPUs' memory by 128-thread kernels that fit next to the persistent
 * interpreter CTA -- no NCCL on the data path, no SM kept free for it, and NO host synchronisation: record counts are read
 * on the device, so the caller orders cuda_stream behind the batch's launch with an event and queues the next pass at once.
 * *step_out: the tag under which zkb_push_result finds this step's data. */
int32_t zkb_push_step(ZkbBatch* b, ZkbComm* c, uint32_t gather_kinds_mask, int32_t dst_rank, uint64_t* step_out, void* cuda_stream);
/* Host: waits until every source rank's pushes of `step` into THIS rank have landed (mailbox tags), then reports, per source
 * rank s, its LOG records for this rank (share_ptrs_out[s], share_records_out[s]) and -- on that step's sink -- the gathered
 * streams (concat_ptrs_out[s * 6 + kind], concat_bytes_out[s * 6 + kind]).  Device pointers, valid until the next step that
 * targets them; timeout_ms = 0 checks once. */
int32_t zkb_push_result(ZkbBatch* b, ZkbComm* c, uint64_t step, uint32_t timeout_ms, void** share_ptrs_out, uint64_t* share_records_out,
                        void** concat_ptrs_out, uint64_t* concat_bytes_out);
/* Makes cuda_stream wait until the LAST collective on `c` has read everything it needs from its batch (its pack kernels):
 * a host loop runs the collectives of pass k on a side stream, orders only this before the restore / launch of pass k + 1
 * on the main stream, and leaves the NCCL transfers in flight underneath that launch. */
int32_t zkb_comm_wait_packed(ZkbComm* c, void* cuda_stream);

/* ---- checkpoint / accounting ---------------------------------------------------------------------- */
/* VmLocalState (+ backends) is a plain cloneable value in the reference (vm_state/mod.rs:53): snapshot keeps a
 * device-side copy of every mutable per-VM array, restore puts it back (asynchronously on `cuda_stream`). */
int32_t zkb_snapshot(ZkbBatch* b);
int32_t zkb_restore(ZkbBatch* b, void* cuda_stream);
/* host<->device bytes moved by this batch's API calls so far (optionally reset the counters) */
int32_t zkb_transfer_stats(ZkbBatch* b, uint64_t* h2d_bytes, uint64_t* d2h_bytes, uint32_t reset);

#ifdef __cplusplus
}
#endif
#endif
