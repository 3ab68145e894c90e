/* Canonical packed witness records (the wire format of the batched witness generator).
 *
 * The reference defines no binary format: `VmWitnessTracer` callbacks receive Rust structs by value
 * (/root/reference/src/witness_trace/mod.rs:11-72).  These records are the build's serialisation of exactly
 * those callback payloads, per VM, in the reference's program order (SURVEY.md §8b "ordering contract", §8d).
 * All scalars little-endian; U256 = 8 x u32 little-endian limbs (== 4 x u64 LE limbs, the layout of
 * ethereum_types::U256 used at src/opcodes/execution/ptr.rs:82); addresses = 20 big-endian bytes.
 *
 * Streams per VM (each a flat array of fixed-size records):
 *   ZKB_STREAM_ROWS      CycleRow        256 B   one per VmState::cycle()      (cycle.rs:257-429)
 *   ZKB_STREAM_MEM       MemoryQueryRec   48 B   add_memory_query + precompile mem witness
 *   ZKB_STREAM_LOG       LogQueryRec     128 B   add_log_query                 (helpers.rs:138-162,196-210)
 *   ZKB_STREAM_DECOMMIT  DecommitRec      48 B   add_decommittment             (helpers.rs:164-194)
 *   ZKB_STREAM_FRAME     FrameRec        128 B   start/finish_execution_context(helpers.rs:225-264)
 *   ZKB_STREAM_REFUND    RefundRec         8 B   record_refund_for_query       (helpers.rs:119-136)
 */
#ifndef ZKB_RECORDS_H
#define ZKB_RECORDS_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum ZkbStreamKind {
  ZKB_STREAM_ROWS = 0,
  ZKB_STREAM_MEM = 1,
  ZKB_STREAM_LOG = 2,
  ZKB_STREAM_DECOMMIT = 3,
  ZKB_STREAM_FRAME = 4,
  ZKB_STREAM_REFUND = 5,
  ZKB_N_STREAMS = 6
};

#define ZKB_ROW_BYTES 256u
#define ZKB_MEM_BYTES 48u
#define ZKB_LOG_BYTES 128u
#define ZKB_DECOMMIT_BYTES 48u
#define ZKB_FRAME_BYTES 128u
#define ZKB_REFUND_BYTES 8u

/* CycleRow.bits (byte) */
#define ZKB_ROWBIT_SRC0_PTR 0x01u   /* src0 is_pointer BEFORE erasure (cycle.rs:352-396, quirk 16) */
#define ZKB_ROWBIT_SRC1_PTR 0x02u
#define ZKB_ROWBIT_DST0_PTR 0x04u
#define ZKB_ROWBIT_DST1_PTR 0x08u
#define ZKB_ROWBIT_PENDING 0x10u    /* pending_exception after the cycle */
#define ZKB_ROWBIT_SKIP 0x20u       /* skip_cycle (execution already ended, cycle.rs:116-129) */
#define ZKB_ROWBIT_DST0_VALID 0x40u /* dst0 field holds a value written this cycle */
#define ZKB_ROWBIT_DST1_VALID 0x80u

/* CycleRow.frame_bits */
#define ZKB_FRAMEBIT_STATIC 0x01u
#define ZKB_FRAMEBIT_LOCAL 0x02u
#define ZKB_FRAMEBIT_KERNEL 0x04u

typedef struct ZkbCycleRow {
  uint32_t cycle;              /* monotonic_cycle_counter at cycle start                       w0  */
  uint32_t timestamp;          /* local_state.timestamp at cycle start                         w1  */
  uint64_t raw_opcode;         /* unmasked 64-bit encoding fetched/forced this cycle           w2-3 */
  uint16_t masked_variant;     /* decode-table index after mask_into_panic / mask_into_nop     w4  */
  uint8_t cond_resolved;       /* resolved condition (cycle.rs:193-210)                             */
  uint8_t error_flags;         /* ErrorFlags bits (helpers.rs:344-352)                              */
  uint16_t pc_before;          /*                                                              w5  */
  uint16_t pc_after;           /* pc of the (new) current frame after the cycle                     */
  uint16_t sp_after;           /*                                                              w6  */
  uint8_t flags_after;         /* bit0 LT/OF, bit1 EQ, bit2 GT (flags.rs:4-8)                       */
  uint8_t bits;                /* ZKB_ROWBIT_*                                                      */
  uint32_t ergs_after;         /* current frame ergs_remaining after the cycle                 w7  */
  uint32_t src0[8];            /* operand values the handler consumed (post swap + erasure)    w8-15 */
  uint32_t src1[8];            /*                                                              w16-23 */
  uint32_t dst0[8];            /* value given to perform_dst0_update; far_call/far ret: new r1 w24-31 */
  uint32_t dst1[8];            /* value given to perform_dst1_update; far_call: new r2         w32-39 */
  uint32_t callstack_depth;    /* callstack.depth() after                                      w40 */
  uint32_t spent_pubdata;      /*                                                              w41 */
  uint32_t memory_page_counter;/*                                                              w42 */
  uint16_t n_mem;              /* records appended to the MEM stream this cycle (mod 2^16)     w43 */
  uint8_t n_log;               /* records appended to the LOG stream this cycle                     */
  uint8_t n_dfr;               /* decommit | frame << 2 | refund << 4                               */
  uint32_t context_u128[4];    /* context_u128_register after                                  w44-47 */
  uint16_t tx_number;          /* tx_number_in_block after                                     w48 */
  uint16_t previous_super_pc;  /* after delayed changes                                             */
  uint32_t ergs_per_pubdata;   /*                                                              w49 */
  uint32_t code_page;          /* current frame after the cycle                                w50 */
  uint32_t base_page;          /*                                                              w51 */
  uint32_t heap_bound;         /*                                                              w52 */
  uint32_t aux_heap_bound;     /*                                                              w53 */
  uint16_t exception_handler;  /*                                                              w54 */
  uint8_t frame_bits;          /* ZKB_FRAMEBIT_* of the current frame after the cycle               */
  uint8_t reserved0;
  uint32_t reserved[9];        /* zero                                                         w55-63 */
} ZkbCycleRow;

/* MemoryQueryRec.origin */
#define ZKB_MEMORIGIN_VM 0u             /* witness_tracer.add_memory_query (helpers.rs:36,70,111) */
#define ZKB_MEMORIGIN_PRECOMPILE_IN 1u  /* mem_witness_in of add_precompile_call_result (helpers.rs:215-221) */
#define ZKB_MEMORIGIN_PRECOMPILE_OUT 2u /* memory_witness_out */

typedef struct ZkbMemoryQueryRec {
  uint32_t timestamp;
  uint32_t page;
  uint32_t index;
  uint8_t memory_type; /* 0 stack, 1 heap, 2 aux heap, 3 fat pointer, 4 code */
  uint8_t rw_flag;
  uint8_t value_is_pointer;
  uint8_t origin;
  uint32_t value[8];
} ZkbMemoryQueryRec;

typedef struct ZkbLogQueryRec {
  uint32_t timestamp;
  uint16_t tx_number_in_block;
  uint8_t aux_byte;
  uint8_t shard_id;
  uint8_t address[20];
  uint8_t rw_flag;
  uint8_t rollback;
  uint8_t is_service;
  uint8_t reserved;
  uint32_t key[8];
  uint32_t read_value[8];
  uint32_t written_value[8];
} ZkbLogQueryRec;

typedef struct ZkbDecommitRec {
  uint32_t timestamp;
  uint32_t memory_page;
  uint16_t decommitted_length;
  uint8_t is_fresh;
  uint8_t reserved0;
  uint32_t reserved1;
  uint32_t hash[8];
} ZkbDecommitRec;

#define ZKB_FRAMEKIND_START 1u
#define ZKB_FRAMEKIND_FINISH 2u

typedef struct ZkbFrameRec {
  uint8_t kind;                /* ZKB_FRAMEKIND_* */
  uint8_t panicked;            /* finish only */
  uint16_t prev_bound_kind;    /* far_call only: 1 = previous context's heap_bound, 2 = its aux_heap_bound was (re)set to
                                  prev_bound_value by this call's memory growth (far_call.rs:330-385); 0 = unchanged */
  uint32_t cycle;
  uint8_t this_address[20];    /* new context (start); zero for finish */
  uint8_t msg_sender[20];
  uint8_t code_address[20];
  uint32_t base_memory_page;
  uint32_t code_page;
  uint16_t sp;
  uint16_t pc;
  uint16_t exception_handler_location;
  uint8_t this_shard_id;
  uint8_t caller_shard_id;
  uint32_t ergs_remaining;
  uint8_t code_shard_id;
  uint8_t is_static;
  uint8_t is_local_frame;
  uint8_t reserved1;
  uint32_t context_u128_value[4];
  uint32_t heap_bound;
  uint32_t aux_heap_bound;
  uint32_t prev_ergs_remaining; /* previous_context fields that changed this cycle (near_call.rs:49-55) */
  uint16_t prev_pc;
  uint16_t prev_sp;
  uint32_t prev_bound_value;
} ZkbFrameRec;

typedef struct ZkbRefundRec {
  uint32_t refund_type; /* 0 = RefundType::None, 1 = RepeatedWrite */
  uint32_t refund_value;
} ZkbRefundRec;

#ifdef __cplusplus
}
#endif
#endif
