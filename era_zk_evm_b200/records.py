"""numpy views of the canonical packed witness records (include/zkb_records.h)."""
import numpy as np

STREAM_ROWS, STREAM_MEM, STREAM_LOG, STREAM_DECOMMIT, STREAM_FRAME, STREAM_REFUND = range(6)
N_STREAMS = 6
STREAM_NAMES = ["rows", "mem", "log", "decommit", "frame", "refund"]
RECORD_BYTES = [256, 48, 128, 48, 128, 8]

ROW_DTYPE = np.dtype([
    ("cycle", "<u4"), ("timestamp", "<u4"), ("raw_opcode", "<u8"),
    ("masked_variant", "<u2"), ("cond_resolved", "u1"), ("error_flags", "u1"),
    ("pc_before", "<u2"), ("pc_after", "<u2"),
    ("sp_after", "<u2"), ("flags_after", "u1"), ("bits", "u1"),
    ("ergs_after", "<u4"),
    ("src0", "<u4", 8), ("src1", "<u4", 8), ("dst0", "<u4", 8), ("dst1", "<u4", 8),
    ("callstack_depth", "<u4"), ("spent_pubdata", "<u4"), ("memory_page_counter", "<u4"),
    ("n_mem", "<u2"), ("n_log", "u1"), ("n_dfr", "u1"),
    ("context_u128", "<u4", 4),
    ("tx_number", "<u2"), ("previous_super_pc", "<u2"),
    ("ergs_per_pubdata", "<u4"), ("code_page", "<u4"), ("base_page", "<u4"),
    ("heap_bound", "<u4"), ("aux_heap_bound", "<u4"),
    ("exception_handler", "<u2"), ("frame_bits", "u1"), ("reserved0", "u1"),
    ("reserved", "<u4", 9),
])
MEM_DTYPE = np.dtype([
    ("timestamp", "<u4"), ("page", "<u4"), ("index", "<u4"),
    ("memory_type", "u1"), ("rw_flag", "u1"), ("value_is_pointer", "u1"), ("origin", "u1"),
    ("value", "<u4", 8),
])
LOG_DTYPE = np.dtype([
    ("timestamp", "<u4"), ("tx_number_in_block", "<u2"), ("aux_byte", "u1"), ("shard_id", "u1"),
    ("address", "u1", 20), ("rw_flag", "u1"), ("rollback", "u1"), ("is_service", "u1"), ("reserved", "u1"),
    ("key", "<u4", 8), ("read_value", "<u4", 8), ("written_value", "<u4", 8),
])
DECOMMIT_DTYPE = np.dtype([
    ("timestamp", "<u4"), ("memory_page", "<u4"), ("decommitted_length", "<u2"), ("is_fresh", "u1"),
    ("reserved0", "u1"), ("reserved1", "<u4"), ("hash", "<u4", 8),
])
FRAME_DTYPE = np.dtype([
    ("kind", "u1"), ("panicked", "u1"), ("prev_bound_kind", "<u2"), ("cycle", "<u4"),
    ("this_address", "u1", 20), ("msg_sender", "u1", 20), ("code_address", "u1", 20),
    ("base_memory_page", "<u4"), ("code_page", "<u4"),
    ("sp", "<u2"), ("pc", "<u2"), ("exception_handler_location", "<u2"),
    ("this_shard_id", "u1"), ("caller_shard_id", "u1"), ("ergs_remaining", "<u4"),
    ("code_shard_id", "u1"), ("is_static", "u1"), ("is_local_frame", "u1"), ("reserved1", "u1"),
    ("context_u128_value", "<u4", 4), ("heap_bound", "<u4"), ("aux_heap_bound", "<u4"),
    ("prev_ergs_remaining", "<u4"), ("prev_pc", "<u2"), ("prev_sp", "<u2"), ("prev_bound_value", "<u4"),
])
REFUND_DTYPE = np.dtype([("refund_type", "<u4"), ("refund_value", "<u4")])

DTYPES = [ROW_DTYPE, MEM_DTYPE, LOG_DTYPE, DECOMMIT_DTYPE, FRAME_DTYPE, REFUND_DTYPE]
for _d, _n in zip(DTYPES, RECORD_BYTES):
    assert _d.itemsize == _n, (_d, _n)


def limbs_to_int(limbs) -> int:
    return sum(int(x) << (32 * i) for i, x in enumerate(limbs))


def int_to_be32(v: int) -> bytes:
    return (v & ((1 << 256) - 1)).to_bytes(32, "big")


def address_bytes(v: int) -> bytes:
    return (v & ((1 << 160) - 1)).to_bytes(20, "big")
