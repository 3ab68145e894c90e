"""ctypes mirror of include/zkb.h, generic over the function prefix.

The product (`GpuVmBatch` in batch.py) instantiates it with the CUDA library and prefix ``zkb_``; the
tests instantiate it with the CPU oracle (``oracle/liborc.so``, prefix ``orc_``) which exports the same
surface.  Nothing in this module imports or loads the oracle.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import records
from .records import address_bytes, int_to_be32

N_STREAMS = records.N_STREAMS

VM_RUNNING, VM_ENDED, VM_UNKNOWN_CODE_HASH, VM_REFERENCE_PANIC = 0, 1, 2, 3
VM_CAP_STREAM, VM_CAP_STACK, VM_CAP_HEAP, VM_CAP_DEPTH, VM_CAP_STORAGE, VM_CAP_PAGES, VM_UNSUPPORTED = range(16, 23)

SCHED_AUTO, SCHED_FREE, SCHED_LOCKSTEP = 0, 1, 2
FLAT_STORAGE_HISTORY, FLAT_EVENT_HISTORY, FLAT_NET_EVENTS, FLAT_NET_L1_MESSAGES = range(4)
FIELD_MEMORY_PAGE_COUNTER, FIELD_ERGS_PER_PUBDATA, FIELD_TX_NUMBER, FIELD_TIMESTAMP = range(4)


class ZkbConfig(C.Structure):
    _fields_ = [("n_vms", C.c_uint32), ("device", C.c_int32), ("witness_mode", C.c_uint32),
                ("cap_records", C.c_uint32 * N_STREAMS), ("stack_words", C.c_uint32), ("heap_bytes", C.c_uint32),
                ("n_heap_slabs", C.c_uint32), ("max_far_depth", C.c_uint32), ("max_depth", C.c_uint32),
                ("storage_slots", C.c_uint32), ("journal_entries", C.c_uint32), ("host_mirror", C.c_uint32),
                ("schedule", C.c_uint32), ("reserved", C.c_uint32 * 3)]


class ZkbFrame(C.Structure):
    _fields_ = [("this_address", C.c_uint8 * 20), ("msg_sender", C.c_uint8 * 20), ("code_address", C.c_uint8 * 20),
                ("base_memory_page", C.c_uint32), ("code_page", C.c_uint32), ("sp", C.c_uint16), ("pc", C.c_uint16),
                ("exception_handler_location", C.c_uint16), ("reserved0", C.c_uint16), ("ergs_remaining", C.c_uint32),
                ("this_shard_id", C.c_uint8), ("caller_shard_id", C.c_uint8), ("code_shard_id", C.c_uint8),
                ("is_static", C.c_uint8), ("is_local_frame", C.c_uint8), ("reserved1", C.c_uint8 * 3),
                ("context_u128_value", C.c_uint32 * 4), ("heap_bound", C.c_uint32), ("aux_heap_bound", C.c_uint32)]


class ZkbLocalState(C.Structure):
    _fields_ = [("previous_code_word", C.c_uint32 * 8), ("previous_code_memory_page", C.c_uint32),
                ("registers", (C.c_uint32 * 8) * 15), ("register_is_pointer", C.c_uint16), ("flags", C.c_uint8),
                ("pending_exception", C.c_uint8), ("timestamp", C.c_uint32), ("monotonic_cycle_counter", C.c_uint32),
                ("spent_pubdata_counter", C.c_uint32), ("memory_page_counter", C.c_uint32),
                ("absolute_execution_step", C.c_uint32), ("current_ergs_per_pubdata_byte", C.c_uint32),
                ("tx_number_in_block", C.c_uint16), ("previous_super_pc", C.c_uint16),
                ("context_u128_register", C.c_uint32 * 4), ("callstack_depth", C.c_uint32),
                ("current_frame", ZkbFrame)]


class ZkbVmStatus(C.Structure):
    _fields_ = [("code", C.c_uint32), ("cycles", C.c_uint32)]


STORAGE_INIT_DTYPE = np.dtype([("shard_id", "u1"), ("reserved", "u1", 3), ("address", "u1", 20),
                               ("key_be", "u1", 32), ("value_be", "u1", 32)])
assert STORAGE_INIT_DTYPE.itemsize == 88


def default_config(n_vms: int, device: int = 0, max_cycles: int = 4096, witness: bool = True) -> ZkbConfig:
    cfg = ZkbConfig()
    cfg.n_vms = n_vms
    cfg.device = device
    cfg.witness_mode = int(witness)
    caps = [max_cycles, max_cycles * 2, max_cycles // 2 + 16, 64, 512, 256]
    for i, c in enumerate(caps):
        cfg.cap_records[i] = c
    cfg.stack_words = 256
    cfg.heap_bytes = 8192
    cfg.n_heap_slabs = 12
    cfg.max_far_depth = 6
    cfg.max_depth = 24
    cfg.storage_slots = 64
    cfg.journal_entries = 128
    cfg.host_mirror = 0
    return cfg


def make_frame(*, this_address=0, msg_sender=0, code_address=0, base_memory_page=8, code_page=8, sp=0, pc=0,
               exception_handler_location=0, ergs_remaining=1 << 30, is_static=False, heap_bound=0,
               aux_heap_bound=0) -> ZkbFrame:
    f = ZkbFrame()
    f.this_address[:] = list(address_bytes(this_address))
    f.msg_sender[:] = list(address_bytes(msg_sender))
    f.code_address[:] = list(address_bytes(code_address))
    f.base_memory_page, f.code_page = base_memory_page, code_page
    f.sp, f.pc, f.exception_handler_location = sp, pc, exception_handler_location
    f.ergs_remaining = ergs_remaining
    f.is_static = int(is_static)
    f.heap_bound, f.aux_heap_bound = heap_bound, aux_heap_bound
    return f


def storage_entries(items) -> np.ndarray:
    """items: iterable of (shard, address:int, key:int, value:int) -> ZkbStorageInit array."""
    items = list(items)
    arr = np.zeros(len(items), dtype=STORAGE_INIT_DTYPE)
    for i, (shard, address, key, value) in enumerate(items):
        arr[i]["shard_id"] = shard
        arr[i]["address"] = np.frombuffer(address_bytes(address), dtype=np.uint8)
        arr[i]["key_be"] = np.frombuffer(int_to_be32(key), dtype=np.uint8)
        arr[i]["value_be"] = np.frombuffer(int_to_be32(value), dtype=np.uint8)
    return arr


class ZkbError(RuntimeError):
    pass


def pack_bytecodes(codes):
    """[bytes] -> (all code words back to back, uint64 word offsets[n + 1])"""
    for c in codes:
        assert len(c) % 32 == 0
    offsets = np.zeros(len(codes) + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum([len(c) // 32 for c in codes], dtype=np.uint64)
    return b"".join(codes), offsets


def hash_bytecodes(lib, prefix: str, codes, marker: int = 0, device: int = 0) -> list:
    """versioned code hashes of `codes` through <prefix>hash_bytecodes (zkb_: the GPU kernel; orc_: the CPU oracle)"""
    fn = getattr(lib, prefix + "hash_bytecodes")
    fn.argtypes = [C.c_int32, C.c_char_p, C.c_void_p, C.c_uint32, C.c_uint8, C.c_void_p]
    fn.restype = C.c_int32
    words, offsets = pack_bytecodes(codes)
    out = np.zeros((len(codes), 32), dtype=np.uint8)
    rc = fn(device, words, offsets.ctypes.data, len(codes), marker, out.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"{prefix}hash_bytecodes failed with status {rc}")
    return [int.from_bytes(out[i].tobytes(), "big") for i in range(len(codes))]


class EncodedWitness:
    """host-side view of a blob produced by fetch_encoded*: canonical records back, byte for byte (zkb_decode_stream)"""

    def __init__(self, lib: C.CDLL, prefix: str, blob: np.ndarray):
        self._blob = np.ascontiguousarray(blob, dtype=np.uint8)
        self._ds, self._dc = getattr(lib, prefix + "decode_stream"), getattr(lib, prefix + "decode_counts")
        self._ds.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
        self._dc.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p]
        self._ds.restype = self._dc.restype = C.c_int32

    def counts(self, vm: int) -> np.ndarray:
        out = np.zeros(8, dtype=np.uint32)
        if self._dc(self._blob.ctypes.data, self._blob.size, vm, out.ctypes.data) != 0:
            raise ZkbError("decode_counts: malformed blob / VM out of range")
        return out

    def read_stream(self, vm: int, kind: int) -> np.ndarray:
        n = C.c_uint64()
        if self._ds(self._blob.ctypes.data, self._blob.size, vm, kind, None, 0, C.byref(n)) != 0:
            raise ZkbError("decode_stream: malformed blob")
        buf = np.zeros(max(n.value, 1), dtype=np.uint8)
        if self._ds(self._blob.ctypes.data, self._blob.size, vm, kind, buf.ctypes.data, n.value, C.byref(n)) != 0:
            raise ZkbError("decode_stream: malformed blob")
        from .records import DTYPES
        return buf[:n.value].view(DTYPES[kind])


class Batch:
    """Thin object wrapper over the C ABI; method names follow the reference's own
    (`populate`, `push_bootloader_context`, `execution_has_ended`, ...)."""

    def __init__(self, lib: C.CDLL, prefix: str, cfg: ZkbConfig):
        self._lib, self._p = lib, prefix
        self.cfg = cfg
        self.n_vms = cfg.n_vms
        self._h = C.c_void_p()
        self._loaded = set()
        self._declare()
        self._check(self._f("create")(C.byref(cfg), C.byref(self._h)))

    # -- plumbing -------------------------------------------------------------------------
    def _f(self, name):
        return getattr(self._lib, self._p + name)

    def _declare(self):
        vp, u32, u64 = C.c_void_p, C.c_uint32, C.c_uint64
        sigs = {
            "create": [C.POINTER(ZkbConfig), C.POINTER(vp)], "destroy": [vp], "reset": [vp],
            "load_bytecode": [vp, C.c_char_p, C.c_char_p, u32],
            "set_block_properties": [vp, C.c_char_p, C.c_uint8],
            "populate_storage": [vp, u32, u32, vp, u32, u32],
            "populate_code": [vp, u32, u32, u32, C.c_char_p],
            "push_bootloader_context": [vp, u32, u32, C.POINTER(ZkbFrame)],
            "populate_heap": [vp, u32, u32, vp, u32, u32],
            "set_register": [vp, u32, u32, u32, vp, C.c_uint8, u32],
            "set_local_field": [vp, u32, u32, u32, u32],
            "run": [vp, u32, vp], "sync": [vp],
            "last_run_ms": [vp, C.POINTER(C.c_float), C.POINTER(u32)],
            "vm_status": [vp, u32, u32, vp],
            "read_local_state": [vp, u32, C.POINTER(ZkbLocalState)],
            "stream_counts": [vp, u32, u32, u32, vp],
            "totals": [vp, C.POINTER(u64), C.POINTER(u64 * N_STREAMS)],
            "read_stream": [vp, u32, u32, vp, u64, C.POINTER(u64)],
            "read_storage": [vp, u32, C.c_uint8, C.c_char_p, C.c_char_p, vp],
            "read_heap": [vp, u32, u32, u32, vp],
            "ingest_bytecodes": [vp, C.c_char_p, vp, u32, vp],
            "flatten_logs": [vp, vp], "flat_counts": [vp, u32, u32, u32, vp, vp],
            "read_flat": [vp, u32, u32, vp, u64, C.POINTER(u64)],
            "read_bytecode": [vp, C.c_char_p, vp, u32, C.POINTER(u32)],
            "set_calldata": [vp, u32, u32, vp, u32, u32],
            "read_calldata": [vp, u32, u32, u32, vp],
            "fetch_encoded": [vp, vp, u64, C.POINTER(u64)],
            "fetch_encoded_async": [vp, vp, u64, C.POINTER(u64), vp],
            "decode_stream": [vp, u64, u32, u32, vp, u64, C.POINTER(u64)],
            "decode_counts": [vp, u64, u32, vp],
            "decode_all": [vp, u64, u32, vp, u64, vp, u32],
            "net_storage_history": [vp, vp, u64, vp, vp, C.POINTER(u64), vp],
        }
        for name, args in sigs.items():
            fn = self._f(name)
            fn.argtypes = args
            fn.restype = C.c_int32
        self._f("last_error").restype = C.c_char_p

    def _check(self, rc: int):
        if rc != 0:
            msg = self._f("last_error")()
            raise ZkbError(f"{self._p}* call failed with status {rc}: {msg.decode() if msg else ''}")

    def close(self):
        if self._h:
            self._f("destroy")(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- population (reference names) -------------------------------------------------------
    def reset(self):
        self._check(self._f("reset")(self._h))

    def load_bytecode(self, code_hash: int, code: bytes):
        assert len(code) % 32 == 0
        if code_hash in self._loaded:      # bytecodes survive reset(); re-running a workload's setup is a no-op here
            return
        self._check(self._f("load_bytecode")(self._h, int_to_be32(code_hash), code, len(code) // 32))
        self._loaded.add(code_hash)

    def read_bytecode(self, code_hash: int) -> bytes:
        """the code words filed under `code_hash` (SimpleDecommitter.known_hashes, decommitter.rs:10-13)"""
        n = C.c_uint32()
        self._check(self._f("read_bytecode")(self._h, int_to_be32(code_hash), None, 0, C.byref(n)))
        out = np.zeros(n.value * 32, dtype=np.uint8)
        self._check(self._f("read_bytecode")(self._h, int_to_be32(code_hash), out.ctypes.data, n.value, C.byref(n)))
        return out.tobytes()

    def set_calldata(self, words: np.ndarray, vm_lo=0, vm_hi=None, per_vm=False):
        """= SimpleMemory::polulate_bootloaders_calldata (memory.rs:293-298); words: uint8[..., 32 * n_words] big-endian"""
        vm_hi = self.n_vms if vm_hi is None else vm_hi
        words = np.ascontiguousarray(words, dtype=np.uint8)
        n_words = (words.size // (vm_hi - vm_lo) if per_vm else words.size) // 32
        self._check(self._f("set_calldata")(self._h, vm_lo, vm_hi, words.ctypes.data, n_words, int(per_vm)))

    def read_calldata(self, vm: int, word_lo: int, n_words: int) -> np.ndarray:
        """= dump_page_content(BOOTLOADER_CALLDATA_PAGE, word_lo .. word_lo + n_words) (memory.rs:300-344)"""
        out = np.zeros((n_words, 32), dtype=np.uint8)
        self._check(self._f("read_calldata")(self._h, vm, word_lo, n_words, out.ctypes.data))
        return out

    # -- encoded transport (include/zkb_codec.h) ---------------------------------------------------------
    def fetch_encoded(self) -> np.ndarray:
        """all six streams of every VM as one lossless blob (uint8 array); decode with EncodedWitness"""
        n = C.c_uint64()
        self._check(self._f("fetch_encoded")(self._h, None, 0, C.byref(n)))
        buf = np.zeros(n.value, dtype=np.uint8)
        self._check(self._f("fetch_encoded")(self._h, buf.ctypes.data, buf.size, C.byref(n)))
        return buf[:n.value]

    def fetch_encoded_async(self, host_ptr: int, host_capacity: int, stream=None) -> int:
        """enqueue encode + ONE D2H copy of the blob into pinned host memory on `stream`; returns the blob size"""
        n = C.c_uint64()
        self._check(self._f("fetch_encoded_async")(self._h, host_ptr, host_capacity, C.byref(n), stream))
        return n.value

    def decode_all(self, blob: np.ndarray, kind: int, n_threads: int = 0, with_offsets: bool = False):
        """host-side bulk decode of stream `kind` of every VM of a blob -> canonical records, VM-major (uint8 array)"""
        blob = np.ascontiguousarray(blob, dtype=np.uint8)
        n_vms = int(blob[:16].view(np.uint32)[2])
        offsets = np.zeros(n_vms + 1, dtype=np.uint64)
        self._check(self._f("decode_all")(blob.ctypes.data, blob.size, kind, None, 0, offsets.ctypes.data, n_threads))
        out = np.empty(int(offsets[-1]), dtype=np.uint8)
        self._check(self._f("decode_all")(blob.ctypes.data, blob.size, kind, out.ctypes.data, out.size, offsets.ctypes.data, n_threads))
        return (out, offsets) if with_offsets else out

    def net_storage_history(self):
        """flatten_and_net_history().1 of every VM (storage.rs:50-73): (records sorted by VM and slot, boundary flags --
        1 = first query of its slot --, per-VM record offsets, number of slots)"""
        from .records import LOG_DTYPE
        offsets = np.zeros(self.n_vms + 1, dtype=np.uint64)
        n_slots = C.c_uint64()
        self._check(self._f("net_storage_history")(self._h, None, 0, None, offsets.ctypes.data, C.byref(n_slots), None))
        total = int(offsets[-1])
        out = np.zeros(max(total, 1), dtype=LOG_DTYPE)
        flags = np.zeros(max(total, 1), dtype=np.uint8)
        if total:
            self._check(self._f("net_storage_history")(self._h, out.ctypes.data, out.nbytes, flags.ctypes.data, offsets.ctypes.data,
                                                        C.byref(n_slots), None))
        return out[:total], flags[:total], offsets, n_slots.value

    def ingest_bytecodes(self, codes) -> list:
        """hash (GPU, row f-4) + populate: returns the versioned code hashes as ints (= zkb_ingest_bytecodes)"""
        words, offsets = pack_bytecodes(codes)
        out = np.zeros((len(codes), 32), dtype=np.uint8)
        self._check(self._f("ingest_bytecodes")(self._h, words, offsets.ctypes.data, len(codes), out.ctypes.data))
        hashes = [int.from_bytes(out[i].tobytes(), "big") for i in range(len(codes))]
        self._loaded.update(hashes)
        return hashes

    def set_block_properties(self, default_aa_code_hash: int, zkporter_is_available: bool = False):
        self._check(self._f("set_block_properties")(self._h, int_to_be32(default_aa_code_hash), int(zkporter_is_available)))

    def populate_storage(self, entries: np.ndarray, vm_lo=0, vm_hi=None, per_vm=False):
        vm_hi = self.n_vms if vm_hi is None else vm_hi
        entries = np.ascontiguousarray(entries, dtype=STORAGE_INIT_DTYPE)
        n = entries.size // (vm_hi - vm_lo) if per_vm else entries.size
        self._check(self._f("populate_storage")(self._h, vm_lo, vm_hi, entries.ctypes.data, n, int(per_vm)))

    def populate_code(self, page: int, code_hash: int, vm_lo=0, vm_hi=None):
        vm_hi = self.n_vms if vm_hi is None else vm_hi
        self._check(self._f("populate_code")(self._h, vm_lo, vm_hi, page, int_to_be32(code_hash)))

    def push_bootloader_context(self, frame: ZkbFrame, vm_lo=0, vm_hi=None):
        vm_hi = self.n_vms if vm_hi is None else vm_hi
        self._check(self._f("push_bootloader_context")(self._h, vm_lo, vm_hi, C.byref(frame)))

    def populate_heap(self, data, vm_lo=0, vm_hi=None, per_vm=False):
        """data: bytes (broadcast) or uint8 array [n_vms_in_range, n_bytes] when per_vm."""
        vm_hi = self.n_vms if vm_hi is None else vm_hi
        arr = np.ascontiguousarray(np.frombuffer(data, dtype=np.uint8) if isinstance(data, (bytes, bytearray)) else data,
                                   dtype=np.uint8)
        n_bytes = arr.shape[-1] if per_vm else arr.size
        self._check(self._f("populate_heap")(self._h, vm_lo, vm_hi, arr.ctypes.data, n_bytes, int(per_vm)))

    def set_register(self, reg: int, value, is_pointer=False, vm_lo=0, vm_hi=None, per_vm=False):
        vm_hi = self.n_vms if vm_hi is None else vm_hi
        if per_vm:
            arr = np.ascontiguousarray(value, dtype=np.uint8)
            assert arr.shape == (vm_hi - vm_lo, 32)
        else:
            arr = np.frombuffer(int_to_be32(value), dtype=np.uint8).copy()
        self._check(self._f("set_register")(self._h, vm_lo, vm_hi, reg, arr.ctypes.data, int(is_pointer), int(per_vm)))

    def set_local_field(self, field: int, value: int, vm_lo=0, vm_hi=None):
        vm_hi = self.n_vms if vm_hi is None else vm_hi
        self._check(self._f("set_local_field")(self._h, vm_lo, vm_hi, field, value))

    # -- execution --------------------------------------------------------------------------
    def run(self, max_cycles_per_vm: int = 0, stream=None, sync=True):
        self._check(self._f("run")(self._h, max_cycles_per_vm, stream))
        if sync:
            self.sync()

    def sync(self):
        self._check(self._f("sync")(self._h))

    def last_run_ms(self):
        ms, n = C.c_float(), C.c_uint32()
        self._check(self._f("last_run_ms")(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    # -- results ----------------------------------------------------------------------------
    def vm_status(self, vm_lo=0, vm_hi=None) -> np.ndarray:
        vm_hi = self.n_vms if vm_hi is None else vm_hi
        out = np.zeros((vm_hi - vm_lo, 2), dtype=np.uint32)
        self._check(self._f("vm_status")(self._h, vm_lo, vm_hi, out.ctypes.data))
        return out

    def execution_has_ended(self) -> bool:
        return bool((self.vm_status()[:, 0] != VM_RUNNING).all())

    def read_local_state(self, vm: int) -> ZkbLocalState:
        st = ZkbLocalState()
        self._check(self._f("read_local_state")(self._h, vm, C.byref(st)))
        return st

    def stream_counts(self, kind: int, vm_lo=0, vm_hi=None) -> np.ndarray:
        vm_hi = self.n_vms if vm_hi is None else vm_hi
        out = np.zeros(vm_hi - vm_lo, dtype=np.uint32)
        self._check(self._f("stream_counts")(self._h, kind, vm_lo, vm_hi, out.ctypes.data))
        return out

    def totals(self):
        cyc = C.c_uint64()
        sb = (C.c_uint64 * N_STREAMS)()
        self._check(self._f("totals")(self._h, C.byref(cyc), C.byref(sb)))
        return cyc.value, list(sb)

    def read_stream(self, vm: int, kind: int) -> np.ndarray:
        n = C.c_uint64()
        self._check(self._f("read_stream")(self._h, vm, kind, None, 0, C.byref(n)))
        buf = np.zeros(n.value, dtype=np.uint8)
        if n.value:
            self._check(self._f("read_stream")(self._h, vm, kind, buf.ctypes.data, n.value, C.byref(n)))
        return buf.view(records.DTYPES[kind])

    # -- post-processing: the backends' flattened histories (storage.rs:34-76, event_sink.rs:66-131) --------
    def flatten_logs(self, stream=None):
        self._check(self._f("flatten_logs")(self._h, stream))

    def flat_counts(self, kind: int, vm_lo=0, vm_hi=None):
        vm_hi = self.n_vms if vm_hi is None else vm_hi
        counts, status = np.zeros(vm_hi - vm_lo, dtype=np.uint32), np.zeros(vm_hi - vm_lo, dtype=np.uint32)
        self._check(self._f("flat_counts")(self._h, kind, vm_lo, vm_hi, counts.ctypes.data, status.ctypes.data))
        return counts, status

    def read_flat(self, vm: int, kind: int) -> np.ndarray:
        n = C.c_uint64()
        self._check(self._f("read_flat")(self._h, vm, kind, None, 0, C.byref(n)))
        buf = np.zeros(n.value, dtype=np.uint8)
        if n.value:
            self._check(self._f("read_flat")(self._h, vm, kind, buf.ctypes.data, n.value, C.byref(n)))
        return buf.view(records.LOG_DTYPE)

    def read_storage(self, vm: int, shard: int, address: int, key: int) -> int:
        out = (C.c_uint8 * 32)()
        self._check(self._f("read_storage")(self._h, vm, shard, address_bytes(address), int_to_be32(key), out))
        return int.from_bytes(bytes(out), "big")

    def read_heap(self, vm: int, byte_offset: int, n_bytes: int) -> bytes:
        out = (C.c_uint8 * n_bytes)()
        self._check(self._f("read_heap")(self._h, vm, byte_offset, n_bytes, out))
        return bytes(out)
