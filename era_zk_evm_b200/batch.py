"""GpuVmBatch — the product entry point: n independent EraVM instances on one B200 behind the C ABI of
include/zkb.h (`libzkb.so`, hand-written sm_100a CUDA).  Stands in for the reference's
`Vec<VmState<InMemoryStorage, SimpleMemory, InMemoryEventSink, DefaultPrecompilesProcessor, SimpleDecommitter, WT>>`
plus the caller loop `while !vm.execution_has_ended() { vm.cycle(&mut tracer)? }`
(/root/reference/src/vm_state/mod.rs:157-216, cycle.rs:257).

There is NO CPU fallback: construction raises if the CUDA extension is not built or no GPU is visible.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _binding, records
from ._binding import ZkbConfig, ZkbError  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ZKB_LIB_PATH") or os.path.join(_HERE, "libzkb.so")   # override: kernel-variant experiments only
_LIB = None


def load_library() -> C.CDLL:
    """dlopen the in-tree CUDA library; never builds, never falls back."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ZkbError(f"{LIB_PATH} is missing: run `python -m era_zk_evm_b200.build` (nvcc, sm_100a). "
                           "This package has no CPU fallback.")
        _LIB = C.CDLL(LIB_PATH)
    return _LIB


def hash_bytecodes(codes, marker: int = 0, device: int = 0) -> list:
    """versioned code hashes (ContractCodeSha256 layout, far_call.rs:169-252) of `codes`, computed by the GPU kernel
    behind zkb_hash_bytecodes (the ingest step before SimpleDecommitter::populate)"""
    return _binding.hash_bytecodes(load_library(), "zkb_", codes, marker, device)


class GpuVmBatch(_binding.Batch):
    def __init__(self, cfg: ZkbConfig):
        super().__init__(load_library(), "zkb_", cfg)
        lib, vp, u32, u64 = self._lib, C.c_void_p, C.c_uint32, C.c_uint64
        lib.zkb_stream_device_view.argtypes = [vp, u32, C.POINTER(vp), C.POINTER(u64)]
        lib.zkb_fetch_stream_packed.argtypes = [vp, u32, vp, u64, vp]
        lib.zkb_pack_stream_device.argtypes = [vp, u32, C.POINTER(vp), C.POINTER(u64), vp]
        lib.zkb_pack_stream_device_async.argtypes = [vp, u32, C.POINTER(vp), C.POINTER(u64), vp]
        lib.zkb_pack_stream_device_async.restype = C.c_int32
        lib.zkb_fetch_stream_packed_async.argtypes = [vp, u32, vp, u64, vp, vp]
        lib.zkb_fetch_stream_packed_async.restype = C.c_int32
        lib.zkb_snapshot.argtypes = [vp]
        lib.zkb_restore.argtypes = [vp, vp]
        lib.zkb_transfer_stats.argtypes = [vp, C.POINTER(u64), C.POINTER(u64), u32]
        for name in ("stream_device_view", "fetch_stream_packed", "pack_stream_device", "snapshot", "restore",
                     "transfer_stats"):
            getattr(lib, "zkb_" + name).restype = C.c_int32

    def stream_device_view(self, kind: int):
        p, stride = C.c_void_p(), C.c_uint64()
        self._check(self._lib.zkb_stream_device_view(self._h, kind, C.byref(p), C.byref(stride)))
        return p.value, stride.value

    def pack_stream_device(self, kind: int, stream=None):
        p, n = C.c_void_p(), C.c_uint64()
        self._check(self._lib.zkb_pack_stream_device(self._h, kind, C.byref(p), C.byref(n), stream))
        return p.value, n.value

    def pack_stream_device_async(self, kind: int, stream=None):
        """enqueue the pack of stream `kind` on `stream` (no synchronisation); returns (device pointer, bytes)"""
        p, n = C.c_void_p(), C.c_uint64()
        self._check(self._lib.zkb_pack_stream_device_async(self._h, kind, C.byref(p), C.byref(n), stream))
        return p.value, n.value

    def fetch_stream_packed(self, kind: int, host_ptr: int | None = None, host_capacity: int = 0):
        """All VMs' records of `kind`, VM-major, as (uint8 array or None, offsets[n_vms + 1])."""
        offsets = np.zeros(self.n_vms + 1, dtype=np.uint64)
        if host_ptr is None:
            total = int(self.stream_counts(kind).astype(np.uint64).sum()) * records.RECORD_BYTES[kind]
            buf = np.empty(max(total, 1), dtype=np.uint8)
            self._check(self._lib.zkb_fetch_stream_packed(self._h, kind, buf.ctypes.data, buf.size, offsets.ctypes.data))
            return buf[:total], offsets
        self._check(self._lib.zkb_fetch_stream_packed(self._h, kind, host_ptr, host_capacity, offsets.ctypes.data))
        return None, offsets

    def fetch_stream_packed_async(self, kind: int, host_ptr: int, host_capacity: int, stream=None, offsets=None):
        """enqueue pack + D2H of stream `kind` into pinned host memory on `stream`; returns the per-VM byte offsets"""
        offsets = np.zeros(self.n_vms + 1, dtype=np.uint64) if offsets is None else offsets
        self._check(self._lib.zkb_fetch_stream_packed_async(self._h, kind, host_ptr, host_capacity, offsets.ctypes.data, stream))
        return offsets

    # -- the device-side consumer ----------------------------------------------------
    def consume(self, cycles_per_snapshot: int, stream=None):
        """per-circuit-batch VmLocalState snapshots every `cycles_per_snapshot` cycles + queue commitments, on the device"""
        lib, vp, u32, u64 = self._lib, C.c_void_p, C.c_uint32, C.c_uint64
        lib.zkb_consume.argtypes = [vp, u32, vp]
        lib.zkb_snapshot_counts.argtypes = [vp, u32, u32, vp]
        lib.zkb_read_snapshots.argtypes = [vp, u32, vp, u64, C.POINTER(u64)]
        lib.zkb_read_queue_digests.argtypes = [vp, u32, u32, vp]
        lib.zkb_fetch_consumed_async.argtypes = [vp, vp, u64, C.POINTER(u64), vp]
        self._check(lib.zkb_consume(self._h, cycles_per_snapshot, stream))

    def fetch_encoded_kinds_async(self, kinds, host_ptr: int, host_capacity: int, stream=None) -> int:
        """encoded blob of a subset of the streams (e.g. the query logs only) -> pinned host memory; returns the blob size"""
        mask = 0
        for k in kinds:
            mask |= 1 << k
        self._lib.zkb_fetch_encoded_kinds_async.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.c_void_p]
        n = C.c_uint64()
        self._check(self._lib.zkb_fetch_encoded_kinds_async(self._h, mask, host_ptr, host_capacity, C.byref(n), stream))
        return n.value

    def snapshot_counts(self) -> np.ndarray:
        out = np.zeros(self.n_vms, dtype=np.uint32)
        self._check(self._lib.zkb_snapshot_counts(self._h, 0, self.n_vms, out.ctypes.data))
        return out

    def read_snapshots(self, vm: int) -> np.ndarray:
        """uint32[n_snapshots, 198]: ZkbSnapshot records of VM `vm` (ZkbLocalState first)"""
        n = C.c_uint64()
        self._check(self._lib.zkb_read_snapshots(self._h, vm, None, 0, C.byref(n)))
        buf = np.zeros(max(n.value // 4, 1), dtype=np.uint32)
        self._check(self._lib.zkb_read_snapshots(self._h, vm, buf.ctypes.data, n.value, C.byref(n)))
        return buf[: n.value // 4].reshape(-1, 198)

    def read_queue_digests(self) -> np.ndarray:
        """uint8[n_vms, 3, 32]: sha256 of the memory / log / decommitment queue of every VM"""
        out = np.zeros((self.n_vms, 3, 32), dtype=np.uint8)
        self._check(self._lib.zkb_read_queue_digests(self._h, 0, self.n_vms, out.ctypes.data))
        return out

    def fetch_consumed_async(self, host_ptr: int, host_capacity: int, stream=None) -> int:
        n = C.c_uint64()
        self._check(self._lib.zkb_fetch_consumed_async(self._h, host_ptr, host_capacity, C.byref(n), stream))
        return n.value

    def snapshot(self):
        self._check(self._lib.zkb_snapshot(self._h))

    def restore(self, stream=None):
        self._check(self._lib.zkb_restore(self._h, stream))

    def transfer_stats(self, reset: bool = False):
        h2d, d2h = C.c_uint64(), C.c_uint64()
        self._check(self._lib.zkb_transfer_stats(self._h, C.byref(h2d), C.byref(d2h), int(reset)))
        return h2d.value, d2h.value
