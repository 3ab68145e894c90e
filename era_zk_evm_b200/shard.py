"""Multi-GPU plumbing: VM batches shard by VM index (every `VmState` owns its backends by value,
/root/reference/src/vm_state/mod.rs:157-175, so no VM ever reads another's state) and the only exchange is the
concatenation of the per-GPU witness / query streams.  One process per GPU; `torch.distributed` (NCCL over
NVLink on the GPU box, gloo in the CPU tests) carries the bytes.  Nothing on the per-cycle path communicates.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def partition(n_total: int, world: int, rank: int) -> tuple[int, int]:
    """static VM-range partition [lo, hi) of rank `rank`; ranges differ by at most one VM"""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class _DevicePtr:
    """exposes a raw device pointer (a packed stream owned by the batch) to torch without a copy"""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def device_bytes_as_tensor(ptr: int, nbytes: int, device: torch.device) -> torch.Tensor:
    if nbytes == 0:
        return torch.empty(0, dtype=torch.uint8, device=device)
    return torch.as_tensor(_DevicePtr(ptr, nbytes), device=device)


class PendingGather:
    """handle of an asynchronous gather: `wait()` returns (concatenated tensor on dst | None, rank byte offsets)"""

    def __init__(self, out, offsets, works, keep=()):
        # `works` may be SHARED between the handles of one gather_many call (one grouped batch of transfers carries
        # every stream): whichever handle is waited on first drains the list in place, the others then find it empty
        self.out, self.offsets, self._works, self._keep = out, offsets, works, keep

    def wait(self):
        while self._works:
            self._works.pop().wait()
        return self.out, self.offsets


def gather_varlen(local: torch.Tensor, dst: int = 0, group=None, async_op: bool = False):
    """Concatenates one variable-length uint8 tensor per rank on rank `dst`, in rank order.
    Returns (concatenated tensor on dst | None elsewhere, byte offsets[world + 1] on every rank); with async_op a
    PendingGather whose payload transfers (NCCL send/recv) are still in flight."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n = torch.tensor([local.numel()], dtype=torch.int64, device=local.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    if rank == dst:
        out = torch.empty(int(offsets[-1]), dtype=torch.uint8, device=local.device)
        out[offsets[rank]: offsets[rank + 1]].copy_(local)
        ops = [dist.P2POp(dist.irecv, out[offsets[r]: offsets[r + 1]], r, group) for r in range(world) if r != dst and sizes[r]]
        pending = PendingGather(out, offsets, dist.batch_isend_irecv(ops) if ops else [], keep=(local,))
    else:
        works = dist.batch_isend_irecv([dist.P2POp(dist.isend, local, dst, group)]) if local.numel() else []
        pending = PendingGather(None, offsets, works, keep=(local,))
    return pending if async_op else pending.wait()


def all_gather_varlen(local: torch.Tensor, group=None):
    """every rank gets the concatenation (one all_gather on max-padded chunks, then a compaction)"""
    world = dist.get_world_size(group)
    n = torch.tensor([local.numel()], dtype=torch.int64, device=local.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    m = max(max(sizes), 1)
    padded = torch.zeros(m, dtype=torch.uint8, device=local.device)
    padded[: local.numel()].copy_(local)
    chunks = [torch.empty(m, dtype=torch.uint8, device=local.device) for _ in range(world)]
    dist.all_gather(chunks, padded, group=group)
    offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    return torch.cat([c[:s] for c, s in zip(chunks, sizes)]), offsets


def gather_stream(batch, kind: int, dst: int = 0, group=None, device: torch.device | None = None, async_op: bool = False,
                  stream_ptr=None):
    """Packs this rank's stream `kind` (VM-major, contiguous: the pack kernel writes the NCCL send buffer directly)
    and concatenates the ranks' buffers on `dst`.  Returns (bytes tensor | None, rank byte offsets,
    per-VM record counts of this rank)."""
    counts = batch.stream_counts(kind)
    if hasattr(batch, "pack_stream_device"):
        ptr, nbytes = batch.pack_stream_device(kind, stream_ptr)
        device = device or torch.device("cuda", torch.cuda.current_device())
        local = device_bytes_as_tensor(ptr, nbytes, device)
    else:   # host-resident batches (CPU tests over gloo)
        parts = [batch.read_stream(vm, kind).view(np.uint8) for vm in range(batch.n_vms)]
        flat = np.concatenate(parts) if parts else np.zeros(0, dtype=np.uint8)
        local = torch.from_numpy(np.ascontiguousarray(flat))
    if async_op:
        return gather_varlen(local, dst, group, async_op=True), counts
    out, offsets = gather_varlen(local, dst, group)
    return out, offsets, counts


def gather_many(locals_: list, dst: int = 0, group=None, copy_stream=None, size_group=None):
    """Concatenation of SEVERAL variable-length uint8 tensors per rank with ONE size exchange and one grouped batch of
    NCCL send/recv: returns a list of PendingGather (payload transfers left in flight).  size_group: a host-side (gloo)
    group for the size exchange -- the sizes are host values already, and exchanging them on the host keeps the caller
    from blocking on its own CUDA stream (the pack kernels) just to read six integers back."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = locals_[0].device
    if size_group is not None:
        n = torch.tensor([t.numel() for t in locals_], dtype=torch.int64)
        sizes = [torch.zeros_like(n) for _ in range(world)]
        dist.all_gather(sizes, n, group=size_group)
        sizes = torch.stack(sizes).numpy()
    else:
        n = torch.tensor([t.numel() for t in locals_], dtype=torch.int64, device=dev)
        sizes = [torch.zeros_like(n) for _ in range(world)]
        dist.all_gather(sizes, n, group=group)
        sizes = torch.stack(sizes).cpu().numpy()                   # [world, k]
    ops, outs = [], []
    for j, local in enumerate(locals_):
        offsets = np.concatenate([[0], np.cumsum(sizes[:, j])]).astype(np.int64)
        if rank == dst:
            out = torch.empty(int(offsets[-1]), dtype=torch.uint8, device=dev)
            if copy_stream is None:
                out[offsets[rank]: offsets[rank + 1]].copy_(local)
            else:   # dst's own share: off the caller's stream (it only has to finish before the pack buffers are reused)
                copy_stream.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(copy_stream):
                    out[offsets[rank]: offsets[rank + 1]].copy_(local)
                out.record_stream(copy_stream)
            ops += [dist.P2POp(dist.irecv, out[offsets[r]: offsets[r + 1]], r, group) for r in range(world) if r != dst and sizes[r, j]]
        else:
            out = None
            if local.numel():
                ops.append(dist.P2POp(dist.isend, local, dst, group))
        outs.append((out, offsets))
    works = dist.batch_isend_irecv(ops) if ops else []
    # every handle carries the SAME work list (waiting twice is harmless): any one of them may be waited on first
    return [PendingGather(out, offsets, works, keep=tuple(locals_)) for j, (out, offsets) in enumerate(outs)]


class PeerSink:
    """Concat buffer on `dst` that EVERY rank maps through CUDA IPC (zkb_peer_sink_create / _open).  Each rank pushes its
    packed streams straight into its slice with zkb_peer_push_async: a copy kernel of a few small CTAs whose stores
    travel over NVLink / NVSwitch peer memory.  The CTAs fit next to the persistent interpreter CTA on an SM, so the
    push of pass k runs underneath the launch of pass k + 1, and the receiving GPU spends no SM on it (NCCL's recv
    kernels slow rank 0's own interpreter launch from 13.2 to 15.5 ms at 8 GPUs).  The only collectives left are the
    host-side size exchange and one 4-byte all_reduce that marks the pushes complete.  Double-buffered: the buffer of
    pass k stays readable on `dst` while pass k + 1 is being gathered."""

    def __init__(self, capacity_bytes: int, device: torch.device, dst: int = 0, group=None, size_group=None, n_buffers: int = 2,
                 n_ctas: int = 12):
        import ctypes as C
        from .batch import load_library
        self._C, self._lib = C, load_library()
        self._lib.zkb_peer_sink_create.argtypes = [C.c_int32, C.c_uint64, C.POINTER(C.c_void_p), C.c_void_p]
        self._lib.zkb_peer_sink_open.argtypes = [C.c_int32, C.c_void_p, C.POINTER(C.c_void_p)]
        self._lib.zkb_peer_sink_close.argtypes = [C.c_int32, C.c_void_p, C.c_uint32]
        self._lib.zkb_peer_push_async.argtypes = [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p]
        self.dst, self.group, self.size_group, self.device, self.n_ctas = dst, group, size_group, device, n_ctas
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.capacity = (int(capacity_bytes) + 255) // 256 * 256
        payload = [None]
        self._ptrs = []
        if self.rank == dst:
            handles = []
            for _ in range(n_buffers):
                p, h = C.c_void_p(), (C.c_uint8 * 64)()
                rc = self._lib.zkb_peer_sink_create(device.index, self.capacity, C.byref(p), h)
                if rc != 0:
                    raise RuntimeError(f"zkb_peer_sink_create failed with status {rc}")
                self._ptrs.append(p.value)
                handles.append(bytes(h))
            payload = [(handles, self.capacity, device.index)]
        dist.broadcast_object_list(payload, src=dst, group=group)
        handles, self.capacity, self.dst_device = payload[0]
        if self.rank != dst:
            for hb in handles:
                p, h = C.c_void_p(), (C.c_uint8 * 64).from_buffer_copy(hb)
                rc = self._lib.zkb_peer_sink_open(device.index, h, C.byref(p))
                if rc != 0:
                    raise RuntimeError(f"zkb_peer_sink_open failed with status {rc}")
                self._ptrs.append(p.value)
        self.side = torch.cuda.Stream(device=device)
        self._flag = torch.zeros(1, dtype=torch.int32, device=device)
        self._turn = 0

    def close(self):
        for p in self._ptrs:
            self._lib.zkb_peer_sink_close(self.device.index, p, 1 if self.rank == self.dst else 0)
        self._ptrs = []

    def gather_many(self, locals_: list):
        """same contract as shard.gather_many: [PendingGather]; the pushes are left in flight on the side stream"""
        n = torch.tensor([t.numel() for t in locals_], dtype=torch.int64, device="cpu" if self.size_group is not None else self.device)
        sizes = [torch.zeros_like(n) for _ in range(self.world)]
        dist.all_gather(sizes, n, group=self.size_group if self.size_group is not None else self.group)
        sizes = torch.stack(sizes).cpu().numpy()                       # [world, k]
        # every stream's concatenation starts on a 256-byte boundary of the sink; rank shares follow each other in it
        base_ptr = self._ptrs[self._turn]
        self._turn = (self._turn + 1) % len(self._ptrs)
        outs, base = [], 0
        self.side.wait_stream(torch.cuda.current_stream(self.device))   # the pack kernels
        with torch.cuda.stream(self.side):
            for j, local in enumerate(locals_):
                offsets = np.concatenate([[0], np.cumsum(sizes[:, j])]).astype(np.int64)
                total = int(offsets[-1])
                if base + total > self.capacity:
                    raise RuntimeError(f"PeerSink: {base + total} bytes exceed the sink capacity {self.capacity}")
                lo, nb = base + int(offsets[self.rank]), int(local.numel())
                if nb:
                    rc = self._lib.zkb_peer_push_async(self.device.index, self.dst_device, local.data_ptr(), base_ptr + lo, nb, self.n_ctas,
                                                       self.side.cuda_stream)
                    if rc != 0:
                        raise RuntimeError(f"zkb_peer_push_async failed with status {rc}")
                out = device_bytes_as_tensor(base_ptr + base, total, self.device) if self.rank == self.dst else None
                outs.append((out, offsets))
                base += (total + 255) // 256 * 256
            work = dist.all_reduce(self._flag, group=self.group, async_op=True)   # behind every rank's pushes
        return [PendingGather(out, offsets, [work], keep=tuple(locals_)) for j, (out, offsets) in enumerate(outs)]


# ---------------------------------------------------------------------------------------------
# the C-ABI collectives (zkb_comm_* / zkb_gather_streams / zkb_exchange_logs): NCCL driven from libzkb.so itself, so a
# host in any language reaches them; torch.distributed is used here only to hand rank 0's NCCL unique id to the others
# ---------------------------------------------------------------------------------------------
def slot_hash64(recs: np.ndarray) -> np.ndarray:
    """numpy restatement of slot_hash64 (csrc/logsort.cuh) over LOG_DTYPE records: the identity hash of a storage slot
    (shard_id, address, key) that decides which rank a query goes to and how zkb_sort_log_queries orders slots"""
    m = np.uint64(0xFF51AFD7ED558CCD)
    h = np.uint64(0x9E3779B97F4A7C15) ^ recs["shard_id"].astype(np.uint64)
    addr_words = np.ascontiguousarray(recs["address"]).view("<u4").reshape(len(recs), 5).astype(np.uint64)
    keys = recs["key"].astype(np.uint64)
    with np.errstate(over="ignore"):
        for i in range(5):
            h = (h ^ addr_words[:, i]) * m
            h ^= h >> np.uint64(32)
        for i in range(8):
            h = (h ^ keys[:, i]) * m
            h ^= h >> np.uint64(32)
    return h


def log_destination(recs: np.ndarray, world: int) -> np.ndarray:
    return ((slot_hash64(recs) >> np.uint64(20)) % np.uint64(world)).astype(np.int64)


class Comm:
    """ZkbComm: one per process / GPU.  Collective constructor (needs an initialised torch.distributed group only to
    broadcast the 128-byte NCCL unique id)."""

    def __init__(self, device: torch.device, group=None):
        import ctypes as C
        from .batch import load_library
        self._C, lib = C, load_library()
        self._lib = lib
        vp, u32, u64, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int32
        lib.zkb_comm_unique_id.argtypes = [vp]
        lib.zkb_comm_create.argtypes = [i32, i32, i32, vp, C.POINTER(vp)]
        lib.zkb_comm_destroy.argtypes = [vp]
        lib.zkb_gather_streams.argtypes = [vp, vp, u32, i32, C.POINTER(vp), vp, vp]
        lib.zkb_exchange_logs.argtypes = [vp, vp, C.POINTER(vp), C.POINTER(u64), vp, vp]
        lib.zkb_comm_wait_packed.argtypes = [vp, vp]
        lib.zkb_push_step.argtypes = [vp, vp, u32, i32, C.POINTER(u64), vp]
        lib.zkb_push_result.argtypes = [vp, vp, u64, u32, vp, vp, vp, vp]
        lib.zkb_exchange_step.argtypes = [vp, vp, u32, i32, C.POINTER(vp), C.POINTER(u64), vp, C.POINTER(vp), vp, vp]
        lib.zkb_last_error.restype = C.c_char_p
        self.rank, self.world, self.device = dist.get_rank(group), dist.get_world_size(group), device
        uid = (C.c_uint8 * 128)()
        if self.rank == 0:
            self._check(lib.zkb_comm_unique_id(uid))
        payload = [bytes(uid)]
        dist.broadcast_object_list(payload, src=0, group=group)
        uid = (C.c_uint8 * 128).from_buffer_copy(payload[0])
        self._h = C.c_void_p()
        self._check(lib.zkb_comm_create(device.index, self.rank, self.world, uid, C.byref(self._h)))

    def _check(self, rc):
        if rc != 0:
            msg = self._lib.zkb_last_error()
            raise RuntimeError(f"zkb comm call failed with status {rc}: {msg.decode() if msg else ''}")

    def close(self):
        if self._h:
            self._lib.zkb_comm_destroy(self._h)
            self._h = self._C.c_void_p()

    def wait_packed(self, stream=None):
        """`stream` waits until the last collective has finished reading the batch (its pack kernels), not its transfers"""
        self._check(self._lib.zkb_comm_wait_packed(self._h, stream))

    def gather_streams(self, batch, kinds, dst: int, stream=None):
        """collective; on `dst`: {kind: (uint8 tensor view of the concat, byte offsets[world + 1] relative to it)}, else {}"""
        C = self._C
        mask = 0
        for k in kinds:
            mask |= 1 << k
        offsets = np.zeros((6, self.world + 1), dtype=np.uint64)
        p = C.c_void_p()
        self._check(self._lib.zkb_gather_streams(batch._h, self._h, mask, dst, C.byref(p), offsets.ctypes.data, stream))
        out = {}
        if self.rank == dst:
            for k in kinds:
                lo, hi = int(offsets[k, 0]), int(offsets[k, self.world])
                out[k] = (device_bytes_as_tensor(p.value + lo, hi - lo, self.device), (offsets[k] - offsets[k, 0]).astype(np.int64))
        return out

    def exchange_step(self, batch, gather_kinds, dst: int, stream=None):
        """collective: exchange_logs + gather_streams with one size exchange; returns ((share, src offsets), {kind: (concat, offsets)})"""
        C = self._C
        mask = 0
        for k in gather_kinds:
            mask |= 1 << k
        p, n, q = C.c_void_p(), C.c_uint64(), C.c_void_p()
        src = np.zeros(self.world + 1, dtype=np.uint64)
        offsets = np.zeros((6, self.world + 1), dtype=np.uint64)
        self._check(self._lib.zkb_exchange_step(batch._h, self._h, mask, dst, C.byref(p), C.byref(n), src.ctypes.data, C.byref(q),
                                                offsets.ctypes.data, stream))
        share = (device_bytes_as_tensor(p.value, n.value * 128, self.device), src.astype(np.int64))
        got = {}
        if self.rank == dst:
            for k in gather_kinds:
                lo, hi = int(offsets[k, 0]), int(offsets[k, self.world])
                got[k] = (device_bytes_as_tensor(q.value + lo, hi - lo, self.device), (offsets[k] - offsets[k, 0]).astype(np.int64))
        return share, got

    def push_step(self, batch, gather_kinds, dst: int, stream=None) -> int:
        """one-sided exchange step over NVLink peer memory (zkb_push_step): fully asynchronous on `stream`; returns the step tag"""
        mask = 0
        for k in gather_kinds:
            mask |= 1 << k
        step = self._C.c_uint64()
        self._check(self._lib.zkb_push_step(batch._h, self._h, mask, dst, self._C.byref(step), stream))
        return step.value

    def push_result(self, batch, step: int, timeout_ms: int = 20000):
        """waits for every source's pushes of `step`; returns ([uint8 tensor per source rank: its LOG records for this rank],
        {kind: [uint8 tensor per source rank]} -- non-empty on that step's sink only)"""
        C, w = self._C, self.world
        sp, sn = (C.c_void_p * w)(), (C.c_uint64 * w)()
        cp, cn = (C.c_void_p * (w * 6))(), (C.c_uint64 * (w * 6))()
        self._check(self._lib.zkb_push_result(batch._h, self._h, step, timeout_ms, sp, sn, cp, cn))
        shares = [device_bytes_as_tensor(sp[s] or 0, int(sn[s]) * 128, self.device) for s in range(w)]
        concat = {}
        for k in range(6):
            if any(cp[s * 6 + k] for s in range(w)):
                concat[k] = [device_bytes_as_tensor(cp[s * 6 + k] or 0, int(cn[s * 6 + k]) if cp[s * 6 + k] else 0, self.device) for s in range(w)]
        return shares, concat

    def exchange_logs(self, batch, stream=None):
        """collective; returns (uint8 tensor view of this rank's share of every rank's LOG records, first record of every
        source rank [world + 1])"""
        C = self._C
        p, n = C.c_void_p(), C.c_uint64()
        src = np.zeros(self.world + 1, dtype=np.uint64)
        self._check(self._lib.zkb_exchange_logs(batch._h, self._h, C.byref(p), C.byref(n), src.ctypes.data, stream))
        return device_bytes_as_tensor(p.value, n.value * 128, self.device), src.astype(np.int64)
