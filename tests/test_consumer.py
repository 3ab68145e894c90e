"""the device-side consumer (zkb_consume, csrc/consume.cuh).  -m gpu: every snapshot's VmLocalState equals
the one the host replay (include/zkb_host.hpp, the C++ mirror of the reference's callback contract) hands to
start_new_execution_cycle at that cycle; the boundary counts equal the rows' per-cycle record counts; the queue commitments
equal a Python sha256 chain over the oracle's records, and the final digests equal hashlib.sha256."""
import ctypes as C
import hashlib
import os
import struct

import numpy as np
import pytest

from era_zk_evm_b200 import records, workloads
from era_zk_evm_b200._binding import ZkbFrame

import test_host_replay as HR

pytestmark = pytest.mark.gpu

_K = [0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3,
      0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da,
      0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13,
      0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070,
      0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208,
      0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2]
_IV = [0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19]
_M = 0xFFFFFFFF


def _compress(st, block):
    w = list(struct.unpack(">16I", block))
    rotr = lambda x, n: ((x >> n) | (x << (32 - n))) & _M
    for t in range(16, 64):
        s0 = rotr(w[t - 15], 7) ^ rotr(w[t - 15], 18) ^ (w[t - 15] >> 3)
        s1 = rotr(w[t - 2], 17) ^ rotr(w[t - 2], 19) ^ (w[t - 2] >> 10)
        w.append((w[t - 16] + s0 + w[t - 7] + s1) & _M)
    a, b, c, d, e, f, g, h = st
    for t in range(64):
        t1 = (h + (rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25)) + ((e & f) ^ (~e & g & _M)) + _K[t] + w[t]) & _M
        t2 = ((rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22)) + ((a & b) ^ (a & c) ^ (b & c))) & _M
        h, g, f, e, d, c, b, a = g, f, e, (d + t1) & _M, c, b, a, (t1 + t2) & _M
    return [(x + y) & _M for x, y in zip(st, [a, b, c, d, e, f, g, h])]


def _padded(rec_bytes: bytes, rec_size: int) -> bytes:
    pad = (-rec_size) % 64
    return b"".join(rec_bytes[i:i + rec_size] + bytes(pad) for i in range(0, len(rec_bytes), rec_size))


@pytest.mark.parametrize("name,kwargs,n,period,deep", [
    ("alu_loop", dict(cycles=100), 4, 16, 4),
    ("erc20", dict(n_transfers=3), 40, 64, 6),
    ("erc20", dict(n_transfers=2), 9, 1, 2),               # a snapshot at every cycle
    ("keccak", dict(n_calls=2, preimage_bytes=200), 6, 7, 3),
    ("storage", dict(n_iters=24), 10, 50, 4),
    ("mixed", dict(n_programs=12), 12 * 32, 37, 24),
])
def test_snapshots_and_queue_commitments(name, kwargs, n, period, deep, oracle_mod):
    from era_zk_evm_b200 import GpuVmBatch, load_library
    from era_zk_evm_b200.batch import LIB_PATH
    shim = HR.build_shim("zkb_", LIB_PATH)
    shim.host_expected_snapshots.restype = C.c_int
    shim.host_expected_snapshots.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(ZkbFrame), C.c_char_p, C.c_uint32, C.c_uint32, C.c_uint32,
                                             C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32), C.c_char_p, C.c_int]
    w = workloads.WORKLOADS[name](**kwargs)
    gpu = GpuVmBatch(w.config(n))
    spy = HR.InitialStateSpy(gpu)
    w.setup(spy, np.arange(n))
    orc = oracle_mod.OracleBatch(w.config(n))
    w.setup(orc, np.arange(n))
    gpu.run()
    orc.run_threads(0, 0)
    gpu.consume(period)
    counts = gpu.snapshot_counts()
    digests = gpu.read_queue_digests()
    cyc = orc.vm_status()[:, 1]
    kinds = (records.STREAM_MEM, records.STREAM_LOG, records.STREAM_DECOMMIT)
    for vm in range(n):
        c_total = int(cyc[vm])
        assert counts[vm] == (0 if c_total == 0 else (c_total - 1) // period + 1)
        # final digests == hashlib over the padded records of the ORACLE's queues (all VMs)
        for q, k in enumerate(kinds):
            raw = orc.read_stream(vm, k).tobytes()
            assert digests[vm, q].tobytes() == hashlib.sha256(_padded(raw, records.RECORD_BYTES[k])).digest(), (vm, q)
    for vm in list(range(min(n, deep))) + [n - 1]:
        snaps = gpu.read_snapshots(vm)
        want = np.zeros((len(snaps) + 2, 170), dtype=np.uint32)
        n_out, err = C.c_uint32(), C.create_string_buffer(512)
        rc = shim.host_expected_snapshots(gpu._h, n, vm, C.byref(spy.frame), spy.regs[vm].tobytes(), spy.ptr_mask, spy.fields.get(0, 8),
                                          spy.fields.get(1, 0), period, want.ctypes.data, len(want), C.byref(n_out), err, 512)
        assert rc == 0, err.value.decode()
        assert n_out.value == len(snaps)
        rows = orc.read_stream(vm, records.STREAM_ROWS)
        before = {int(r["cycle"]): i for i, r in enumerate(rows)}
        cum = np.zeros((len(rows) + 1, 3), dtype=np.int64)
        cum[1:, 0] = np.cumsum(rows["n_mem"])
        cum[1:, 1] = np.cumsum(rows["n_log"])
        cum[1:, 2] = np.cumsum(rows["n_dfr"] & 3)
        chains = []
        for q, k in enumerate(kinds):
            raw = _padded(orc.read_stream(vm, k).tobytes(), records.RECORD_BYTES[k])
            per = (records.RECORD_BYTES[k] + 63) // 64
            st, states = list(_IV), [list(_IV)]
            for i in range(0, len(raw), 64):
                st = _compress(st, raw[i:i + 64])
                if (i // 64 + 1) % per == 0:
                    states.append(st)
            chains.append(states)
        for j, s in enumerate(snaps):
            assert s[:170].tobytes() == want[j].tobytes(), f"vm {vm} snapshot {j} (cycle {s[170]}): VmLocalState differs from the host replay"
            cycle = int(s[170])
            idx = before.get(cycle, len(rows))
            assert s[171:174].tolist() == cum[idx].tolist(), (vm, j)
            for q in range(3):
                assert s[174 + 8 * q: 182 + 8 * q].tolist() == chains[q][int(cum[idx][q])], (vm, j, q)
