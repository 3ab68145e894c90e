"""CPU suite: include/zkb_host.hpp (the C++ mirror of the reference's VmState / VmWitnessTracer surface).  Every VM's
recorded streams are replayed callback by callback through a recording tracer; the replay must consume every record
exactly once, respect the reference's sub-timestamps, and the `VmLocalState` it reconstructs for
start_new_execution_cycle / end_execution_cycle must end equal to the batch's own final local state."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from era_zk_evm_b200 import records, workloads
from era_zk_evm_b200._binding import ZkbFrame, make_frame

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_shim(prefix: str, lib_path: str) -> C.CDLL:
    out_dir = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, f"libhostreplay_{prefix}.so")
    src = os.path.join(ROOT, "tests", "host_replay.cpp")
    deps = [src, os.path.join(ROOT, "include", "zkb_host.hpp"), os.path.join(ROOT, "include", "zkb.h"),
            os.path.join(ROOT, "include", "zkb_records.h"), os.path.join(ROOT, "include", "zkb_codec.h")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-fPIC", "-shared", f"-DZKB_HOST_PREFIX={prefix}", "-o", out, src,
                               lib_path, f"-Wl,-rpath,{os.path.dirname(lib_path)}"])
    lib = C.CDLL(out)
    lib.host_replay_check.restype = C.c_int
    lib.host_replay_check.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(ZkbFrame), C.c_char_p, C.c_uint32, C.c_uint32,
                                      C.c_uint32, C.POINTER(C.c_uint64 * 10), C.c_char_p, C.c_int, C.c_void_p, C.c_uint64]
    return lib


class InitialStateSpy:
    """wraps a batch during Workload.setup to remember what the host put into registers / local fields / boot frame"""

    def __init__(self, inner):
        self._b = inner
        self.regs = np.zeros((inner.n_vms, 15, 32), dtype=np.uint8)
        self.ptr_mask = 0
        self.fields = {}
        self.frame = None

    def __getattr__(self, name):
        return getattr(self._b, name)

    def set_register(self, reg, value, is_pointer=False, vm_lo=0, vm_hi=None, per_vm=False):
        if per_vm:
            self.regs[:, reg, :] = np.asarray(value, dtype=np.uint8)
        else:
            self.regs[:, reg, :] = np.frombuffer(int(value).to_bytes(32, "big"), dtype=np.uint8)
        if is_pointer:
            self.ptr_mask |= 1 << reg
        return self._b.set_register(reg, value, is_pointer, vm_lo, vm_hi, per_vm)

    def set_local_field(self, field, value, vm_lo=0, vm_hi=None):
        self.fields[field] = value
        return self._b.set_local_field(field, value, vm_lo, vm_hi)

    def push_bootloader_context(self, frame, vm_lo=0, vm_hi=None):
        self.frame = frame
        return self._b.push_bootloader_context(frame, vm_lo, vm_hi)


def replay_all(shim, batch, spy, vms):
    totals = np.zeros(10, dtype=np.uint64)
    blob = batch.fetch_encoded()          # replay_encoded (straight from the transport blob) is checked against replay
    for vm in vms:
        counts = (C.c_uint64 * 10)()
        err = C.create_string_buffer(512)
        rc = shim.host_replay_check(batch._h, batch.n_vms, vm, C.byref(spy.frame), spy.regs[vm].tobytes(), spy.ptr_mask,
                                    spy.fields.get(0, 8), spy.fields.get(1, 0), C.byref(counts), err, 512, blob.ctypes.data, blob.size)
        assert rc == 0, f"vm {vm}: {err.value.decode()}"
        c = np.array(list(counts), dtype=np.uint64)
        n = [len(batch.read_stream(vm, k)) for k in range(records.N_STREAMS)]
        assert c[0] == c[1] == n[records.STREAM_ROWS]
        assert c[2] + c[8] == n[records.STREAM_MEM]                    # VM queries + precompile witness = MEM stream
        assert c[4] == n[records.STREAM_LOG] and c[5] == n[records.STREAM_DECOMMIT] and c[3] == n[records.STREAM_REFUND]
        assert c[7] + c[9] + 1 == n[records.STREAM_FRAME]              # + the bootloader push
        totals += c
    return totals


@pytest.mark.parametrize("name,kwargs,n", [
    ("alu_loop", dict(cycles=100), 4),
    ("erc20", dict(n_transfers=3), 66),
    ("keccak", dict(n_calls=2, preimage_bytes=200), 6),
    ("storage", dict(n_iters=24), 10),
    ("mixed", dict(n_programs=12), 12 * 32),
])
def test_replay_reconstructs_the_local_state(name, kwargs, n, oracle_mod):
    shim = build_shim("orc_", os.path.join(ROOT, "oracle", "liborc.so"))
    w = workloads.WORKLOADS[name](**kwargs)
    b = oracle_mod.OracleBatch(w.config(n))
    spy = InitialStateSpy(b)
    w.setup(spy, np.arange(n))
    b.run_threads(0, 0)
    totals = replay_all(shim, b, spy, range(n))
    assert totals[0] == b.totals()[0]


def test_read_bytecode_and_bootloader_calldata_roundtrip(oracle_mod):
    """zkb_read_bytecode = SimpleDecommitter.known_hashes (decommitter.rs:10-13); zkb_set_calldata / zkb_read_calldata =
    polulate_bootloaders_calldata + dump_page_content of BOOTLOADER_CALLDATA_PAGE (memory.rs:293-344) -- through the
    oracle's mirror of the boundary here, through libzkb.so in the -m gpu suite"""
    _check_bytecode_and_calldata(oracle_mod.OracleBatch)


def _check_bytecode_and_calldata(batch_cls):
    w = workloads.Erc20(n_transfers=1)
    b = batch_cls(w.config(5))
    w.setup(b, np.arange(5))
    for h, code in w.codes.values():
        assert b.read_bytecode(h) == code
    with pytest.raises(Exception):
        b.read_bytecode(12345)
    rng = np.random.default_rng(7)
    shared = rng.integers(0, 256, size=3 * 32, dtype=np.uint8)
    b.set_calldata(shared)
    per_vm = rng.integers(0, 256, size=(2, 2 * 32), dtype=np.uint8)
    b.set_calldata(per_vm, vm_lo=1, vm_hi=3, per_vm=True)
    assert b.read_calldata(0, 0, 3).tobytes() == shared.tobytes()
    assert b.read_calldata(4, 1, 4).tobytes() == shared[32:].tobytes() + bytes(64)     # beyond the page: zero
    assert b.read_calldata(2, 0, 2).tobytes() == per_vm[1].tobytes()
    b.run() if hasattr(b, "run") else None
    b.close()


@pytest.mark.gpu
def test_read_bytecode_and_bootloader_calldata_roundtrip_gpu():
    from era_zk_evm_b200 import GpuVmBatch
    _check_bytecode_and_calldata(GpuVmBatch)
