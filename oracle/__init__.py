"""TEST INFRASTRUCTURE: Python handle on the CPU oracle (oracle/liborc.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import os
import subprocess

from era_zk_evm_b200 import _binding

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    path = os.path.join(_HERE, "liborc.so")
    srcs = [os.path.join(_HERE, f) for f in ("zkvm_oracle.cpp", "u256.hpp", "hashes.hpp", "secp256k1.hpp")] + \
        [os.path.join(_HERE, "..", "include", "zkb.h"), os.path.join(_HERE, "..", "include", "zkb_records.h"),
         os.path.join(_HERE, "..", "era_zk_evm_b200", "csrc", "isa_tables.inc")]
    stale = force or not os.path.exists(path) or any(
        os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(path) for s in srcs)
    if stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return path


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_run_threads.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
        _LIB.orc_run_threads.restype = C.c_int32
    return _LIB


class OracleBatch(_binding.Batch):
    def __init__(self, cfg):
        super().__init__(lib(), "orc_", cfg)

    def run_threads(self, max_cycles_per_vm: int = 0, n_threads: int = 0):
        self._check(self._lib.orc_run_threads(self._h, max_cycles_per_vm, n_threads))


def hash_bytecodes(codes, marker: int = 0) -> list:
    """versioned code hashes through the oracle's restatement (orc_hash_bytecodes)"""
    return _binding.hash_bytecodes(lib(), "orc_", codes, marker)


def _hash_fn(name):
    fn = getattr(lib(), name)
    fn.argtypes = [C.c_char_p, C.c_uint64, C.c_char_p]
    fn.restype = None

    def call(data: bytes) -> bytes:
        out = C.create_string_buffer(32)
        fn(data, len(data), out)
        return out.raw
    return call


def keccak256(data: bytes) -> bytes:
    return _hash_fn("orc_keccak256")(data)


def sha3_256(data: bytes) -> bytes:
    return _hash_fn("orc_sha3_256")(data)


def sha256(data: bytes) -> bytes:
    return _hash_fn("orc_sha256")(data)


def u256_op(op: int, a: int, b: int):
    fn = lib().orc_u256_op
    fn.argtypes = [C.c_uint32, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p]
    fn.restype = None
    o0, o1, fl = C.create_string_buffer(32), C.create_string_buffer(32), C.create_string_buffer(1)
    fn(op, a.to_bytes(32, "big"), b.to_bytes(32, "big"), o0, o1, fl)
    return int.from_bytes(o0.raw, "big"), int.from_bytes(o1.raw, "big"), fl.raw[0]


def keccak_precompile_harness(data: bytes, unalignment: int):
    fn = lib().orc_keccak_precompile_harness
    fn.argtypes = [C.c_char_p, C.c_uint32, C.c_uint32, C.c_char_p, C.POINTER(C.c_uint32)]
    fn.restype = C.c_int32
    out = C.create_string_buffer(32)
    n = C.c_uint32()
    rc = fn(data, len(data), unalignment, out, C.byref(n))
    assert rc == 0
    return out.raw, n.value


def ecrecover(digest: bytes, r: int, s: int, v_odd: bool):
    """(ok, 20-byte address) — the oracle's restatement of the ecrecover precompile's recovery"""
    fn = lib().orc_ecrecover
    fn.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_uint32, C.c_char_p]
    fn.restype = C.c_int32
    out = C.create_string_buffer(20)
    ok = fn(digest, r.to_bytes(32, "big"), s.to_bytes(32, "big"), int(v_odd), out)
    return bool(ok), out.raw
