#!/usr/bin/env python
"""Row j3: measured integer-pipe peaks of the GPU and the throughput of the octet-distributed U256 primitives
(zkb_alu_microbench, csrc/alubench.cuh).  Prints one JSON object; run on the GPU box:
    python tools/alu_microbench.py > gpurun_out/alu_microbench.json
Utilisation = U256 ops/s x the thread-level integer instructions ONE op needs at the minimum (the limb operations of the
schoolbook algorithm: add/sub 8 limb adds, mul 64 32x32->64 multiply-adds, shl 8 funnel shifts) / the measured peak of
the pipe those run on.  The votes / shuffles / selects around them are the overhead the figure exposes."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from era_zk_evm_b200 import load_library  # noqa: E402

OPS = ["imad", "iadd3", "lop3", "u256_add", "u256_sub", "u256_mul", "u256_divmod_256_by_128", "u256_shl"]
ITERS = [4096, 4096, 4096, 20000, 20000, 4000, 600, 20000]
MIN_LIMB_OPS = {"u256_add": ("iadd3", 8), "u256_sub": ("iadd3", 8), "u256_mul": ("imad", 64), "u256_shl": ("iadd3", 8)}


def main():
    lib = load_library()
    lib.zkb_alu_microbench.argtypes = [C.c_int32, C.c_uint32, C.c_uint32, C.POINTER(C.c_double), C.POINTER(C.c_float)]
    out = {}
    for op, (name, iters) in enumerate(zip(OPS, ITERS)):
        rate, ms = C.c_double(), C.c_float()
        rc = lib.zkb_alu_microbench(0, op, iters, C.byref(rate), C.byref(ms))
        if rc != 0:
            raise SystemExit(f"zkb_alu_microbench({name}) failed: {rc}")
        out[name] = {"ops_per_s": rate.value, "kernel_ms": ms.value, "iters": iters}
    for name, (pipe, n) in MIN_LIMB_OPS.items():
        out[name]["limb_ops_per_op"] = n
        out[name]["pipe_utilisation"] = out[name]["ops_per_s"] * n / out[pipe]["ops_per_s"]
    out["note"] = ("imad / iadd3 / lop3: thread-level instructions per second at 64 warps per SM, 8 independent chains per thread "
                   "(the measured integer peaks of this GPU); u256_*: 256-bit operations per second of the octet-distributed primitives")
    print(json.dumps(out))


if __name__ == "__main__":
    main()
