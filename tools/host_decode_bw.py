#!/usr/bin/env python
"""Throughput of the host decoder (zkb_decode_all in libzkb.so, include/zkb_codec.h) over an oracle-encoded ERC-20 blob:
canonical GB/s written per stream and thread count.  CPU only.  Usage: python tools/host_decode_bw.py [n_vms]"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from era_zk_evm_b200 import load_library, records, workloads  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    w = workloads.Erc20(n_transfers=8)
    b = oracle.OracleBatch(w.config(n))
    w.setup(b, np.arange(n))
    b.run_threads(0, 0)
    blob = b.fetch_encoded()
    lib = load_library()
    lib.zkb_decode_all.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32]
    print(f"{n} VMs, blob {blob.size / 1e6:.1f} MB")
    for kind in (records.STREAM_ROWS, records.STREAM_MEM, records.STREAM_LOG):
        offsets = np.zeros(n + 1, dtype=np.uint64)
        lib.zkb_decode_all(blob.ctypes.data, blob.size, kind, None, 0, offsets.ctypes.data, 1)
        out = np.zeros(int(offsets[-1]), dtype=np.uint8)
        for threads in (1, os.cpu_count()):
            best = 1e9
            for _ in range(3):
                t0 = time.perf_counter()
                assert lib.zkb_decode_all(blob.ctypes.data, blob.size, kind, out.ctypes.data, out.size, offsets.ctypes.data, threads) == 0
                best = min(best, time.perf_counter() - t0)
            print(f"  {records.STREAM_NAMES[kind]:5s} {out.size / 1e6:8.1f} MB canonical, {threads:2d} threads: {out.size / best / 1e9:6.2f} GB/s")


if __name__ == "__main__":
    main()
