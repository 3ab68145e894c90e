#!/usr/bin/env python
"""Exercises every non-interpreter kernel once on an ERC-20 batch (pack, flatten, transport encoder, device-side consumer,
log sort / per-slot netting, bytecode hashing), so that ONE ncu run captures them all:
    ncu --set full --clock-control none -k regex:'zkb_(pack|flatten|encode_kernel|consume|logsort|hash|restore)' \
        -o gpurun_out/ncu_r02_aux python tools/aux_kernels.py [n_vms]
Also prints each step's wall time (host-synchronised) when run without a profiler."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from era_zk_evm_b200 import GpuVmBatch, hash_bytecodes, records, workloads  # noqa: E402


def timed(name, fn):
    import torch
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    print(f"{name:28s} {(time.perf_counter() - t0) * 1e3:9.2f} ms", file=sys.stderr)
    return out


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 14208
    w = workloads.Erc20(n_transfers=8)
    b = GpuVmBatch(w.config(n))
    w.setup(b, np.arange(n))
    timed("run", lambda: b.run())
    cycles, sbytes = b.totals()
    print(f"{n} VMs, {cycles} cycles, stream bytes {sbytes}", file=sys.stderr)
    for k in (records.STREAM_ROWS, records.STREAM_MEM, records.STREAM_LOG):
        timed(f"pack {records.STREAM_NAMES[k]}", lambda k=k: b.pack_stream_device(k))
    timed("flatten_logs", lambda: b.flatten_logs())
    timed("net_storage_history", lambda: b.net_storage_history())
    timed("fetch_encoded (encode + D2H)", lambda: b.fetch_encoded())
    timed("consume(256)", lambda: b.consume(256))
    rng = np.random.default_rng(7)
    codes = [rng.integers(0, 256, size=32 * int(m), dtype=np.uint8).tobytes() for m in rng.integers(1, 400, size=4096) | 1]
    timed("hash_bytecodes x4096", lambda: hash_bytecodes(codes))
    b.close()


if __name__ == "__main__":
    main()
