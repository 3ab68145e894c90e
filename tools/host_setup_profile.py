#!/usr/bin/env python
"""Where the host-side time of one end-to-end sub-batch goes (bench.py e2e loop, thread A): wall time of every C-ABI call
behind GpuVmBatch.reset + Workload.setup + run, per call name, over a few repetitions.  Usage (GPU box):
    python tools/host_setup_profile.py [n_vms]"""
import collections
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from era_zk_evm_b200 import GpuVmBatch, workloads  # noqa: E402


class Timed:
    def __init__(self, inner, acc):
        object.__setattr__(self, "_inner", inner)
        object.__setattr__(self, "_acc", acc)

    def __getattr__(self, name):
        v = getattr(self._inner, name)
        if not callable(v):
            return v

        def call(*a, **k):
            t0 = time.perf_counter()
            try:
                return v(*a, **k)
            finally:
                self._acc[name] += time.perf_counter() - t0
        return call


def main():
    import torch
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 14208
    w = workloads.Erc20(n_transfers=8)
    ids = np.arange(n, dtype=np.uint64)
    b = GpuVmBatch(w.config(n))
    acc = collections.defaultdict(float)
    tb = Timed(b, acc)
    reps, total = 6, 0.0
    for r in range(reps + 1):
        if r == 1:
            acc.clear()
            total = 0.0
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        tb.reset()
        w.setup(tb, ids)
        tb.run(sync=False)
        total += time.perf_counter() - t0
        b.sync()
    print(f"{n} VMs, {reps} repetitions: reset + setup + run (launch only) = {total / reps * 1e3:.2f} ms per sub-batch")
    for k, v in sorted(acc.items(), key=lambda kv: -kv[1]):
        print(f"  {k:28s} {v / reps * 1e3:8.3f} ms")
    print(f"  (python glue)                {(total - sum(acc.values())) / reps * 1e3:8.3f} ms")


if __name__ == "__main__":
    main()
