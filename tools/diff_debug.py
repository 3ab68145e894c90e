#!/usr/bin/env python
"""GPU-vs-oracle differential debugger: runs a workload on both and prints, per diverging VM, the first differing
record with the surrounding rows.  Usage (on a GPU box): python tools/diff_debug.py mixed '{"n_programs": 24}' 768"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

import oracle  # noqa: E402
from era_zk_evm_b200 import GpuVmBatch, isa, records, workloads  # noqa: E402
from parity_util import compare_batches, describe_row  # noqa: E402


def main():
    name = sys.argv[1]
    kwargs = json.loads(sys.argv[2]) if len(sys.argv) > 2 else {}
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 256
    w = workloads.WORKLOADS[name](**kwargs)
    ids = list(range(n))
    cfg = w.config(n)
    gpu, orc = GpuVmBatch(cfg), oracle.OracleBatch(cfg)
    w.setup(gpu, ids)
    w.setup(orc, ids)
    gpu.run()
    orc.run_threads(0, 0)
    gs, os_ = gpu.vm_status(), orc.vm_status()
    bad = [vm for vm in ids if tuple(gs[vm]) != tuple(os_[vm])]
    print(f"{len(bad)} VMs with differing status; first: {bad[:10]}")
    shown = 0
    for vm in ids:
        problems = compare_batches(gpu, orc, vms=[vm], max_report=4)
        if not problems:
            continue
        print("=" * 100)
        print(f"vm {vm}: gpu status {tuple(gs[vm])} oracle status {tuple(os_[vm])}")
        for p in problems:
            print(p)
        gr, orr = gpu.read_stream(vm, 0), orc.read_stream(vm, 0)
        k = len(gr)
        print(f"-- gpu emitted {k} rows; oracle {len(orr)}; oracle rows around the stop:")
        for i in range(max(0, k - 3), min(len(orr), k + 2)):
            print(f"   [{i}] {describe_row(orr[i])}  raw={int(orr[i]['raw_opcode']):#018x}")
        shown += 1
        if shown >= int(os.environ.get("MAX_SHOW", "4")):
            break


if __name__ == "__main__":
    main()
