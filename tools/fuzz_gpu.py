#!/usr/bin/env python
"""Differential fuzzing on a GPU box: RNG-generated mixed-opcode programs, CUDA batch vs CPU oracle, many seeds.
Usage: python tools/fuzz_gpu.py [n_seeds] [first_seed]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

import oracle  # noqa: E402
from era_zk_evm_b200 import GpuVmBatch, workloads  # noqa: E402
from parity_util import compare_batches  # noqa: E402


def main():
    n_seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    first = int(sys.argv[2], 0) if len(sys.argv) > 2 else 1
    bad = 0
    t0 = time.time()
    total_cycles = 0
    for seed in range(first, first + n_seeds):
        w = workloads.Mixed(n_programs=16, seed=seed, target_cycles=300 + 97 * (seed % 7))
        n = 16 * 32
        ids = list(range(n))
        cfg = w.config(n)
        cfg.schedule = 1 + seed % 2
        gpu, orc = GpuVmBatch(cfg), oracle.OracleBatch(cfg)
        w.setup(gpu, ids)
        w.setup(orc, ids)
        gpu.run()
        orc.run_threads(0, 0)
        problems = compare_batches(gpu, orc, max_report=2, allow_capacity_stops=gpu.n_vms)
        st = orc.vm_status()
        total_cycles += int(st[:, 1].sum())
        caps = int((gpu.vm_status()[:, 0] >= 16).sum())
        if problems:
            bad += 1
            print(f"seed {seed}: MISMATCH ({caps} VMs hit a capacity status)")
            for p in problems:
                print("   ", p[:1500])
        gpu.close()
        orc.close()
    print(f"fuzz: {n_seeds} seeds, {total_cycles} oracle cycles compared, {bad} seeds with mismatches, {time.time() - t0:.1f} s")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
