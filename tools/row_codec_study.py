#!/usr/bin/env python
"""Feasibility data for a lossless TRANSPORT encoding of the witness streams (end-to-end lever): how many
of the bytes that cross PCIe are zero / unchanged from the previous record?  Runs the ERC-20 workload on the CPU oracle
(test infrastructure) and reports, per stream, the share of all-zero 8-byte words and -- for cycle rows -- of 8-byte words
equal to the same word of the previous row of the same VM.  Usage: python tools/row_codec_study.py [n_vms]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from era_zk_evm_b200 import records, workloads  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    w = workloads.Erc20(n_transfers=8)
    b = oracle.OracleBatch(w.config(n))
    w.setup(b, list(range(n)))
    b.run_threads(0, 0)
    total_raw = total_enc = 0
    for kind in range(records.N_STREAMS):
        raw = zero = same = 0
        for vm in range(n):
            rec = b.read_stream(vm, kind)
            if len(rec) == 0:
                continue
            words = rec.view(np.uint8).reshape(len(rec), -1).view(np.uint64)          # [records, 8-byte words]
            raw += words.size
            z = words == 0
            zero += int(z.sum())
            if kind == records.STREAM_ROWS and len(rec) > 1:
                eq = words[1:] == words[:-1]
                same += int((eq & ~z[1:]).sum())
        if raw == 0:
            continue
        # encoding model: one presence bit per 8-byte word (+ for rows: "same as previous row" also elided)
        kept = raw - zero - same
        enc = kept * 8 + raw // 8
        total_raw += raw * 8
        total_enc += enc
        print(f"{records.STREAM_NAMES[kind]:9s} {raw * 8 / n:10.0f} B/VM  zero words {100 * zero / raw:5.1f} %  "
              f"unchanged (rows) {100 * same / raw:5.1f} %  -> encoded {100 * enc / (raw * 8):5.1f} % of raw")
    print(f"all streams: encoded / raw = {total_enc / total_raw:.3f}  (PCIe floor would move from 302 ms to {302 * total_enc / total_raw:.0f} ms per 65 536-VM batch)")


if __name__ == "__main__":
    main()
