#!/usr/bin/env python
"""Regenerates tests/golden/codec_v2.json: sha256 of the scalar encoder's blob (include/zkb_codec.h, format version 2) on
three small seeded workloads run on the CPU oracle.  Run only after a DELIBERATE change of the wire format."""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from era_zk_evm_b200 import workloads  # noqa: E402

CASES = (("erc20", dict(n_transfers=2), 12), ("mixed", dict(n_programs=6), 40), ("storage", dict(n_iters=24), 9))


def main():
    out = {}
    for name, kw, n in CASES:
        w = workloads.WORKLOADS[name](**kw)
        b = oracle.OracleBatch(w.config(n))
        w.setup(b, np.arange(n))
        b.run_threads(0, 1)
        blob = b.fetch_encoded()
        out[name] = {"kwargs": kw, "n_vms": n, "blob_bytes": int(blob.size), "raw_bytes": int(sum(b.totals()[1])),
                     "blob_sha256": hashlib.sha256(blob.tobytes()).hexdigest()}
    doc = {"format": "zkb_codec.h version 2",
           "note": "sha256 of the oracle encoder's blob on three small seeded workloads: pins the wire format "
                   "(tests/test_codec.py::test_wire_format_is_pinned); regenerate with tests/golden/make_codec_golden.py after a DELIBERATE format change",
           "cases": out}
    with open(os.path.join(ROOT, "tests", "golden", "codec_v2.json"), "w") as f:
        json.dump(doc, f, indent=1)


if __name__ == "__main__":
    main()
