#!/usr/bin/env python3
"""Writes tests/golden/ck_from_label.json: the first points of DlogGroup::from_label(b"ck", n) per curve as computed by
oracle/h2c.py.  A REGRESSION fixture of the oracle, not a reference-written vector: the reference holds no point of its
commitment key; the reference cannot run here (Rust)."""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import h2c  # noqa: E402

out = {
    "generator": "tools/make_ck_golden.py (oracle/h2c.py); NOT produced by the reference",
    "label": "ck",
    "uniform_bytes_ck_1": h2c.uniform_bytes(b"ck", 2)[1].hex(),
    "from_label_ck": {str(c): [["%064x" % x, "%064x" % y] for x, y in h2c.from_label(c, b"ck", 4)] for c in range(4)},
}
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "ck_from_label.json")
with open(path, "w") as f:
    json.dump(out, f, indent=1)
    f.write("\n")
