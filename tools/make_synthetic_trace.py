#!/usr/bin/env python3
"""Writes the small SYNTHETIC traces committed under tests/golden/traces/ (header flag bit 0 set): one frame's worth of slots
(14 Hash4, 6 Hash8, 1 Commitment, 3 BitDecomp; most of them dummies) with witnesses produced by the ORACLE, and one
64-term commitment.  They exercise the trace format, the loader and both replay paths; they pin nothing about the reference.
A trace written by lurk-beta itself (integration/rust/trace_export.patch) replaces them as the real pin."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import capi as oracle          # noqa: E402
import lurk_beta_b200.trace as T            # noqa: E402
from util import random_elements            # noqa: E402

out = os.path.join(ROOT, "tests", "golden", "traces")
os.makedirs(out, exist_ok=True)
rng = np.random.default_rng(2026)
slots = []
for typ, count in (("Hash4", 14), ("Hash8", 6), ("Commitment", 1), ("BitDecomp", 3)):
    a = T.SLOT_ARITY[typ]
    for k in range(count):
        dummy = rng.random() < 0.7
        pre = np.zeros((a or 1) * 32, dtype=np.uint8) if dummy else random_elements(0, a or 1, seed=int(rng.integers(1 << 30)), shape="lem" if a else "witness")
        wit = oracle.poseidon_witness_batch(0, a, pre) if a else oracle.bitdecomp_witness_batch(0, pre)
        slots.append(T.Slot(typ, dummy, wit))
T.write_slots(os.path.join(out, "slots_synthetic_bn256.bin"), 0, slots, synthetic=True)
bases = oracle.gen_bases(0, 64)
sc = random_elements(0, 64, seed=9, shape="witness")
pt = oracle.msm(0, bases, sc)
T.write_commit(os.path.join(out, "commit_synthetic_bn254.bin"), 0, bases, sc, pt[:64], not pt[64:].any(), synthetic=True)
# N3: the head of the commitment key (kind 0 = from_label on Grumpkin, the secondary circuit's Pedersen key; kind 1 = powers of tau on BN254 G1)
from oracle import h2c, kzg, spec          # noqa: E402
T.write_key(os.path.join(out, "ck_synthetic_grumpkin.bin"), 1, 0, b"ck", np.frombuffer(h2c.from_label_bytes(1, b"ck", 16), dtype=np.uint8), synthetic=True)
g = spec.ec_mul(20260924, spec.CURVES[0]["gen"], spec.FIELD_MODULUS[1])
beta = 0x1b2c3d4e5f60718293a4b5c6d7e8f90a1b2c3d4e5f60718293a4b5c6d7e8f9 % spec.FIELD_MODULUS[0]
pts = kzg.powers_of_tau(0, g, beta, 12)
lab = g[0].to_bytes(32, "little") + g[1].to_bytes(32, "little") + beta.to_bytes(32, "little")
T.write_key(os.path.join(out, "ck_synthetic_kzg_bn254.bin"), 0, 1, lab,
            np.frombuffer(b"".join(x.to_bytes(32, "little") + y.to_bytes(32, "little") for x, y in pts), dtype=np.uint8), synthetic=True)
print("wrote", sorted(os.listdir(out)))
