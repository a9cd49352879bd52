"""Shared helpers for the parity tests: seeded inputs in the shapes SURVEY.md 8(d) prescribes."""
import numpy as np

from oracle import spec

import json
import os

# reference golden digests, BN256 Fr; the fixture cites the reference test each one comes from
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_goldens.json")) as _f:
    REFERENCE = json.load(_f)
GOLDEN = {k: int(v["hex"], 16) for k, v in REFERENCE["poseidon_digests"].items()}
# ExprTag values used by the goldens (src/tag.rs): Nil=0, Cons=1, Sym=2, Fun=3, Num=4, Str=6, Char=7
TAG_SYM, TAG_NUM, TAG_STR, TAG_CHAR, TAG_NIL = 2, 4, 6, 7, 0


def random_elements(field_id, count, seed, shape="uniform"):
    """uint8 array of `count` canonical elements.
    uniform: uniform in [0, p).  lem: even positions are tags (u16), odd uniform.  witness: 40% in {0,1},
    10% < 2^16, 50% uniform."""
    p = spec.FIELD_MODULUS[field_id]
    rng = np.random.default_rng(seed)
    raw = rng.integers(0, 256, size=(count, 32), dtype=np.uint8)
    top_bits = p.bit_length() - 248
    raw[:, 31] &= (1 << (top_bits - 1)) - 1          # < 2^(bits-1) < p: uniform enough, always reduced
    if shape == "lem":
        tags = rng.integers(0, 0x3014, size=count).astype(np.uint16)
        even = np.arange(count) % 2 == 0
        raw[even] = 0
        raw[even, 0] = (tags[even] & 0xff).astype(np.uint8)
        raw[even, 1] = (tags[even] >> 8).astype(np.uint8)
    elif shape == "witness":
        u = rng.random(count)
        small = u < 0.4
        raw[small] = 0
        raw[small, 0] = rng.integers(0, 2, size=int(small.sum()), dtype=np.uint8)
        mid = (u >= 0.4) & (u < 0.5)
        raw[mid, 2:] = 0
    elif shape != "uniform":
        raise ValueError(shape)
    return raw.reshape(-1)


def ints(buf):
    b = np.ascontiguousarray(buf, dtype=np.uint8).tobytes()
    return [int.from_bytes(b[i:i + 32], "little") for i in range(0, len(b), 32)]


def pack(vals):
    return np.frombuffer(b"".join(int(v).to_bytes(32, "little") for v in vals), dtype=np.uint8).copy()
