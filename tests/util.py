"""Shared helpers for the parity tests: seeded inputs in the shapes SURVEY.md 8(d) prescribes."""
import numpy as np

from oracle import spec

GOLDEN = {   # reference golden digests, BN256 Fr
    "G1": 0x1ca5b207085f3f0f324a2e0704b18fff1cda2e2d686aa85343fea91df77bf35b,   # src/coprocessor/trie/mod.rs:932,983
    "G2": 0x0637ddaef5cd53ba6711c328952208d846222066701e10c34d3a6df7350de8aa,   # :936,992
    "G3": 0x08127a45502f5939273edd1957c8748ae39992e2a459d99f999992a842df99a5,   # :940,1001
    "G4": 0x12c2ef2ab5df25442fe23d8711bf985f02c39e83930517f7103d4bd4228c6cfb,   # :1010
    "G5": 0x2bfc4f437d5ca652511d67e06201b4fdf95c314c85ea987988746a253071bed6,   # src/lem/tests/eval_tests.rs:3868
    "G6": 0x1d501baeefe83acf0e7137180b091834f542a5059dbaf99ec82c5e19d3bb9201,   # src/lem/store.rs:1473
    "G7": 0x0df269cc1a453b80d4694fe3e54f0ff2d68bfa6a6dd6320446af03691112e89d,   # src/lem/tests/eval_tests.rs:1944
    "G8": 0x2e78db30531cf5ddd836d2b5594d2895a78c71de06abf212c4bcb0de268d4557,   # src/lem/tests/eval_tests.rs:1955
    "G10": 0x21ad1dd339f26bb824ab861dbcf110c1bcb3b7658eea4b5e84780a3b4958bf95,  # StandardTrie root after insert 123 -> 456 (eval_tests.rs:3904)
}
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
