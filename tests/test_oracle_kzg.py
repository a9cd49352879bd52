"""N3 (powers-of-tau key) / N4 (HyperKZG prover) on the CPU: the oracle's prover against the verifier's algebra for a key of known
beta, and the host build of the product's segment functions (kzg.cuh: up-sweep / down-sweep recurrence, fold, fixed-base windows)
against the oracle."""
import ctypes
import hashlib
import os
import random
import subprocess

import pytest

from oracle import kzg, sumcheck as sc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def pack(vals):
    return b"".join(int(v).to_bytes(32, "little") for v in vals)


def unpack(buf, n):
    return [int.from_bytes(buf[32 * i:32 * i + 32], "little") for i in range(n)]


@pytest.mark.parametrize("l", [1, 2, 3, 7])
def test_oracle_prover_satisfies_the_verifier_algebra(spec, l):
    p = spec.FIELD_MODULUS[0]
    rnd = random.Random(l)
    P = [rnd.randrange(p) for _ in range(1 << l)]
    x = [rnd.randrange(p) for _ in range(l)]
    beta = rnd.randrange(p)
    commit = lambda f: sum(c * pow(beta, i, p) for i, c in enumerate(f)) % p        # discrete log of commit(f) for the key beta^i g
    ch = lambda rd, msg: int.from_bytes(hashlib.sha256(repr((rd, msg)).encode()).digest(), "little") % p
    pr = kzg.prove(0, commit, P, x, ch)
    ev = sc.mle_eval(P, x, p)
    assert len(pr["com"]) == l - 1 and len(pr["v"]) == 3 and len(pr["v"][0]) == l
    assert kzg.verify_known_beta(0, None, beta, commit(P), x, ev, pr["com"], pr["v"], pr["w"], pr["u"][0], pr["q"])
    assert not kzg.verify_known_beta(0, None, beta, commit(P), x, (ev + 1) % p, pr["com"], pr["v"], pr["w"], pr["u"][0], pr["q"])
    bad_w = [pr["w"][0], (pr["w"][1] + 1) % p, pr["w"][2]]
    assert not kzg.verify_known_beta(0, None, beta, commit(P), x, ev, pr["com"], pr["v"], bad_w, pr["u"][0], pr["q"])


@pytest.fixture(scope="module", params=["emulated_gpu_limbs", "host_fast_path"])
def kzglib(request, tmp_path_factory):
    out = str(tmp_path_factory.mktemp("kzg") / f"libkzg_{request.param}.so")
    flags = ["-DLURK_HOST_EMULATE_CC"] if request.param == "emulated_gpu_limbs" else []
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", *flags, "-I",
                           os.path.join(ROOT, "lurk-beta_b200", "csrc"), "-x", "c++",
                           os.path.join(ROOT, "tests", "csrc", "kzg_host_test.cc"), "-o", out])
    lib = ctypes.CDLL(out)
    lib.kzg_test_offset.restype = ctypes.c_size_t
    lib.kzg_test_offset.argtypes = [ctypes.c_size_t, ctypes.c_int]
    return lib


@pytest.mark.parametrize("field", [0, 2])
def test_product_witness_recurrence_by_segments(kzglib, spec, field):
    """h = B / (X - u) through the up-sweep / down-sweep for lengths that exercise 1, 2 and 3 levels and ragged last segments"""
    p = spec.FIELD_MODULUS[field]
    rnd = random.Random(field)
    for n in (1, 2, 31, 32, 33, 64, 1000, 1024, 1025, 2048, 40000):
        B = [rnd.randrange(p) for _ in range(n)]
        u = rnd.randrange(p) if n != 64 else 0
        h = ctypes.create_string_buffer(32 * n)
        ev = ctypes.create_string_buffer(32)
        assert kzglib.kzg_test_witness(field, pack(B), n, pack([u]), h, ev) == 0
        assert unpack(h.raw, n) == kzg.witness_poly(B, u, p), n
        assert unpack(ev.raw, 1)[0] == kzg.poly_eval(B, u, p)


def test_product_fold_and_layout(kzglib, spec):
    p = spec.FIELD_MODULUS[0]
    rnd = random.Random(9)
    P = [rnd.randrange(p) for _ in range(64)]
    x = [rnd.randrange(p) for _ in range(6)]
    polys = kzg.fold_chain(P, x, p)
    for i in range(5):
        out = ctypes.create_string_buffer(32 * len(polys[i + 1]))
        assert kzglib.kzg_test_fold(0, pack(polys[i]), len(polys[i + 1]), pack([x[5 - i]]), out) == 0
        assert unpack(out.raw, len(polys[i + 1])) == polys[i + 1]
    off = 0
    for j in range(6):
        assert kzglib.kzg_test_offset(64, j) == off
        off += 64 >> j


@pytest.mark.parametrize("curve", [0, 2])
def test_product_fixed_base_windows(kzglib, spec, curve):
    C = spec.CURVES[curve]
    pb, q = spec.FIELD_MODULUS[C["base"]], spec.FIELD_MODULUS[C["scalar"]]
    g = spec.ec_mul(12345, C["gen"], pb)
    table = []
    base = g
    for w in range(32):
        acc = None
        for d in range(1, 256):
            acc = spec.ec_add(acc, base, pb)
            table.append(acc)
        base = spec.ec_add(acc, base, pb)
    tb = b"".join(pack(pt) for pt in table)
    rnd = random.Random(curve)
    out = ctypes.create_string_buffer(64)
    for s in [1, 2, 255, 256, q - 1, rnd.randrange(q), rnd.randrange(q), 1 << 200]:
        assert kzglib.kzg_test_fixed_mul(curve, tb, pack([s]), out) == 0
        assert tuple(unpack(out.raw, 2)) == spec.ec_mul(s, g, pb), s
