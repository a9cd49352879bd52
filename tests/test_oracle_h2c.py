"""N3 (commitment-key generation) on the CPU: the oracle's own pins, the host build of the product's hash-to-curve templates
against the oracle, and the host-only entry points of the library.

Nothing here is pinned by the reference (no point of the key exists in it -- SURVEY.md 8(c)); what IS pinned:
SHAKE256 / BLAKE2b against hashlib, the iso-curve coefficients by the group order, the isogeny constants typed into the product
(pasta_curves' published ISOGENY_CONSTANTS) against the oracle's Velu derivation, Z by the RFC 9380 criteria, every output on its
curve, and a regression fixture written by tools/make_ck_golden.py."""
import ctypes
import hashlib
import json
import os
import random
import subprocess

import numpy as np
import pytest

from oracle import h2c

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PREFIX = "from_uniform_bytes"


def on_curve(curve, pt):
    p, b = h2c.base_modulus(curve), h2c.curve_b(curve)
    x, y = pt
    return (y * y - x * x * x - b) % p == 0


@pytest.mark.parametrize("curve", [0, 1, 2, 3])
def test_oracle_points_on_curve_and_fixture(curve):
    pts = h2c.from_label(curve, b"ck", 8)
    assert all(on_curve(curve, pt) for pt in pts)
    assert len(set(pts)) == 8
    with open(os.path.join(ROOT, "tests", "golden", "ck_from_label.json")) as f:
        fix = json.load(f)
    want = [(int(x, 16), int(y, 16)) for x, y in fix["from_label_ck"][str(curve)]]
    assert pts[:len(want)] == want
    assert h2c.uniform_bytes(b"ck", 2)[1].hex() == fix["uniform_bytes_ck_1"]


def test_ck_size_rule():
    assert h2c.ck_size(1_114_100, 911_900) == 1 << 21          # fib rc = 100
    assert h2c.ck_size(3, 5, 100) == 128 and h2c.ck_size(0, 0) == 1 and h2c.ck_size(1 << 20, 7) == 1 << 20


@pytest.mark.parametrize("curve", [2, 3])
def test_iso_curve_has_the_target_group_order(curve):
    """an isogenous curve has as many points as its target: [order] P = O for a few points of y^2 = x^3 + a x + 1265"""
    from oracle import spec
    p, a, b = h2c.base_modulus(curve), h2c.ISO_A[curve], h2c.ISO_B
    order = spec.FIELD_MODULUS[spec.CURVES[curve]["scalar"]]
    x, found = 1, 0
    while found < 3:
        g = (x * x * x + a * x + b) % p
        if h2c.is_square(g, p) and g:
            assert h2c.ec_mul(order, (x, h2c.sqrt(g, p)), a, p) is None
            found += 1
        x += 1
    # Z = -13: non-square, g(b / (Z a)) square (RFC 9380 6.6.2 conditions 1 and 4)
    Z = h2c.SSWU_Z % p
    xz = b * pow(Z * a, -1, p) % p
    assert not h2c.is_square(Z, p) and h2c.is_square((xz ** 3 + a * xz + b) % p, p)


@pytest.mark.parametrize("curve", [0, 1])
def test_svdw_z_is_what_find_z_svdw_returns(curve):
    assert h2c.svdw_z_is_valid(curve, 1)            # ctr = 1 is the first candidate find_z_svdw tries


@pytest.mark.parametrize("curve", [2, 3])
def test_isogeny_maps_iso_curve_onto_target(curve):
    p, a, b = h2c.base_modulus(curve), h2c.ISO_A[curve], h2c.ISO_B
    c = h2c.isogeny_constants(curve)
    assert c[0] * 9 % p == 1 and c[12] == p - 540
    rnd = random.Random(curve)
    pts = []
    while len(pts) < 6:
        x = rnd.randrange(p)
        g = (x ** 3 + a * x + b) % p
        if h2c.is_square(g, p):
            pts.append((x, h2c.sqrt(g, p)))
    imgs = [h2c.iso_map(curve, pt) for pt in pts]
    assert all(on_curve(curve, im) for im in imgs)
    # a group homomorphism: phi(P + Q) = phi(P) + phi(Q)
    s = h2c.ec_add(pts[0], pts[1], a, p)
    assert h2c.iso_map(curve, s) == h2c.ec_add(imgs[0], imgs[1], 0, p)


@pytest.fixture(scope="module", params=["emulated_gpu_limbs", "host_fast_path"])
def h2clib(request, tmp_path_factory):
    out = str(tmp_path_factory.mktemp("h2c") / f"libh2c_{request.param}.so")
    flags = ["-DLURK_HOST_EMULATE_CC"] if request.param == "emulated_gpu_limbs" else []
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", *flags, "-I",
                           os.path.join(ROOT, "lurk-beta_b200", "csrc"), "-x", "c++",
                           os.path.join(ROOT, "tests", "csrc", "h2c_host_test.cc"), "-o", out])
    return ctypes.CDLL(out)


def _pt(buf):
    return int.from_bytes(buf.raw[:32], "little"), int.from_bytes(buf.raw[32:64], "little")


def test_product_hashes_against_hashlib(h2clib):
    rnd = random.Random(1)
    out = ctypes.create_string_buffer(64)
    for n in (0, 1, 3, 64, 127, 128, 129, 255, 256, 257, 1000):
        d = bytes(rnd.randrange(256) for _ in range(n))
        h2clib.h2c_test_blake2b(d, n, out)
        assert out.raw == hashlib.blake2b(d).digest(), n
    for n, m, step in ((0, 32, 0), (2, 1000, 32), (135, 500, 7), (136, 272, 136), (137, 5000, 33), (500, 64, 0)):
        d = bytes(rnd.randrange(256) for _ in range(n))
        o = ctypes.create_string_buffer(m)
        h2clib.h2c_test_shake256(d, n, o, m, step)
        assert o.raw == hashlib.shake_256(d).digest(m), (n, m)


@pytest.mark.parametrize("field", [0, 1, 2, 3])
def test_product_fixed_sequence_sqrt(h2clib, spec, field):
    p = spec.FIELD_MODULUS[field]
    rnd = random.Random(field)
    o = ctypes.create_string_buffer(32)
    for x in [0, 1, 4, p - 1, 2] + [rnd.randrange(p) for _ in range(60)]:
        sq = h2clib.h2c_test_sqrt(field, x.to_bytes(32, "little"), o)
        assert sq == int(h2c.is_square(x, p)), (field, x)
        if sq:
            assert int.from_bytes(o.raw, "little") ** 2 % p == x


@pytest.mark.parametrize("curve", [0, 1, 2, 3])
def test_product_hash_to_curve_against_oracle(h2clib, curve):
    p = h2c.base_modulus(curve)
    rnd = random.Random(100 + curve)
    o64 = ctypes.create_string_buffer(64)
    # constants: computed in C++ (SVDW) / typed in (isogeny) vs computed / derived by the oracle
    if curve < 2:
        o = ctypes.create_string_buffer(128)
        assert h2clib.h2c_test_svdw_constants(curve, o) == 0
        assert [int.from_bytes(o.raw[32 * i:32 * i + 32], "little") for i in range(4)] == list(h2c.svdw_constants(curve))
    else:
        o = ctypes.create_string_buffer(14 * 32)
        assert h2clib.h2c_test_iso_constants(curve, o) == 0
        got = [int.from_bytes(o.raw[32 * i:32 * i + 32], "little") for i in range(14)]
        assert got[0] == h2c.ISO_A[curve] and got[1:] == h2c.isogeny_constants(curve)
    # hash_to_field for several message lengths and prefixes
    for prefix, ml in ((PREFIX, 0), (PREFIX, 1), (PREFIX, 31), (PREFIX, 32), (PREFIX, 40), ("z", 64), ("a-longer-domain-prefix", 33)):
        msg = bytes(rnd.randrange(256) for _ in range(ml))
        assert h2clib.h2c_test_hash_to_field(curve, prefix.encode(), msg, ml, o64) == 0
        assert list(_pt(o64)) == h2c.hash_to_field(curve, prefix, msg), (curve, prefix, ml)
    # what does not fit one BLAKE2b block is refused
    assert h2clib.h2c_test_hash_to_field(curve, b"p" * 70, bytes(32), 32, o64) == -1
    assert h2clib.h2c_test_hash_to_field(curve, PREFIX.encode(), bytes(65), 65, o64) == -1
    # the map alone, incl. the exceptional inputs u = 0, +-1
    for u in [0, 1, p - 1, 2] + [rnd.randrange(p) for _ in range(40)]:
        assert h2clib.h2c_test_map(curve, u.to_bytes(32, "little"), o64) == 0
        assert _pt(o64) == (h2c.svdw_map(curve, u) if curve < 2 else h2c.sswu_map(curve, u)), (curve, u)
    # whole points of from_label(b"ck")
    want = h2c.from_label(curve, b"ck", 24)
    for ub, e in zip(h2c.uniform_bytes(b"ck", 24), want):
        assert h2clib.h2c_test_point(curve, PREFIX.encode(), ub, 32, o64) == 0
        assert _pt(o64) == e, curve


def test_library_host_only_entry_points(L):
    rnd = random.Random(5)
    for n, m in ((0, 1), (2, 32), (2, 64 * 1024 + 5), (200, 137)):
        d = bytes(rnd.randrange(256) for _ in range(n))
        assert L.shake256(d, m) == hashlib.shake_256(d).digest(m)
    assert L.ck_size(1_114_100, 911_900) == 1 << 21 and L.ck_size(3, 5, 100) == 128 and L.ck_size(0, 0, 0) == 1
    if L._capi.lib().lurk_device_count() == 0:
        with pytest.raises(L.LurkError) as e:
            L.from_label(0, b"ck", 4)
        assert e.value.code == L._capi.ERR_NOGPU
        with pytest.raises(L.LurkError) as e:
            L.hash_to_curve_batch(0, PREFIX, np.zeros(32, dtype=np.uint8), 32)
        assert e.value.code == L._capi.ERR_NOGPU


@pytest.mark.parametrize("curve", [0, 1, 2, 3])
def test_c_port_of_from_label_equals_python_restatement(oracle, curve):
    """oracle.c: oracle_hash_to_curve_batch (the CPU baseline of the N3 row) against oracle/h2c.py -- with the host build of the CUDA
    templates that makes three implementations agreeing point for point"""
    n = 150
    assert oracle.from_label(curve, b"ck", n, nthreads=4).tobytes() == h2c.from_label_bytes(curve, b"ck", n)
    assert oracle.from_label(curve, b"", 3).tobytes() == h2c.from_label_bytes(curve, b"", 3)


def test_hash_known_answers_independent_of_hashlib(h2clib):
    """RFC 7693 Appendix A (BLAKE2b-512 of "abc") and the FIPS 202 SHAKE256 empty-message vector, as literals: pins both the product's
    implementations and the way the oracle drives hashlib (digest size, XOF)"""
    blake_abc = bytes.fromhex("ba80a53f981c4d0d6a2797b69f12f6e94c212f14685ac4b74b12bb6fdbffa2d1"
                              "7d87c5392aab792dc252d5de4533cc9518d38aa8dbf1925ab92386edd4009923")
    shake_empty = bytes.fromhex("46b9dd2b0ba88d13233b3feb743eeb243fcd52ea62b81b82b50c27646ed5762f"
                                "d75dc4ddd8c0f200cb05019d67b592f6fc821c49479ab48640292eacb3b7c4be")
    assert hashlib.blake2b(b"abc", digest_size=64).digest() == blake_abc
    assert hashlib.shake_256(b"").digest(64) == shake_empty
    out = ctypes.create_string_buffer(64)
    h2clib.h2c_test_blake2b(b"abc", 3, out)
    assert out.raw == blake_abc
    h2clib.h2c_test_shake256(b"", 0, out, 64, 0)
    assert out.raw == shake_empty


def _pasta_vectors():
    with open(os.path.join(ROOT, "tests", "golden", "pasta_hash_to_curve_vectors.json")) as f:
        return json.load(f)["vectors"]


def _affine_of(v, p):
    x, z = int(v["x"], 16), int(v["z"], 16)
    zi = pow(z, -1, p)
    ax = x * zi * zi % p
    ay = int(v["y"], 16) * zi * zi * zi % p if v["y"] else None
    return ax, ay


@pytest.mark.parametrize("v", _pasta_vectors(), ids=lambda v: v["curve"])
def test_pasta_curves_own_hash_to_curve_vectors(h2clib, v):
    """pasta_curves' unit-test vectors (Jacobian x, y, z of hash_to_curve("z.cash:test")(msg)): the oracle AND the host build of the CUDA
    templates reproduce them -- the external pin of hash_to_field (BLAKE2b XMD, DST layout), SSWU (Z = -13, iso-curve coefficients, sign rule),
    the sum on the iso-curve and the 3-isogeny.  This is what makes from_label on Pallas / Vesta 'pinned' rather than 'restated'."""
    c, msg = v["curve_id"], v["message_ascii"].encode()
    p, b = h2c.base_modulus(c), h2c.curve_b(c)
    ax, ay = _affine_of(v, p)
    got = h2c.hash_to_curve(c, v["domain_prefix"], msg)
    assert got[0] == ax and (ay is None or got[1] == ay)
    assert (got[1] ** 2 - got[0] ** 3 - b) % p == 0
    out = ctypes.create_string_buffer(64)
    assert h2clib.h2c_test_point(c, v["domain_prefix"].encode(), msg, len(msg), out) == 0
    assert _pt(out) == got
