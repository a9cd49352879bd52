"""GPU parity for N3 (commitment-key generation): `DlogGroup::from_label` / `hash_to_curve` through the C ABI
against oracle/h2c.py, bit-exact as canonical affine points; size-independent properties at the real key size (2^21 points for
fib rc = 100: every point on the curve, no duplicates, sampled points equal the oracle, chunk boundaries of the host/GPU
pipeline).  Parity with Arecibo's actual key is unpinned (tests/test_oracle_h2c.py says what is pinned instead)."""
import hashlib
import time

import numpy as np
import pytest

from oracle import h2c
from util import ints, random_elements

pytestmark = pytest.mark.gpu
CURVES = [0, 1, 2, 3]


def pts_of(buf):
    v = ints(buf)
    return list(zip(v[0::2], v[1::2]))


@pytest.mark.parametrize("curve", CURVES)
def test_from_label_matches_oracle(L, curve):
    n = 300
    got = L.from_label(curve, b"ck", n)
    assert got.tobytes() == h2c.from_label_bytes(curve, b"ck", n)
    # Montgomery output = canonical output converted (what lurk_msm_ctx_create consumes either way)
    p = h2c.base_modulus(curve)
    mont = pts_of(L.from_label(curve, b"ck", 16, fmt=L.FMT_MONTGOMERY))
    R = (1 << 256) % p
    assert [(x * R % p, y * R % p) for x, y in pts_of(got)[:16]] == mont
    # another label, empty label, n = 0 and n = 1
    assert L.from_label(curve, b"another label", 5).tobytes() == h2c.from_label_bytes(curve, b"another label", 5)
    assert L.from_label(curve, b"", 3).tobytes() == h2c.from_label_bytes(curve, b"", 3)
    assert L.from_label(curve, b"ck", 0).size == 0
    assert L.from_label(curve, b"ck", 1).tobytes() == got.tobytes()[:64]


@pytest.mark.parametrize("curve", CURVES)
def test_hash_to_curve_batch_matches_oracle(L, curve):
    rng = np.random.default_rng(curve)
    for prefix, ml, n in (("from_uniform_bytes", 32, 64), ("x", 1, 33), ("some-domain", 31, 40),
                          ("another-domain-prefix", 48, 17), ("z", 64, 9)):
        msgs = rng.integers(0, 256, size=max(n * ml, 0), dtype=np.uint8)
        got = pts_of(L.hash_to_curve_batch(curve, prefix, msgs, ml))
        want = []
        for i in range(n):
            P = h2c.hash_to_curve(curve, prefix, msgs[i * ml:(i + 1) * ml].tobytes())
            want.append(P if P is not None else (0, 0))
        assert got == want, (curve, prefix, ml)
    # special messages: all zero, all 0xff
    msgs = np.concatenate([np.zeros(32, dtype=np.uint8), np.full(32, 255, dtype=np.uint8)])
    want = [h2c.hash_to_curve(curve, "from_uniform_bytes", m.tobytes()) for m in (msgs[:32], msgs[32:])]
    assert pts_of(L.hash_to_curve_batch(curve, "from_uniform_bytes", msgs, 32)) == want


def test_hash_to_curve_argument_errors(L):
    msgs = np.zeros(64, dtype=np.uint8)
    with pytest.raises(L.LurkError) as e:
        L.hash_to_curve_batch(0, "p" * 70, msgs, 32)            # DST does not fit the single-block layout
    assert e.value.code == L._capi.ERR_ARG
    with pytest.raises(L.LurkError) as e:
        L.hash_to_curve_batch(7, "x", msgs, 32)
    assert e.value.code == L._capi.ERR_ARG
    with pytest.raises(L.LurkError) as e:
        L.from_label(0, b"ck", 4, fmt=9)
    assert e.value.code == L._capi.ERR_ARG


@pytest.mark.parametrize("curve", [0, 2])
def test_setup_key_commits_like_the_oracle(L, oracle, spec, curve):
    """CommitmentKey.setup (key generated into device memory, never on the host) + commit == oracle MSM over the oracle's key"""
    n = 700
    ck = L.CommitmentKey.setup(curve, b"ck", n)
    bases = np.frombuffer(h2c.from_label_bytes(curve, b"ck", n), dtype=np.uint8)
    sc = random_elements(spec.CURVES[curve]["scalar"], n, seed=3, shape="witness")
    assert np.array_equal(ck.commit(sc), oracle.msm(curve, bases, sc, nthreads=4))
    ck.precompute()
    assert np.array_equal(ck.commit(sc), oracle.msm(curve, bases, sc, nthreads=4))


def _on_curve_all(curve, buf):
    """vectorised on-curve test of n points with Python ints in object arrays (a few seconds for 2^17 points)"""
    p, b = h2c.base_modulus(curve), h2c.curve_b(curve)
    v = np.array(ints(buf), dtype=object)
    x, y = v[0::2], v[1::2]
    return bool(np.all((y * y - x * x * x - b) % p == 0))


@pytest.mark.parametrize("curve", CURVES)
def test_chunked_pipeline_across_chunk_boundaries(L, curve):
    """n spans three chunks of the host-XOF / GPU pipeline (2^16 points each) with a ragged tail"""
    n = (1 << 17) + 77
    got = L.from_label(curve, b"ck", n)
    stream = hashlib.shake_256(b"ck").digest(32 * n)
    pts = pts_of(got)
    for i in [0, 1, 65535, 65536, 65537, 131071, 131072, n - 1]:
        P = h2c.hash_to_curve(curve, "from_uniform_bytes", stream[32 * i:32 * i + 32])
        assert pts[i] == P, (curve, i)
    assert _on_curve_all(curve, got)
    assert len(set(pts)) == n


def test_full_size_key_bn254(L):
    """the key of the headline configuration: 2^21 BN254 G1 points (fib rc = 100), generated into device
    memory; prints the wall time (XOF on one host thread overlapped with the kernel)"""
    import torch
    curve, n = 0, 1 << 21
    buf = torch.empty(n * 64, dtype=torch.uint8, device="cuda")
    import ctypes as C
    lib = L._capi.lib()
    L._capi.check(lib.lurk_ck_generate_dev(curve, b"ck", 2, 1 << 16, C.c_void_p(buf.data_ptr()), None))   # warm-up
    torch.cuda.synchronize()
    t0 = time.time()
    L._capi.check(lib.lurk_ck_generate_dev(curve, b"ck", 2, n, C.c_void_p(buf.data_ptr()), None))
    torch.cuda.synchronize()
    dt = time.time() - t0
    print(f"\n2^21-point BN254 G1 key: {dt * 1e3:.1f} ms wall ({n / dt / 1e6:.2f} M points/s)")
    # device-side format is Montgomery: convert a sample on the host
    p = h2c.base_modulus(curve)
    Rinv = pow(1 << 256, -1, p)
    host = buf.cpu().numpy()
    stream = hashlib.shake_256(b"ck").digest(32 * n)
    rng = np.random.default_rng(1)
    for i in [0, n - 1] + [int(k) for k in rng.integers(0, n, size=14)]:
        x, y = ints(host[64 * i:64 * i + 64])
        assert (x * Rinv % p, y * Rinv % p) == h2c.hash_to_curve(curve, "from_uniform_bytes", stream[32 * i:32 * i + 32]), i
    # no duplicates anywhere (compare the x coordinates as raw bytes)
    xs = host.reshape(n, 64)[:, :32]
    assert np.unique(xs.view([("b", "V32")]).reshape(-1)).size == n
    # a commitment over the generated key behaves linearly: commit(s) + commit(t) = commit(s + t) for small s, t
    ck = L.CommitmentKey.from_device(curve, buf.data_ptr(), n)
    m = 1 << 12
    s = np.zeros(m * 32, dtype=np.uint8); s[0::32] = rng.integers(1, 100, size=m, dtype=np.uint8)
    t = np.zeros(m * 32, dtype=np.uint8); t[0::32] = rng.integers(1, 100, size=m, dtype=np.uint8)
    st = np.zeros(m * 32, dtype=np.uint8); st[0::32] = s[0::32] + t[0::32]
    both = np.concatenate([ck.commit(s), ck.commit(t)])
    assert np.array_equal(L.point_sum(curve, both), ck.commit(st))


@pytest.mark.parametrize("curve", [0, 3])
def test_key_slice_of_a_sharded_key(L, curve):
    """lurk_ck_generate_range_dev: a rank's contiguous slice [first, first + n) equals the same range of the whole key"""
    import torch
    import ctypes as C
    whole = L.from_label(curve, b"ck", 70000)
    for first, n in ((0, 10), (1, 1), (65530, 20), (4095, 130), (69990, 10)):
        buf = torch.empty(n * 64, dtype=torch.uint8, device="cuda")
        L._capi.check(L._capi.lib().lurk_ck_generate_range_dev(curve, b"ck", 2, first, n, C.c_void_p(buf.data_ptr()), None))
        L._capi.check(L._capi.lib().lurk_convert_dev(L._capi.lib() and {0: 1, 1: 0, 2: 3, 3: 2}[curve], C.c_void_p(buf.data_ptr()), 2 * n, L.FMT_CANONICAL,
                                                    C.c_void_p(buf.data_ptr()), None))
        torch.cuda.synchronize()
        assert buf.cpu().numpy().tobytes() == whole[64 * first:64 * (first + n)].tobytes(), (first, n)


def test_pasta_curves_own_vectors_on_the_gpu(L):
    """tests/golden/pasta_hash_to_curve_vectors.json (pasta_curves 0.5.0 unit tests) through lurk_hash_to_curve_batch"""
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with open(os.path.join(root, "tests", "golden", "pasta_hash_to_curve_vectors.json")) as f:
        vectors = json.load(f)["vectors"]
    for v in vectors:
        c, msg = v["curve_id"], v["message_ascii"].encode()
        p = h2c.base_modulus(c)
        x, z = int(v["x"], 16), int(v["z"], 16)
        zi = pow(z, -1, p)
        got = pts_of(L.hash_to_curve_batch(c, v["domain_prefix"], np.frombuffer(msg, dtype=np.uint8), len(msg)))[0]
        assert got[0] == x * zi * zi % p
        if v["y"]:
            assert got[1] == int(v["y"], 16) * zi * zi * zi % p
        assert got == h2c.hash_to_curve(c, v["domain_prefix"], msg)
