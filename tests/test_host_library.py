"""CPU-side checks: the C-ABI library loads and exports every symbol include/lurk_b200.h declares, its host-only
entry points agree with the oracle, the GPU limb algorithms (compiled for the host with the carry flag emulated)
agree with Python big-int arithmetic, and compute calls fail loudly without a GPU."""
import ctypes
import os
import random
import re
import subprocess

import numpy as np
import pytest

from util import ints, pack

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(L):
    header = open(os.path.join(ROOT, "include", "lurk_b200.h")).read()
    declared = set(re.findall(r"\b(lurk_[a-z0-9_]+)\s*\(", header))
    declared -= {"lurk_dag_node"}
    assert declared, "no declarations parsed"
    lib = L._capi.lib()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in L._capi.PROTOTYPES, f"{name} has no ctypes prototype"
    assert set(L._capi.PROTOTYPES) <= declared


def test_compute_fails_loudly_without_gpu(L):
    if L._capi.lib().lurk_device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(L.LurkError) as e:
        L.PoseidonCache(0).hash4([1, 2, 3, 4])
    assert e.value.code == L._capi.ERR_NOGPU
    with pytest.raises(L.LurkError):
        L.CommitmentKey(0, np.zeros(64, dtype=np.uint8))


@pytest.mark.parametrize("field", [0, 1, 2, 3])
@pytest.mark.parametrize("arity", [3, 4, 6, 8])
def test_product_poseidon_constants_match_spec(L, spec, field, arity):
    # the library generates its constants itself (host C++); the oracle generates them in Python
    c = L.HashConstants(field).constants(arity)
    P = spec.params(field, arity)
    assert (c["full_rounds"], c["partial_rounds"]) == (P["rf"], P["rp"])
    assert c["round_constants"] == P["rc"]
    assert c["mds"] == P["mds"]


def test_witness_block_sizes(L):
    ST = L.SlotType
    assert [L.compute_witness_size(s, 0) for s in (ST.Hash4, ST.Hash6, ST.Hash8, ST.Commitment, ST.BitDecomp)] == [293, 343, 396, 268, 354]
    assert [L.compute_witness_size(ST.BitDecomp, f) for f in (2, 3, 0, 1)] == [298, 301, 354, 364]   # multiframe.rs:495-498


@pytest.mark.parametrize("curve", [0, 1, 2, 3])
def test_host_point_sum_and_synthetic_bases(L, oracle, spec, curve):
    n = 300
    bases = L.synthetic_bases(curve, n, start=7)
    assert np.array_equal(bases, oracle.gen_bases(curve, n, start=7))
    C = spec.CURVES[curve]
    assert spec.on_curve(curve, (ints(bases)[0], ints(bases)[1]))
    pts = np.zeros(96 * 5, dtype=np.uint8)
    for k in range(4):                       # 4 finite points + 1 identity
        pts[96 * k:96 * k + 64] = bases[64 * k:64 * k + 64]
        pts[96 * k + 64] = 1
    assert np.array_equal(L.point_sum(curve, pts), oracle.point_sum(curve, pts))
    # P + (-P)
    pb = spec.FIELD_MODULUS[C["base"]]
    x, y = ints(bases)[0], ints(bases)[1]
    two = np.concatenate([pack([x, y, 1]), pack([x, pb - y, 1])])
    assert not L.point_sum(curve, two).any()


@pytest.fixture(scope="module", params=["emulated_gpu_limbs", "host_fast_path"])
def fieldlib(request, tmp_path_factory):
    out = str(tmp_path_factory.mktemp("fht") / f"libfht_{request.param}.so")
    flags = ["-DLURK_HOST_EMULATE_CC"] if request.param == "emulated_gpu_limbs" else []
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", *flags, "-I",
                           os.path.join(ROOT, "lurk-beta_b200", "csrc"), "-x", "c++",
                           os.path.join(ROOT, "tests", "csrc", "field_host_test.cc"), "-o", out])
    return ctypes.CDLL(out)


@pytest.mark.parametrize("field", [0, 1, 2, 3])
def test_limb_arithmetic_against_bigints(fieldlib, spec, field):
    p = spec.FIELD_MODULUS[field]
    rnd = random.Random(field)
    cases = [(0, 0), (1, 1), (p - 1, p - 1), (p - 1, 1), (1 << 253, (1 << 253) - 1)]
    cases += [(rnd.randrange(p), rnd.randrange(p)) for _ in range(400)]
    ops = {0: lambda a, b: a * b % p, 1: lambda a, b: (a + b) % p, 2: lambda a, b: (a - b) % p,
           4: lambda a, b: -a % p, 5: lambda a, b: pow(a, 5, p), 6: lambda a, b: a * a % p}
    out = ctypes.create_string_buffer(32)
    for a, b in cases:
        for op, fn in ops.items():
            assert fieldlib.fe_test_op(field, op, a.to_bytes(32, "little"), b.to_bytes(32, "little"), out) == 0
            assert int.from_bytes(out.raw, "little") == fn(a, b), (field, op)
    for a, _ in cases[:12]:
        fieldlib.fe_test_op(field, 3, a.to_bytes(32, "little"), bytes(32), out)
        assert int.from_bytes(out.raw, "little") == pow(a, p - 2, p)
    # binary-GCD inversion (the one-thread normalisation on the fold's critical chain): every case incl. 0 -> 0
    for a, b in cases:
        for x in (a, b):
            fieldlib.fe_test_op(field, 7, x.to_bytes(32, "little"), bytes(32), out)
            assert int.from_bytes(out.raw, "little") == pow(x, p - 2, p), (field, x)
    # lazy dot products (one reduction per MDS row)
    for trial in range(300):
        k = rnd.randint(1, 15)
        A = [p - 1] * k if trial < 30 else [rnd.randrange(p) for _ in range(k)]
        B = [p - 1] * k if trial < 30 else [rnd.randrange(p) for _ in range(k)]
        fieldlib.fe_test_dot(field, k, b"".join(x.to_bytes(32, "little") for x in A), b"".join(x.to_bytes(32, "little") for x in B), out)
        assert int.from_bytes(out.raw, "little") == sum(x * y for x, y in zip(A, B)) % p


@pytest.mark.parametrize("curve", [0, 1, 2, 3])
def test_host_group_law_properties(L, spec, curve):
    """host-side XYZZ arithmetic of the library (used for the N-GPU combine and the final MSM window combine) against the
    affine chord-and-tangent law of the Python spec: random multiples of G, doubling, inverse pairs, identity handling"""
    from hypothesis import given, settings, strategies as st
    C = spec.CURVES[curve]
    pb = spec.FIELD_MODULUS[C["base"]]
    G = C["gen"]
    mults = {k: spec.ec_mul(k, G, pb) for k in range(1, 40)}

    def rec(pt):
        return pack([0, 0, 0]) if pt is None else pack([pt[0], pt[1], 1])

    @settings(max_examples=60, deadline=None)
    @given(st.lists(st.integers(min_value=-39, max_value=39), min_size=0, max_size=12))
    def check(ks):
        pts, want = [], None
        for k in ks:
            if k == 0:
                pts.append(None)
                continue
            p = mults[abs(k)]
            if k < 0:
                p = (p[0], (pb - p[1]) % pb)
            pts.append(p)
            want = spec.ec_add(want, p, pb)
        buf = np.concatenate([rec(p) for p in pts]) if pts else np.zeros(0, dtype=np.uint8)
        got = L.point_sum(curve, buf)
        assert np.array_equal(got, rec(want))
    check()


def test_plain_c_client_of_the_abi(tmp_path):
    """include/lurk_b200.h is strict C99 and a gcc-built C program can drive the library (what a cgo / bindgen shim needs):
    host-only entry points work, compute entry points fail loudly without a GPU (tests/csrc/c_abi_client.c)"""
    exe = str(tmp_path / "c_abi_client")
    libdir = os.path.join(ROOT, "lurk-beta_b200")
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "csrc", "c_abi_client.c"), "-o", exe, "-L", libdir, "-llurk_b200",
                           "-Wl,-rpath," + libdir])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert "c_abi_client ok" in out.stdout


# ff::PrimeField::ROOT_OF_UNITY of the reference's field types, as published in halo2curves (bn256::Fr, GENERATOR 7;
# bn256::Fq: -1) and pasta_curves (Fq = pallas::Scalar, Fp = pallas::Base; GENERATOR 5)
PUBLISHED_ROOT_OF_UNITY = {
    0: 0x03ddb9f5166d18b798865ea93dd31f743215cf6dd39329c8d34f1ed960c37c9c,
    2: 0x2de6a9b8746d3f589e5c4dfd492ae26e9bb97ea3c106f049a70e2c1102b6d05f,
    3: 0x2bce74deac30ebda362120830561f81aea322bf2b7bb7584bdad6fabd87ea32f,
}


@pytest.mark.parametrize("field", [0, 1, 2, 3])
def test_ntt_root_of_unity_is_the_field_types_own(fieldlib, spec, field):
    """the NTT kernels (csrc/ntt.cu: omega = ROOT^(2^(s - log n))), the oracle and the reference's field types agree on the
    2^s-th root of unity: a transform computed with another primitive root is a permutation of this one"""
    p = spec.FIELD_MODULUS[field]
    out = ctypes.create_string_buffer(32)
    s = fieldlib.fe_test_root(field, out)
    root = int.from_bytes(out.raw, "little")
    assert s == spec.TWO_ADICITY[field] and root == spec.root_of_unity(field, s)
    assert pow(root, 1 << s, p) == 1 and pow(root, 1 << (s - 1), p) == p - 1
    assert root == PUBLISHED_ROOT_OF_UNITY.get(field, p - 1)


def test_plain_c_fold_driver_fails_loudly_without_gpu(tmp_path):
    """tests/csrc/fold_client.c (the fold context through plain C99): without a CUDA device the contexts cannot be created
    and the driver reports it; on the GPU box the same program folds a chain (tests/test_gpu_zz_fold_size_commit.py)"""
    exe = str(tmp_path / "fold_client")
    libdir = os.path.join(ROOT, "lurk-beta_b200")
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "csrc", "fold_client.c"),
                           "-o", exe, "-L", libdir, "-llurk_b200", "-Wl,-rpath," + libdir])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert "fold_client ok" in out.stdout


def test_n3_n4_entry_points_reject_bad_arguments_and_have_no_cpu_fallback(L):
    """argument validation of the N3 / N4 entry points happens before any device work; with valid arguments and no GPU they fail with
    LURK_ERR_NOGPU (no CPU fallback anywhere in the product)"""
    import ctypes as C
    lib, E = L._capi.lib(), L._capi
    cb = E.CHALLENGE_FN(lambda user, rnd, msg, n, out: 0)
    buf = np.zeros(4096, dtype=np.uint8)
    ptrs = (C.c_void_p * 4)(*[C.c_void_p(buf.ctypes.data)] * 4)       # never dereferenced on these paths
    nr = (C.c_int * 1)(3)
    z32 = E.np_ptr(np.zeros(32, dtype=np.uint8))
    assert lib.lurk_sumcheck_prove_dev(0, 9, ptrs, 3, z32, cb, None, None, None, None, 0, None) == E.ERR_ARG          # unknown kind
    assert lib.lurk_sumcheck_prove_dev(0, 0, ptrs, 41, z32, cb, None, None, None, None, 0, None) == E.ERR_ARG         # too many rounds
    assert lib.lurk_sumcheck_prove_dev(0, 0, ptrs, 3, z32, cb, None, None, None, None, 7, None) == E.ERR_ARG          # bad format
    assert lib.lurk_sumcheck_prove_batch_dev(0, 0, 0, ptrs, nr, z32, None, cb, None, None, None, None, 0, None) == E.ERR_ARG   # no instance
    assert lib.lurk_sumcheck_prove_batch_dev(0, 0, 61, ptrs, nr, z32, None, cb, None, None, None, None, 0, None) == E.ERR_ARG  # too many
    assert lib.lurk_eq_evals_dev(0, z32, 33, C.c_void_p(buf.ctypes.data), 0, None) == E.ERR_ARG
    assert lib.lurk_ipa_fold_scalars_dev(0, C.c_void_p(buf.ctypes.data), 6, z32, z32, 0, None) == E.ERR_ARG             # n not a power of two
    assert lib.lurk_ipa_fold_bases_dev(0, C.c_void_p(buf.ctypes.data), 1, z32, z32, 0, None) == E.ERR_ARG
    assert lib.lurk_hyperkzg_prove_dev(0, None, C.c_void_p(buf.ctypes.data), z32, 3, cb, None, None, None, None, 0, None) == E.ERR_ARG
    assert lib.lurk_ck_powers_dev(0, None, z32, 4, C.c_void_p(buf.ctypes.data), 0, None) == E.ERR_ARG
    assert lib.lurk_ck_generate_range_dev(9, b"ck", 2, 0, 4, C.c_void_p(buf.ctypes.data), None) == E.ERR_ARG            # unknown curve
    assert lib.lurk_msm_ctx_info(None, None, None) == E.ERR_ARG
    assert len(lib.lurk_last_error()) > 0
    if lib.lurk_device_count() == 0:
        assert lib.lurk_sumcheck_prove_dev(0, 0, ptrs, 3, z32, cb, None, None, None, None, 0, None) == E.ERR_NOGPU
        assert lib.lurk_eq_evals_dev(0, z32, 1, C.c_void_p(buf.ctypes.data), 0, None) == E.ERR_NOGPU
        assert lib.lurk_inner_product_dev(0, C.c_void_p(buf.ctypes.data), C.c_void_p(buf.ctypes.data), 4, z32, 0, None) == E.ERR_NOGPU
        assert lib.lurk_ck_powers_dev(0, E.np_ptr(np.zeros(64, dtype=np.uint8)), z32, 4, C.c_void_p(buf.ctypes.data), 0, None) == E.ERR_NOGPU
        assert lib.lurk_ck_generate_range_dev(0, b"ck", 2, 0, 4, C.c_void_p(buf.ctypes.data), None) == E.ERR_NOGPU
