"""N4 on the CPU: the oracle's sum-check prover against its own verifier and against independently evaluated multilinear
extensions; the host build of the product's per-index templates (sumcheck.cuh) against the oracle, round by round."""
import ctypes
import hashlib
import os
import random
import subprocess

import pytest

from oracle import sumcheck as sc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def fs_challenge(p, tag=b""):
    """a stand-in transcript: r = sha256(tag | round | message) mod p (the real Keccak256Transcript is the caller's)"""
    def f(rnd, evals):
        data = tag + bytes([rnd]) + b"".join(int(e).to_bytes(32, "little") for e in evals)
        return int.from_bytes(hashlib.sha256(data).digest() + hashlib.sha256(data + b"x").digest(), "little") % p
    return f


@pytest.mark.parametrize("field", [0, 2])
@pytest.mark.parametrize("kind", ["quad", "cubic"])
def test_oracle_prover_is_accepted_by_the_verifier(spec, field, kind):
    p = spec.FIELD_MODULUS[field]
    rnd = random.Random(field * 7 + len(kind))
    for l in (0, 1, 2, 5):
        n = 1 << l
        k = 2 if kind == "quad" else 4
        polys = [[rnd.randrange(p) for _ in range(n)] for _ in range(k)]
        comb = sc.comb_quad if kind == "quad" else sc.comb_cubic
        claim = sum(comb(*[P[i] for P in polys], p) for i in range(n)) % p
        rounds, rs, finals, last = sc.prove(polys, kind, claim, fs_challenge(p), p)
        assert len(rounds) == l
        assert sc.verify(rounds, rs, claim, 2 if kind == "quad" else 3, p) == last
        assert comb(*finals, p) == last                                      # the verifier's final check
        assert finals == [sc.mle_eval(P, rs, p) for P in polys]              # evaluations of the multilinear extensions at r
        if l:
            bad = [list(e) for e in rounds]
            bad[0][0] = (bad[0][0] + 1) % p
            assert sc.verify(bad, rs, claim, 2 if kind == "quad" else 3, p) is None


def test_eq_table_is_the_multilinear_extension_of_equality(spec):
    p = spec.FIELD_MODULUS[0]
    rnd = random.Random(3)
    tau = [rnd.randrange(p) for _ in range(4)]
    table = sc.eq_evals(tau, p)
    assert sum(table) % p == 1
    for i, v in enumerate(table):
        want = 1
        for j, t in enumerate(tau):
            want = want * (t if (i >> (3 - j)) & 1 else 1 - t) % p
        assert v == want
    assert sc.eq_evals([1, 0, 1, 1], p)[0b1011] == 1


@pytest.fixture(scope="module", params=["emulated_gpu_limbs", "host_fast_path"])
def sclib(request, tmp_path_factory):
    out = str(tmp_path_factory.mktemp("sc") / f"libsc_{request.param}.so")
    flags = ["-DLURK_HOST_EMULATE_CC"] if request.param == "emulated_gpu_limbs" else []
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", *flags, "-I",
                           os.path.join(ROOT, "lurk-beta_b200", "csrc"), "-x", "c++",
                           os.path.join(ROOT, "tests", "csrc", "sumcheck_host_test.cc"), "-o", out])
    return ctypes.CDLL(out)


def pack(vals):
    return b"".join(int(v).to_bytes(32, "little") for v in vals)


def unpack(buf, n):
    return [int.from_bytes(buf[32 * i:32 * i + 32], "little") for i in range(n)]


@pytest.mark.parametrize("field", [0, 1, 2, 3])
@pytest.mark.parametrize("kind", ["quad", "cubic"])
def test_product_round_arithmetic_against_oracle(sclib, spec, field, kind):
    """every round of a 2^5 sum-check: the product's bind + evaluation templates give the oracle's round polynomial"""
    p = spec.FIELD_MODULUS[field]
    rnd = random.Random(11 + field)
    k, e = (2, 2) if kind == "quad" else (4, 3)
    n = 32
    polys = [[rnd.randrange(p) for _ in range(n)] for _ in range(k)]
    polys[0][3] = 0
    polys[1][5] = p - 1
    comb = sc.comb_quad if kind == "quad" else sc.comb_cubic
    claim = sum(comb(*[P[i] for P in polys], p) for i in range(n)) % p
    rounds, rs, finals, last = sc.prove(polys, kind, claim, fs_challenge(p, b"t"), p)
    cur = [list(P) for P in polys]
    length = n
    for j in range(5):
        buf = ctypes.create_string_buffer(pack([v for P in cur for v in P + [0] * (length - len(P))]), k * length * 32)
        ev = ctypes.create_string_buffer(32 * e)
        r_prev = pack([rs[j - 1]]) if j else bytes(32)
        assert sclib.sc_test_round(field, 0 if kind == "quad" else 1, buf, length, 1 if j else 0, r_prev, ev) == 0
        got = unpack(ev.raw, e)
        assert got[0] == rounds[j][0] and got[1:] == rounds[j][2:], (field, kind, j)
        if j:
            length //= 2
            allv = unpack(buf.raw, k * length * 2)
            cur = [allv[2 * length * i:2 * length * i + length] for i in range(k)]
        # claim bookkeeping (host side of the library): interpolation at the challenge
        out = ctypes.create_string_buffer(32)
        assert sclib.sc_test_interpolate(field, pack(rounds[j]), e + 1, pack([rs[j]]), out) == 0
        assert unpack(out.raw, 1)[0] == sc.uni_eval_from_evals(rounds[j], rs[j], p)
    assert [sc.bind_top(P, rs[4], p)[0] for P in cur] == finals


@pytest.mark.parametrize("curve", [0, 1, 2, 3])
def test_product_ipa_fold_arithmetic_against_oracle(sclib, spec, curve):
    C = spec.CURVES[curve]
    pb, q = spec.FIELD_MODULUS[C["base"]], spec.FIELD_MODULUS[C["scalar"]]
    rnd = random.Random(curve)
    out = ctypes.create_string_buffer(64)
    for _ in range(6):
        lo, hi, x, y = (rnd.randrange(q) for _ in range(4))
        assert sclib.sc_test_fold_scalar(C["scalar"], pack([lo]), pack([hi]), pack([x]), pack([y]), out) == 0
        assert unpack(out.raw, 1)[0] == (lo * x + hi * y) % q
    G = C["gen"]
    P, Q = spec.ec_mul(rnd.randrange(q), G, pb), spec.ec_mul(rnd.randrange(q), G, pb)
    cases = [(rnd.randrange(q), rnd.randrange(q)), (1, 1), (0, 5), (7, 0), (q - 1, 1), (rnd.randrange(1 << 128), rnd.randrange(q))]
    for x, y in cases:
        assert sclib.sc_test_fold_point(curve, pack(P), pack(Q), pack([x]), pack([y]), out) == 0
        want = sc.ipa_fold_bases(curve, [P, Q], x, y)[0]
        assert tuple(unpack(out.raw, 2)) == (want if want is not None else (0, 0)), (curve, x, y)
    # P = Q and P = -Q exercise the doubling / cancellation branches of the interleaved ladder
    for Q2, (x, y) in ((P, (3, 4)), ((P[0], pb - P[1]), (5, 5)), ((P[0], pb - P[1]), (6, 2))):
        assert sclib.sc_test_fold_point(curve, pack(P), pack(Q2), pack([x]), pack([y]), out) == 0
        want = sc.ipa_fold_bases(curve, [P, Q2], x, y)[0]
        assert tuple(unpack(out.raw, 2)) == (want if want is not None else (0, 0))


@pytest.mark.parametrize("kind", ["quad", "cubic"])
def test_oracle_batched_prover_final_check(spec, kind):
    """batched sum-check over instances of different sizes: the verifier's final check sum_i coeff_i comb(final_i) eq-free form:
    the last claim equals the combination of the instances' final evaluations, and every instance's final evaluations are its
    multilinear extensions at the LAST nr_i challenges"""
    p = spec.FIELD_MODULUS[0]
    rnd = random.Random(5)
    k = 2 if kind == "quad" else 4
    sizes = [5, 3, 0, 5, 1]
    insts = [[[rnd.randrange(p) for _ in range(1 << l)] for _ in range(k)] for l in sizes]
    comb = sc.comb_quad if kind == "quad" else sc.comb_cubic
    claims = [sum(comb(*[P[i] for P in polys], p) for i in range(len(polys[0]))) % p for polys in insts]
    coeffs = [rnd.randrange(p) for _ in sizes]
    rounds, rs, finals, last = sc.prove_batch(insts, kind, claims, coeffs, fs_challenge(p, b"b"), p)
    assert len(rounds) == 5
    e0 = sum(c * (1 << (5 - l)) * cl for c, l, cl in zip(coeffs, sizes, claims)) % p
    assert sc.verify(rounds, rs, e0, 2 if kind == "quad" else 3, p) == last
    assert last == sum(c * comb(*f, p) for c, f in zip(coeffs, finals)) % p
    for polys, l, f in zip(insts, sizes, finals):
        assert f == [sc.mle_eval(P, rs[5 - l:], p) for P in polys]
    # one instance with coefficient 1 is the plain prover
    single = sc.prove(insts[0], kind, claims[0], fs_challenge(p, b"b"), p)
    again = sc.prove_batch([insts[0]], kind, [claims[0]], [1], fs_challenge(p, b"b"), p)
    assert (single[0], single[1], single[2]) == (again[0], again[1], again[2][0])


@pytest.mark.parametrize("kind", ["quad", "cubic"])
def test_c_port_of_the_sumcheck_prover_equals_python(oracle, spec, kind):
    """oracle.c: oracle_sc_* (the CPU baseline of the N4 sum-check rows) against oracle/sumcheck.py, multi-threaded"""
    import numpy as np
    from util import ints, random_elements
    p = spec.FIELD_MODULUS[2]
    k = 2 if kind == "quad" else 4
    for l in (0, 1, 7):
        bufs = [random_elements(2, 1 << l, seed=5 * l + i) for i in range(k)]
        polys = [ints(b) for b in bufs]
        comb = sc.comb_quad if kind == "quad" else sc.comb_cubic
        claim = sum(comb(*[P[i] for P in polys], p) for i in range(1 << l)) % p
        want = sc.prove(polys, kind, claim, fs_challenge(p, b"c"), p)
        got = oracle.sumcheck_prove(2, kind, bufs, l, claim, fs_challenge(p, b"c"), nthreads=3)
        assert (got[0], got[1], got[2]) == (want[0], want[1], want[2])
