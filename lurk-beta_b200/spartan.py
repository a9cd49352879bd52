"""Host-side mirror of the Arecibo interfaces behind `compress` (reference src/proof/nova.rs:341-356 -> CompressedSNARK::prove
-> spartan::snark::RelaxedR1CSSNARK::prove): SumcheckProof::prove_quad / prove_cubic_with_additive_term, EqPolynomial::evals,
MultilinearPolynomial::evaluate and provider::ipa_pc::InnerProductArgument::prove, all on device-resident vectors through the C ABI
(include/lurk_b200.h, N4).  The transcript is the caller's: `challenge(round, message_bytes) -> int`."""
import ctypes as C

import numpy as np

from . import _capi

QUAD, CUBIC = _capi.SUMCHECK_QUAD, _capi.SUMCHECK_CUBIC


def _callback(challenge, errors):
    def cb(user, rnd, msg, msg_len, out):
        try:
            r = int(challenge(rnd, bytes(msg[i] for i in range(msg_len))))
            for i, byte in enumerate(r.to_bytes(32, "little")):
                out[i] = byte
            return 0
        except Exception as e:          # never unwind through the C frames
            errors.append(e)
            return 1
    return _capi.CHALLENGE_FN(cb)


def _ints(buf):
    return [int.from_bytes(buf[i:i + 32].tobytes(), "little") for i in range(0, buf.size, 32)]


def _fe(x):
    return np.frombuffer(int(x).to_bytes(32, "little"), dtype=np.uint8).copy()


def sumcheck_prove(field_id, kind, poly_ptrs, num_rounds, claim, challenge, stream=0):
    """poly_ptrs: device pointers of the 2 (QUAD) or 4 (CUBIC) polynomials, 2^num_rounds Montgomery elements each (consumed).
    claim: int.  challenge(round, message) -> int (canonical).  Returns (round_evals [[int]], challenges [int], final_evals [int])."""
    k, deg1 = (2, 3) if kind == QUAD else (4, 4)
    ptrs = (C.c_void_p * k)(*[C.c_void_p(p) for p in poly_ptrs])
    rounds = np.zeros(max(1, num_rounds) * deg1 * 32, dtype=np.uint8)
    chal = np.zeros(max(1, num_rounds) * 32, dtype=np.uint8)
    fin = np.zeros(k * 32, dtype=np.uint8)
    errors = []
    cb = _callback(challenge, errors)
    rc = _capi.lib().lurk_sumcheck_prove_dev(field_id, kind, ptrs, num_rounds, _capi.np_ptr(_fe(claim)), cb, None, _capi.np_ptr(rounds),
                                             _capi.np_ptr(chal), _capi.np_ptr(fin), _capi.FMT_CANONICAL, C.c_void_p(stream))
    if errors:
        raise errors[0]
    _capi.check(rc)
    ev = _ints(rounds)
    return [ev[i * deg1:(i + 1) * deg1] for i in range(num_rounds)], _ints(chal)[:num_rounds], _ints(fin)


def sumcheck_prove_batch(field_id, kind, instances, claims, coeffs, challenge, stream=0):
    """SumcheckProof::prove_*_batch.  instances: [(poly_ptrs, num_rounds)], claims / coeffs: ints.
    Returns (round_evals, challenges, final_evals per instance)."""
    k, deg1 = (2, 3) if kind == QUAD else (4, 4)
    n = len(instances)
    flat = [p for ptrs, _ in instances for p in ptrs]
    ptrs = (C.c_void_p * (n * k))(*[C.c_void_p(p) for p in flat])
    nr = (C.c_int * n)(*[r for _, r in instances])
    mx = max(r for _, r in instances)
    rounds = np.zeros(max(1, mx) * deg1 * 32, dtype=np.uint8)
    chal = np.zeros(max(1, mx) * 32, dtype=np.uint8)
    fin = np.zeros(n * k * 32, dtype=np.uint8)
    cl = np.concatenate([_fe(c) for c in claims])
    co = np.concatenate([_fe(c) for c in coeffs])
    errors = []
    cb = _callback(challenge, errors)
    rc = _capi.lib().lurk_sumcheck_prove_batch_dev(field_id, kind, n, ptrs, nr, _capi.np_ptr(cl), _capi.np_ptr(co), cb, None, _capi.np_ptr(rounds),
                                                   _capi.np_ptr(chal), _capi.np_ptr(fin), _capi.FMT_CANONICAL, C.c_void_p(stream))
    if errors:
        raise errors[0]
    _capi.check(rc)
    ev, fi = _ints(rounds), _ints(fin)
    return [ev[i * deg1:(i + 1) * deg1] for i in range(mx)], _ints(chal)[:mx], [fi[i * k:(i + 1) * k] for i in range(n)]


def eq_evals(field_id, tau, d_out_ptr, out_fmt=_capi.FMT_MONTGOMERY, stream=0):
    """EqPolynomial::new(tau).evals() into device memory (2^len(tau) elements in out_fmt); tau: ints"""
    R = 1 << 256
    if out_fmt == _capi.FMT_MONTGOMERY:
        p = int.from_bytes(field_modulus(field_id), "little")
        tau = [t * R % p for t in tau]
    buf = np.frombuffer(b"".join(int(t).to_bytes(32, "little") for t in tau) or bytes(32), dtype=np.uint8).copy()
    _capi.check(_capi.lib().lurk_eq_evals_dev(field_id, _capi.np_ptr(buf), len(tau), C.c_void_p(d_out_ptr), out_fmt, C.c_void_p(stream)))


def field_modulus(field_id):
    out = np.zeros(32, dtype=np.uint8)
    _capi.check(_capi.lib().lurk_field_modulus(field_id, _capi.np_ptr(out)))
    return out.tobytes()


def inner_product(field_id, d_a_ptr, d_b_ptr, n, stream=0):
    out = np.zeros(32, dtype=np.uint8)
    _capi.check(_capi.lib().lurk_inner_product_dev(field_id, C.c_void_p(d_a_ptr), C.c_void_p(d_b_ptr), n, _capi.np_ptr(out),
                                                   _capi.FMT_CANONICAL, C.c_void_p(stream)))
    return int.from_bytes(out.tobytes(), "little")


def ipa_fold_scalars(field_id, d_a_ptr, n, x, y, stream=0):
    _capi.check(_capi.lib().lurk_ipa_fold_scalars_dev(field_id, C.c_void_p(d_a_ptr), n, _capi.np_ptr(_fe(x)), _capi.np_ptr(_fe(y)),
                                                      _capi.FMT_CANONICAL, C.c_void_p(stream)))


def ipa_fold_bases(curve_id, d_bases_ptr, n, x, y, stream=0):
    _capi.check(_capi.lib().lurk_ipa_fold_bases_dev(curve_id, C.c_void_p(d_bases_ptr), n, _capi.np_ptr(_fe(x)), _capi.np_ptr(_fe(y)),
                                                    _capi.FMT_CANONICAL, C.c_void_p(stream)))


def ipa_prove(curve_id, ck, ck_c, d_a_ptr, d_b_ptr, log_n, challenge, stream=0):
    """InnerProductArgument::prove's rounds under the key of CommitmentKey `ck` (not consumed).  ck_c: (x, y) canonical ints.
    Returns (L points, R points, a_final, b_final); points are (x, y) tuples or None for the identity."""
    gc = np.concatenate([_fe(ck_c[0]), _fe(ck_c[1])])
    Ls = np.zeros(max(1, log_n) * 96, dtype=np.uint8)
    Rs = np.zeros(max(1, log_n) * 96, dtype=np.uint8)
    af, bf = np.zeros(32, dtype=np.uint8), np.zeros(32, dtype=np.uint8)
    errors = []
    cb = _callback(challenge, errors)
    rc = _capi.lib().lurk_ipa_prove_dev(curve_id, ck._ctx, _capi.np_ptr(gc), C.c_void_p(d_a_ptr), C.c_void_p(d_b_ptr), log_n, cb,
                                        None, _capi.np_ptr(Ls), _capi.np_ptr(Rs), _capi.np_ptr(af), _capi.np_ptr(bf), _capi.FMT_CANONICAL,
                                        C.c_void_p(stream))
    if errors:
        raise errors[0]
    _capi.check(rc)

    def pts(buf):
        out = []
        for i in range(log_n):
            b = buf[96 * i:96 * i + 96].tobytes()
            z = int.from_bytes(b[64:], "little")
            out.append((int.from_bytes(b[:32], "little"), int.from_bytes(b[32:64], "little")) if z else None)
        return out
    return pts(Ls), pts(Rs), int.from_bytes(af.tobytes(), "little"), int.from_bytes(bf.tobytes(), "little")


def hyperkzg_prove(curve_id, ck, d_poly_ptr, point, challenge, stream=0):
    """provider::hyperkzg::EvaluationEngine::prove.  ck: a CommitmentKey on the KZG key; point: ints.  challenge(round, message) -> int
    with round 0 = commitments, 1 = evaluations, 2 = witness commitments.  Returns (com points, v [3][l] ints, w points)."""
    l = len(point)
    pt = np.frombuffer(b"".join(int(x).to_bytes(32, "little") for x in point), dtype=np.uint8).copy()
    com = np.zeros(max(1, l - 1) * 96, dtype=np.uint8)
    w = np.zeros(3 * 96, dtype=np.uint8)
    v = np.zeros(3 * l * 32, dtype=np.uint8)
    errors = []
    cb = _callback(challenge, errors)
    rc = _capi.lib().lurk_hyperkzg_prove_dev(curve_id, ck._ctx, C.c_void_p(d_poly_ptr), _capi.np_ptr(pt), l, cb, None, _capi.np_ptr(com),
                                             _capi.np_ptr(w), _capi.np_ptr(v), _capi.FMT_CANONICAL, C.c_void_p(stream))
    if errors:
        raise errors[0]
    _capi.check(rc)

    def pts(buf, k):
        out = []
        for i in range(k):
            b = buf[96 * i:96 * i + 96].tobytes()
            out.append((int.from_bytes(b[:32], "little"), int.from_bytes(b[32:64], "little")) if int.from_bytes(b[64:], "little") else None)
        return out
    vi = _ints(v)
    return pts(com, l - 1), [vi[t * l:(t + 1) * l] for t in range(3)], pts(w, 3)
