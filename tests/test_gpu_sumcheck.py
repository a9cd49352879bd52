"""GPU parity for N4 (the data-parallel loops of `compress`): sum-check prover rounds, eq table, inner product and
the inner-product argument's folding rounds through the C ABI against oracle/sumcheck.py -- bit-exact round messages for the same
challenges at sizes the oracle finishes in seconds, and at 2^21 (the outer sum-check of a fib rc = 100 step circuit) through
size-independent properties: the oracle VERIFIER accepts the GPU prover's transcript and the final evaluations equal the
multilinear extensions evaluated independently (eq table + inner product on untouched copies).
The challenge function is a stand-in (sha256); Arecibo's Keccak256Transcript is the caller's side of the callback."""
import hashlib
import time

import numpy as np
import pytest

from oracle import sumcheck as sc
from util import ints, pack, random_elements

pytestmark = pytest.mark.gpu


def fs_challenge(p, tag=b""):
    def f(rnd, evals):
        data = tag + bytes([rnd]) + b"".join(int(e).to_bytes(32, "little") for e in evals)
        return int.from_bytes(hashlib.sha256(data).digest() + hashlib.sha256(data + b"x").digest(), "little") % p
    return f


def msg_challenge(p, tag=b""):
    """the same function on the raw callback message (canonical 32-byte elements)"""
    inner = fs_challenge(p, tag)

    def f(rnd, msg):
        return inner(rnd, [int.from_bytes(msg[i:i + 32], "little") for i in range(0, len(msg), 32)])
    return f


def to_device(L, field, canon_buf):
    """canonical host elements -> Montgomery device tensor"""
    import torch
    import ctypes as C
    t = torch.from_numpy(np.ascontiguousarray(canon_buf, dtype=np.uint8)).cuda()
    L._capi.check(L._capi.lib().lurk_convert_dev(field, C.c_void_p(t.data_ptr()), t.numel() // 32, L.FMT_MONTGOMERY, C.c_void_p(t.data_ptr()), None))
    return t


def from_device(L, field, t):
    import ctypes as C
    c = t.clone()
    L._capi.check(L._capi.lib().lurk_convert_dev(field, C.c_void_p(c.data_ptr()), c.numel() // 32, L.FMT_CANONICAL, C.c_void_p(c.data_ptr()), None))
    return ints(c.cpu().numpy())


@pytest.mark.parametrize("field", [0, 1, 2, 3])
def test_eq_table_and_inner_product(L, spec, field):
    import torch
    p = spec.FIELD_MODULUS[field]
    rng = np.random.default_rng(field)
    for l in (0, 1, 2, 3, 4, 5, 9, 12):
        tau = ints(random_elements(field, max(l, 1), seed=l))[:l]
        out = torch.empty((1 << l) * 32, dtype=torch.uint8, device="cuda")
        L.spartan.eq_evals(field, tau, out.data_ptr())
        assert from_device(L, field, out) == sc.eq_evals(tau, p), (field, l)
        L.spartan.eq_evals(field, tau, out.data_ptr(), out_fmt=L.FMT_CANONICAL)
        torch.cuda.synchronize()
        assert ints(out.cpu().numpy()) == sc.eq_evals(tau, p)
    for n in (1, 2, 31, 1000, (1 << 16) + 3):
        a, b = random_elements(field, n, seed=n), random_elements(field, n, seed=n + 1)
        da, db = to_device(L, field, a), to_device(L, field, b)
        assert L.spartan.inner_product(field, da.data_ptr(), db.data_ptr(), n) == sc.inner_product(ints(a), ints(b), p), (field, n)
    assert L.spartan.inner_product(field, 0, 0, 0) == 0


@pytest.mark.parametrize("field", [0, 2])
@pytest.mark.parametrize("kind", ["quad", "cubic"])
@pytest.mark.parametrize("l", [0, 1, 2, 3, 7, 11])
def test_sumcheck_rounds_match_oracle(L, spec, field, kind, l):
    p = spec.FIELD_MODULUS[field]
    n, k = 1 << l, (2 if kind == "quad" else 4)
    bufs = [random_elements(field, n, seed=17 * l + i, shape="witness" if i == 3 else "uniform") for i in range(k)]
    polys = [ints(b) for b in bufs]
    comb = sc.comb_quad if kind == "quad" else sc.comb_cubic
    claim = sum(comb(*[P[i] for P in polys], p) for i in range(n)) % p
    want_rounds, want_rs, want_fin, last = sc.prove(polys, kind, claim, fs_challenge(p, b"g"), p)
    dev = [to_device(L, field, b) for b in bufs]
    rounds, rs, fin = L.spartan.sumcheck_prove(field, L.spartan.QUAD if kind == "quad" else L.spartan.CUBIC, [d.data_ptr() for d in dev], l, claim,
                                               msg_challenge(p, b"g"))
    assert rounds == want_rounds and rs == want_rs and fin == want_fin
    assert sc.verify(rounds, rs, claim, 2 if kind == "quad" else 3, p) == last == comb(*fin, p)
    # the polynomials were bound in place: slot 0 holds the final evaluation
    assert [from_device(L, field, d[:32])[0] for d in dev] == want_fin


def test_sumcheck_other_fields_and_errors(L, spec):
    for field in (1, 3):
        p = spec.FIELD_MODULUS[field]
        bufs = [random_elements(field, 64, seed=5 + i) for i in range(2)]
        polys = [ints(b) for b in bufs]
        claim = sc.inner_product(polys[0], polys[1], p)
        want = sc.prove(polys, "quad", claim, fs_challenge(p), p)
        dev = [to_device(L, field, b) for b in bufs]
        got = L.spartan.sumcheck_prove(field, L.spartan.QUAD, [d.data_ptr() for d in dev], 6, claim, msg_challenge(p))
        assert (got[0], got[1], got[2]) == (want[0], want[1], want[2])
    dev = [to_device(L, 0, random_elements(0, 8, seed=i)) for i in range(2)]
    with pytest.raises(L.LurkError) as e:      # a challenge >= p is refused (LURK_ERR_RANGE), nothing is silently reduced
        L.spartan.sumcheck_prove(0, L.spartan.QUAD, [d.data_ptr() for d in dev], 3, 0, lambda r, m: (1 << 256) - 1)
    assert e.value.code == L._capi.ERR_RANGE
    with pytest.raises(ZeroDivisionError):     # an exception inside the callback aborts the proof and is re-raised
        L.spartan.sumcheck_prove(0, L.spartan.QUAD, [d.data_ptr() for d in dev], 3, 0, lambda r, m: 1 // 0)
    with pytest.raises(L.LurkError) as e:
        L.spartan.sumcheck_prove(0, 7, [d.data_ptr() for d in dev], 3, 0, lambda r, m: 1)
    assert e.value.code == L._capi.ERR_ARG


def test_outer_sumcheck_full_size(L, spec):
    """2^21 rows (fib rc = 100: 1 114 100 constraints padded): claim = sum_x eq(tau, x) (Az(x) Bz(x) - (u Cz + E)(x)) with a residual
    on a few rows, proven on the GPU; the oracle verifier accepts and the four final evaluations are the multilinear extensions
    at the challenge point, recomputed from untouched copies with the eq-table and inner-product kernels"""
    import torch
    field, l = 0, 21
    p = spec.FIELD_MODULUS[field]
    n = 1 << l
    rng = np.random.default_rng(2)
    tau = ints(random_elements(field, l, seed=99))
    A = torch.empty(n * 32, dtype=torch.uint8, device="cuda")
    L.spartan.eq_evals(field, tau, A.data_ptr())
    Bh, Ch = random_elements(field, n, seed=1), random_elements(field, n, seed=2, shape="witness")
    Bi, Ci = ints(Bh), ints(Ch)
    Di = [b * c % p for b, c in zip(Bi, Ci)]
    touched = [int(x) for x in rng.integers(0, n, size=5)]
    for t in touched:
        Di[t] = (Di[t] + 1 + t) % p                      # residual -(1 + t) on row t
    eq_t = from_device(L, field, A)
    claim = sum(-(1 + t) * eq_t[t] for t in set(touched)) % p
    B, Cc, D = to_device(L, field, Bh), to_device(L, field, Ch), to_device(L, field, pack(Di))
    keep = [x.clone() for x in (A, B, Cc, D)]
    torch.cuda.synchronize()
    t0 = time.time()
    rounds, rs, fin = L.spartan.sumcheck_prove(field, L.spartan.CUBIC, [x.data_ptr() for x in (A, B, Cc, D)], l, claim, msg_challenge(p, b"outer"))
    dt = time.time() - t0
    print(f"\nouter sum-check, 2^21 rows x 4 polynomials, 21 rounds (Python callback transcript): {dt * 1e3:.1f} ms")
    last = sc.verify(rounds, rs, claim, 3, p)
    assert last is not None and last == sc.comb_cubic(*fin, p)
    eq_r = torch.empty(n * 32, dtype=torch.uint8, device="cuda")
    L.spartan.eq_evals(field, rs, eq_r.data_ptr())
    assert [L.spartan.inner_product(field, k.data_ptr(), eq_r.data_ptr(), n) for k in keep] == fin
    # eq(tau, .) evaluated at r has the closed form prod_j (tau_j r_j + (1 - tau_j)(1 - r_j))
    want = 1
    for t, r in zip(tau, rs):
        want = want * (t * r + (1 - t) * (1 - r)) % p
    assert fin[0] == want


@pytest.mark.parametrize("curve", [0, 2])
def test_ipa_folds_match_oracle(L, oracle, spec, curve):
    import torch
    C = spec.CURVES[curve]
    pb, q = spec.FIELD_MODULUS[C["base"]], spec.FIELD_MODULUS[C["scalar"]]
    n = 16
    a = random_elements(C["scalar"], n, seed=1)
    x, y = ints(random_elements(C["scalar"], 2, seed=2))
    da = to_device(L, C["scalar"], a)
    L.spartan.ipa_fold_scalars(C["scalar"], da.data_ptr(), n, x, y)
    assert from_device(L, C["scalar"], da)[:n // 2] == sc.ipa_fold_scalars(ints(a), x, y, q)
    bases = oracle.gen_bases(curve, n)
    bases[64 * 3:64 * 4] = bases[64 * 11:64 * 12]           # G[3] = G[3 + n/2]: the P = Q branch
    G = [(v[0], v[1]) for v in zip(ints(bases)[0::2], ints(bases)[1::2])]
    dG = to_device(L, C["base"], bases)
    L.spartan.ipa_fold_bases(curve, dG.data_ptr(), n, x, y)
    got = from_device(L, C["base"], dG)
    want = sc.ipa_fold_bases(curve, G, x, y)
    assert list(zip(got[0:n:2], got[1:n:2])) == [w if w is not None else (0, 0) for w in want]


@pytest.mark.parametrize("curve,log_n", [(0, 4), (2, 4), (1, 7)])
def test_ipa_prove_rounds_and_verifier_relation(L, oracle, spec, curve, log_n):
    """every round message equals the oracle's recomputation (MSM by the C oracle, folds in Python), and the verifier's relation
    holds: commit(a'; G') + a' b' ck_c = P + sum_i (r_i^2 L_i + r_i^-2 R_i) with P = commit(a; G) + <a, b> ck_c"""
    Cv = spec.CURVES[curve]
    pb, q = spec.FIELD_MODULUS[Cv["base"]], spec.FIELD_MODULUS[Cv["scalar"]]
    n = 1 << log_n
    a_h, b_h = random_elements(Cv["scalar"], n, seed=3), random_elements(Cv["scalar"], n, seed=4, shape="witness")
    bases = oracle.gen_bases(curve, n + 1, start=5)
    Gs = list(zip(ints(bases)[0::2], ints(bases)[1::2]))
    G, gc = Gs[:n], Gs[n]
    a, b = ints(a_h), ints(b_h)
    add = lambda P, Q: spec.ec_add(P, Q, pb)
    mul = lambda k, P: spec.ec_mul(k % q, P, pb)
    P0 = add(spec.msm_naive(curve, G, a), mul(sc.inner_product(a, b, q), gc))
    da, db = to_device(L, Cv["scalar"], a_h), to_device(L, Cv["scalar"], b_h)
    ck = L.CommitmentKey(curve, bases[:64 * n])
    if log_n == 7:
        ck.precompute()                      # the fixed-base table of the key serves every round

    def chal(rnd, msg):
        return 1 + int.from_bytes(hashlib.sha256(bytes([rnd]) + msg).digest()[:16], "little")      # 128-bit, non-zero

    Ls, Rs, a_fin, b_fin = L.spartan.ipa_prove(curve, ck, gc, da.data_ptr(), db.data_ptr(), log_n, chal)
    # oracle recomputation, round by round
    acc = P0
    for rnd in range(log_n):
        h = len(a) // 2
        cl, cr = sc.inner_product(a[:h], b[h:], q), sc.inner_product(a[h:], b[:h], q)
        Lw = add(spec.msm_naive(curve, G[h:], a[:h]), mul(cl, gc))
        Rw = add(spec.msm_naive(curve, G[:h], a[h:]), mul(cr, gc))
        assert Ls[rnd] == Lw and Rs[rnd] == Rw, rnd
        enc = lambda P: (pack([P[0], P[1], 1]) if P is not None else np.zeros(96, dtype=np.uint8)).tobytes()
        r = chal(rnd, enc(Lw) + enc(Rw))
        ri = pow(r, -1, q)
        acc = add(acc, add(mul(r * r, Lw), mul(ri * ri, Rw)))
        a = sc.ipa_fold_scalars(a, r, ri, q)
        b = sc.ipa_fold_scalars(b, ri, r, q)
        G = sc.ipa_fold_bases(curve, G, ri, r)
    assert (a_fin, b_fin) == (a[0], b[0])
    # the key is not consumed: committing under it still gives the oracle's commitment
    assert np.array_equal(ck.commit(a_h), oracle.msm(curve, bases[:64 * n], a_h, nthreads=4))
    assert add(mul(a_fin, G[0]), mul(a_fin * b_fin, gc)) == acc


@pytest.mark.parametrize("kind", ["quad", "cubic"])
def test_batched_sumcheck_matches_oracle(L, spec, kind):
    """SuperNova's batched sum-check: instances of 2^9, 2^4, 2^0, 2^9, 2^1 elements with random coefficients"""
    field = 0
    p = spec.FIELD_MODULUS[field]
    k = 2 if kind == "quad" else 4
    sizes = [9, 4, 0, 9, 1]
    bufs = [[random_elements(field, 1 << l, seed=100 * i + j) for j in range(k)] for i, l in enumerate(sizes)]
    insts = [[ints(b) for b in polys] for polys in bufs]
    comb = sc.comb_quad if kind == "quad" else sc.comb_cubic
    claims = [sum(comb(*[P[i] for P in polys], p) for i in range(len(polys[0]))) % p for polys in insts]
    coeffs = ints(random_elements(field, len(sizes), seed=7))
    want = sc.prove_batch(insts, kind, claims, coeffs, fs_challenge(p, b"B"), p)
    dev = [[to_device(L, field, b) for b in polys] for polys in bufs]
    got = L.spartan.sumcheck_prove_batch(field, L.spartan.QUAD if kind == "quad" else L.spartan.CUBIC,
                                         [([d.data_ptr() for d in polys], l) for polys, l in zip(dev, sizes)], claims, coeffs, msg_challenge(p, b"B"))
    assert got[0] == want[0] and got[1] == want[1] and got[2] == want[2]
    e0 = sum(c * (1 << (9 - l)) * cl for c, l, cl in zip(coeffs, sizes, claims)) % p
    assert sc.verify(got[0], got[1], e0, 2 if kind == "quad" else 3, p) == sum(c * comb(*f, p) for c, f in zip(coeffs, got[2])) % p


def test_provers_from_concurrent_host_threads(L, spec):
    """the reference calls its prover from rayon workers: four host threads run sum-checks at once (per-thread reduction scratch,
    callbacks re-entering Python) and every transcript equals the oracle's"""
    import threading
    field = 0
    p = spec.FIELD_MODULUS[field]
    l = 9
    results, errors = {}, []

    def work(t):
        try:
            bufs = [random_elements(field, 1 << l, seed=1000 * t + i) for i in range(4)]
            polys = [ints(b) for b in bufs]
            claim = sum(sc.comb_cubic(*[P[i] for P in polys], p) for i in range(1 << l)) % p
            dev = [to_device(L, field, b) for b in bufs]
            for _ in range(3):
                work_copy = [d.clone() for d in dev]
                got = L.spartan.sumcheck_prove(field, L.spartan.CUBIC, [d.data_ptr() for d in work_copy], l, claim, msg_challenge(p, bytes([t])))
            results[t] = (got, sc.prove(polys, "cubic", claim, fs_challenge(p, bytes([t])), p))
        except Exception as e:       # pragma: no cover
            errors.append(e)

    threads = [threading.Thread(target=work, args=(t,)) for t in range(4)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors
    for t in range(4):
        got, want = results[t]
        assert (got[0], got[1], got[2]) == (want[0], want[1], want[2]), t
