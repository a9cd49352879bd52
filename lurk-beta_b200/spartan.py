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



# ------------------------------------------------------------------------------------------------ RelaxedR1CSSNARK::prove, the GPU half
class DeviceCSR:
    """a CSR matrix resident on the GPU (row_ptr u64, col u32, val Montgomery field elements)"""

    def __init__(self, field_id, rows, row_ptr, col, val_canonical):
        import torch
        self.rows = rows
        self.rp = torch.from_numpy(np.ascontiguousarray(row_ptr, dtype=np.uint64).view(np.int64)).cuda()
        self.col = torch.from_numpy(np.ascontiguousarray(col, dtype=np.uint32).view(np.int32)).cuda()
        self.val = torch.from_numpy(np.ascontiguousarray(val_canonical, dtype=np.uint8).reshape(-1)).cuda()
        if self.val.numel():
            _capi.check(_capi.lib().lurk_convert_dev(field_id, C.c_void_p(self.val.data_ptr()), self.val.numel() // 32, _capi.FMT_MONTGOMERY,
                                                     C.c_void_p(self.val.data_ptr()), None))

    def mv(self, field_id, d_z_ptr, d_y_ptr):
        _capi.check(_capi.lib().lurk_spmv_csr_dev(field_id, C.c_void_p(self.rp.data_ptr()), C.c_void_p(self.col.data_ptr()), C.c_void_p(self.val.data_ptr()),
                                                  self.rows, C.c_void_p(d_z_ptr), C.c_void_p(d_y_ptr), None))


def padded_and_transposed(mats, n_w, num_vars, rows_pad):
    """host-side set-up (once per circuit shape): columns re-based onto the padded z = (W | 0.. | u | X | 0..) of length 2 num_vars, and the
    transposes (for compute_eval_table_sparse: sum_row eq(rx)[row] M[row][col]) as CSR over 2 num_vars rows.  mats: [(row_ptr, col, val bytes)]"""
    fwd, tr = [], []
    for rp, col, val in mats:
        rp = np.asarray(rp, dtype=np.uint64)
        col = np.asarray(col, dtype=np.int64)
        colm = np.where(col < n_w, col, num_vars + (col - n_w))
        val = np.ascontiguousarray(val, dtype=np.uint8).reshape(-1, 32)
        nrows = len(rp) - 1
        fwd.append((nrows, rp, colm.astype(np.uint32), val.reshape(-1)))
        row_of = np.repeat(np.arange(nrows, dtype=np.int64), np.diff(rp.astype(np.int64)))
        order = np.argsort(colm, kind="stable")
        trp = np.concatenate([[0], np.cumsum(np.bincount(colm, minlength=2 * num_vars))]).astype(np.uint64)
        tr.append((2 * num_vars, trp, row_of[order].astype(np.uint32), val[order].reshape(-1)))
    return fwd, tr


class RelaxedR1CSProver:
    """Control flow of Arecibo's spartan::snark::RelaxedR1CSSNARK::prove (reached from `compress`, reference src/proof/nova.rs:341-356) over the
    C-ABI primitives, every vector device-resident: multiply_vec (SpMV x3), EqPolynomial::evals, the outer (cubic) and inner (quadratic)
    sum-checks, compute_eval_table_sparse (transposed SpMV x3 + AXPY x2) and the evaluation claims.  The Fiat-Shamir transcript is a
    callable `challenge(label, data) -> int`; the polynomial-commitment openings (hyperkzg_prove / ipa_prove) are separate calls."""

    def __init__(self, field_id, mats, n_w, n_x):
        self.field = field_id
        self.p = int.from_bytes(field_modulus(field_id), "little")
        self.n_w, self.n_x = n_w, n_x
        self.rows = len(mats[0][0]) - 1
        self.log_rows = max(1, (self.rows - 1).bit_length())
        self.num_vars = 1 << max(1, (max(n_w, n_x + 1) - 1).bit_length())
        fwd, tr = padded_and_transposed(mats, n_w, self.num_vars, 1 << self.log_rows)
        self.M = [DeviceCSR(field_id, *m) for m in fwd]
        self.MT = [DeviceCSR(field_id, *m) for m in tr]

    def _mont(self, x):
        return _fe(int(x) * (1 << 256) % self.p)

    def _axpy(self, a, b, r, out):
        _capi.check(_capi.lib().lurk_axpy_dev(self.field, C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), _capi.np_ptr(self._mont(r)), a.numel() // 32,
                                              C.c_void_p(out.data_ptr()), None))

    def pad_z(self, d_W, u, X):
        """(W | 0.. | u | X | 0..), Montgomery, 2 num_vars elements; d_W: device tensor of n_w Montgomery elements"""
        import torch
        z = torch.zeros(2 * self.num_vars * 32, dtype=torch.uint8, device="cuda")
        z[:self.n_w * 32] = d_W[:self.n_w * 32]
        tail = np.concatenate([self._mont(u)] + [self._mont(x) for x in X])
        z[self.num_vars * 32:self.num_vars * 32 + tail.size] = torch.from_numpy(tail).cuda()
        return z

    def prove(self, d_z, d_E, u, challenge, timings=None):
        """d_z: padded z (pad_z); d_E: device tensor of `rows` Montgomery elements.  Returns the transcript the verifier needs plus the points
        (rx, ry) at which E and W have to be opened."""
        import time
        import torch
        f, p, s, nv = self.field, self.p, self.log_rows, self.num_vars
        n_rows_pad = 1 << s
        t = time.perf_counter

        def mark(name, t0):
            if timings is not None:
                torch.cuda.synchronize()
                timings[name] = timings.get(name, 0.0) + (t() - t0) * 1e3

        t0 = t()
        Az, Bz, Cz = (torch.zeros(n_rows_pad * 32, dtype=torch.uint8, device="cuda") for _ in range(3))
        for M, y in zip(self.M, (Az, Bz, Cz)):
            M.mv(f, d_z.data_ptr(), y.data_ptr())
        E = torch.zeros(n_rows_pad * 32, dtype=torch.uint8, device="cuda")
        E[:self.rows * 32] = d_E[:self.rows * 32]
        uCzE = torch.empty_like(E)
        self._axpy(E, Cz, u, uCzE)
        mark("multiply_vec + u Cz + E", t0)
        t0 = t()
        tau = [challenge("tau", i) % p for i in range(s)]
        eq_tau = torch.empty(n_rows_pad * 32, dtype=torch.uint8, device="cuda")
        eq_evals(f, tau, eq_tau.data_ptr())
        mark("eq(tau)", t0)
        t0 = t()
        work = [eq_tau, Az.clone(), Bz.clone(), uCzE]
        outer_rounds, rx, fin = sumcheck_prove(f, CUBIC, [w.data_ptr() for w in work], s, 0,
                                               lambda rnd, msg: challenge("outer", (rnd, _ints(np.frombuffer(msg, dtype=np.uint8)))) % p)
        mark("outer sum-check", t0)
        t0 = t()
        eq_rx = torch.empty(n_rows_pad * 32, dtype=torch.uint8, device="cuda")
        eq_evals(f, rx, eq_rx.data_ptr())
        claims = (fin[1], fin[2], inner_product(f, Cz.data_ptr(), eq_rx.data_ptr(), n_rows_pad), inner_product(f, E.data_ptr(), eq_rx.data_ptr(), n_rows_pad))
        mark("claims at rx", t0)
        t0 = t()
        r = challenge("inner_r", claims) % p
        ys = [torch.empty(2 * nv * 32, dtype=torch.uint8, device="cuda") for _ in range(3)]
        for M, y in zip(self.MT, ys):
            M.mv(f, eq_rx.data_ptr(), y.data_ptr())
        abc = torch.empty_like(ys[0])
        self._axpy(ys[0], ys[1], r, abc)
        self._axpy(abc, ys[2], r * r % p, abc)
        mark("eval table (transposed SpMV)", t0)
        t0 = t()
        joint = (claims[0] + r * claims[1] + r * r * claims[2]) % p
        zc = d_z.clone()
        inner_rounds, ry, fin2 = sumcheck_prove(f, QUAD, [abc.data_ptr(), zc.data_ptr()], nv.bit_length(), joint,
                                                lambda rnd, msg: challenge("inner", (rnd, _ints(np.frombuffer(msg, dtype=np.uint8)))) % p)
        mark("inner sum-check", t0)
        t0 = t()
        eq_ry = torch.empty(nv * 32, dtype=torch.uint8, device="cuda")
        eq_evals(f, ry[1:], eq_ry.data_ptr())
        eval_W = inner_product(f, d_z.data_ptr(), eq_ry.data_ptr(), nv)
        mark("eval W", t0)
        return dict(outer_rounds=outer_rounds, inner_rounds=inner_rounds, claims=claims, eval_W=eval_W, rx=rx, ry=ry, E_padded=E)
