"""
ORACLE (test infrastructure, NOT product code): ctypes front end of oracle/liboracle.so
(the plain-C restatement in oracle/oracle.c), with Poseidon constants installed from the
from-spec Python restatement oracle/spec.py.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from . import spec

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        _LIB = C.CDLL(path)
    return _LIB


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def fes_to_bytes(vals):
    return np.frombuffer(b"".join(spec.fe_to_bytes(v) for v in vals), dtype=np.uint8).copy()


_installed = set()


def install_params(field_id, arity):
    if (field_id, arity) in _installed:
        return
    P = spec.params(field_id, arity)
    t = P["t"]
    flat = lambda m: [x for row in m for x in row]
    arrs = [fes_to_bytes([P["domain_tag"]]), fes_to_bytes(P["rc"]), fes_to_bytes(flat(P["mds"])),
            fes_to_bytes(P["compressed"]), fes_to_bytes(flat(P["pre_sparse"])),
            fes_to_bytes([x for s in P["sparse"] for x in s["w_hat"]]),
            fes_to_bytes([x for s in P["sparse"] for x in s["v_rest"]])]
    rc = lib().oracle_set_poseidon_params(field_id, arity, P["rf"], P["rp"], *[_ptr(a) for a in arrs])
    assert rc == 0
    _installed.add((field_id, arity))


def threads():
    return lib().oracle_max_threads()


def witness_block(field_id, arity):
    P = spec.params(field_id, arity)
    return arity + 3 * (P["t"] * P["rf"] + P["rp"]) + 1


def poseidon_hash_batch(field_id, arity, pre, mode=1, nthreads=1):
    """pre: uint8 array n*arity*32 canonical LE -> uint8 array n*32"""
    install_params(field_id, arity)
    pre = np.ascontiguousarray(pre, dtype=np.uint8).reshape(-1)
    n = pre.size // (32 * arity)
    out = np.zeros(n * 32, dtype=np.uint8)
    rc = lib().oracle_poseidon_hash_batch(field_id, arity, _ptr(pre), C.c_size_t(n), _ptr(out), mode, nthreads)
    if rc:
        raise ValueError(f"oracle_poseidon_hash_batch rc={rc}")
    return out


def poseidon_witness_batch(field_id, arity, pre, nthreads=1):
    install_params(field_id, arity)
    pre = np.ascontiguousarray(pre, dtype=np.uint8).reshape(-1)
    n = pre.size // (32 * arity)
    out = np.zeros(n * 32 * witness_block(field_id, arity), dtype=np.uint8)
    rc = lib().oracle_poseidon_witness_batch(field_id, arity, _ptr(pre), C.c_size_t(n), _ptr(out), nthreads)
    if rc:
        raise ValueError(f"oracle_poseidon_witness_batch rc={rc}")
    return out


def bitdecomp_size(field_id):
    return lib().oracle_bitdecomp_size(field_id)


def bitdecomp_witness_batch(field_id, vals, nthreads=1):
    vals = np.ascontiguousarray(vals, dtype=np.uint8).reshape(-1)
    n = vals.size // 32
    out = np.zeros(n * 32 * bitdecomp_size(field_id), dtype=np.uint8)
    rc = lib().oracle_bitdecomp_witness_batch(field_id, _ptr(vals), C.c_size_t(n), _ptr(out), nthreads)
    if rc:
        raise ValueError(f"oracle_bitdecomp_witness_batch rc={rc}")
    return out


DAG_NODE = np.dtype([("kind", "u1"), ("pad", "u1"), ("tag", "<u2", (4,)), ("child", "<u4", (4,))], align=True)   # C layout, 28 bytes
assert DAG_NODE.itemsize == 28


def dag_hash(field_id, nodes, atoms):
    for a in (3, 4, 6, 8):
        install_params(field_id, a)
    nodes = np.ascontiguousarray(nodes, dtype=DAG_NODE)
    atoms = np.ascontiguousarray(atoms, dtype=np.uint8).reshape(-1)
    out = np.zeros(len(nodes) * 32, dtype=np.uint8)
    rc = lib().oracle_dag_hash(field_id, _ptr(nodes), C.c_size_t(len(nodes)), _ptr(atoms),
                               C.c_size_t(atoms.size // 32), _ptr(out))
    if rc:
        raise ValueError(f"oracle_dag_hash rc={rc}")
    return out


def msm(curve_id, bases, scalars, nthreads=1, naive=False):
    bases = np.ascontiguousarray(bases, dtype=np.uint8).reshape(-1)
    scalars = np.ascontiguousarray(scalars, dtype=np.uint8).reshape(-1)
    n = scalars.size // 32
    assert bases.size >= 64 * n
    out = np.zeros(96, dtype=np.uint8)
    if naive:
        rc = lib().oracle_msm_naive(curve_id, _ptr(bases), _ptr(scalars), C.c_size_t(n), _ptr(out))
    else:
        rc = lib().oracle_msm_pippenger(curve_id, _ptr(bases), _ptr(scalars), C.c_size_t(n), _ptr(out), nthreads)
    if rc:
        raise ValueError(f"oracle msm rc={rc}")
    return out


def gen_bases(curve_id, n, start=0):
    g = spec.CURVES[curve_id]["gen"]
    gen = fes_to_bytes(g)
    out = np.zeros(64 * n, dtype=np.uint8)
    rc = lib().oracle_gen_bases(curve_id, _ptr(gen), C.c_uint64(start), C.c_size_t(n), _ptr(out))
    assert rc == 0
    return out


def point_sum(curve_id, pts96):
    pts96 = np.ascontiguousarray(pts96, dtype=np.uint8).reshape(-1)
    out = np.zeros(96, dtype=np.uint8)
    rc = lib().oracle_point_sum(curve_id, _ptr(pts96), C.c_size_t(pts96.size // 96), _ptr(out))
    assert rc == 0
    return out


def axpy(field_id, a, b, r, nthreads=1):
    a = np.ascontiguousarray(a, dtype=np.uint8).reshape(-1)
    b = np.ascontiguousarray(b, dtype=np.uint8).reshape(-1)
    r = np.ascontiguousarray(r, dtype=np.uint8).reshape(-1)
    out = np.zeros_like(a)
    lib().oracle_axpy(field_id, _ptr(a), _ptr(b), _ptr(r), C.c_size_t(a.size // 32), _ptr(out), nthreads)
    return out


def spmv(field_id, row_ptr, col, val, z, nthreads=1):
    row_ptr = np.ascontiguousarray(row_ptr, dtype=np.uint64)
    col = np.ascontiguousarray(col, dtype=np.uint32)
    val = np.ascontiguousarray(val, dtype=np.uint8).reshape(-1)
    z = np.ascontiguousarray(z, dtype=np.uint8).reshape(-1)
    rows = len(row_ptr) - 1
    y = np.zeros(rows * 32, dtype=np.uint8)
    lib().oracle_spmv(field_id, _ptr(row_ptr), _ptr(col), _ptr(val), C.c_size_t(rows), _ptr(z), _ptr(y), nthreads)
    return y


def cross_term(field_id, az1, bz1, cz1, az2, bz2, cz2, u1, u2, nthreads=1):
    arrs = [np.ascontiguousarray(x, dtype=np.uint8).reshape(-1) for x in (az1, bz1, cz1, az2, bz2, cz2, u1, u2)]
    n = arrs[0].size // 32
    out = np.zeros(n * 32, dtype=np.uint8)
    lib().oracle_cross_term(field_id, *[_ptr(x) for x in arrs], C.c_size_t(n), _ptr(out), nthreads)
    return out


def ntt(field_id, data, inverse=False, nthreads=1):
    data = np.ascontiguousarray(data, dtype=np.uint8).reshape(-1).copy()
    n = data.size // 32
    log_n = n.bit_length() - 1
    p = spec.FIELD_MODULUS[field_id]
    w = spec.root_of_unity(field_id, log_n)
    if inverse:
        w = pow(w, p - 2, p)
    root = fes_to_bytes([w])
    lib().oracle_ntt(field_id, _ptr(data), log_n, _ptr(root), nthreads)
    if inverse:
        ninv = fes_to_bytes([pow(n, p - 2, p)])
        zero = np.zeros_like(data)
        data = axpy(field_id, zero, data, ninv, nthreads)
    return data


def from_label(curve_id, label, n, nthreads=1):
    """C port of DlogGroup::from_label (oracle.c: oracle_hash_to_curve_batch) -- the CPU baseline of the N3 row; the curve constants are
    the ones oracle/h2c.py computes / derives.  Returns n * 64 bytes (affine x | y canonical, identity = zeros)."""
    import hashlib
    from . import h2c
    p = h2c.base_modulus(curve_id)
    if h2c.CURVE_H2C_METHOD[curve_id] == "SVDW":
        method, consts = 0, [h2c.curve_b(curve_id), h2c.SVDW_Z % p, *h2c.svdw_constants(curve_id)]
    else:
        method, consts = 1, [h2c.curve_b(curve_id), h2c.SSWU_Z % p, h2c.ISO_A[curve_id], h2c.ISO_B, *h2c.isogeny_constants(curve_id)]
    cb = fes_to_bytes(consts)
    dst = np.frombuffer(h2c.dst_prime(curve_id, "from_uniform_bytes"), dtype=np.uint8).copy()
    msgs = np.frombuffer(hashlib.shake_256(bytes(label)).digest(32 * n), dtype=np.uint8).copy() if n else np.zeros(1, dtype=np.uint8)
    out = np.zeros(64 * n, dtype=np.uint8)
    rc = lib().oracle_hash_to_curve_batch(curve_id, method, _ptr(cb), _ptr(dst), C.c_size_t(dst.size), _ptr(msgs), C.c_size_t(32), C.c_size_t(n), _ptr(out),
                                          nthreads)
    if rc:
        raise ValueError(f"oracle_hash_to_curve_batch rc={rc}")
    return out


def sumcheck_prove(field_id, kind, polys_bytes, log_n, claim, challenge, nthreads=1):
    """C port of oracle/sumcheck.py: prove (oracle.c: oracle_sc_*), the CPU baseline of the N4 sum-check rows.
    polys_bytes: k buffers of 2^log_n canonical elements.  Returns (rounds, challenges, finals) like sumcheck.prove."""
    from . import sumcheck as sc
    p = spec.FIELD_MODULUS[field_id]
    k, E = (2, 2) if kind == "quad" else (4, 3)
    flat = np.concatenate([np.ascontiguousarray(b, dtype=np.uint8).reshape(-1) for b in polys_bytes])
    l = lib()
    l.oracle_sc_new.restype = C.c_void_p
    h = C.c_void_p(l.oracle_sc_new(field_id, 0 if kind == "quad" else 1, _ptr(flat), C.c_size_t(1 << log_n)))
    try:
        rounds, rs = [], []
        ev = np.zeros(32 * E, dtype=np.uint8)
        for rnd in range(log_n):
            l.oracle_sc_round(h, _ptr(ev), nthreads)
            e = [int.from_bytes(ev[32 * i:32 * i + 32].tobytes(), "little") for i in range(E)]
            evals = [e[0], (claim - e[0]) % p] + e[1:]
            r = challenge(rnd, evals) % p
            rounds.append(evals)
            rs.append(r)
            claim = sc.uni_eval_from_evals(evals, r, p)
            rc = l.oracle_sc_bind(h, _ptr(fes_to_bytes([r])), nthreads)
            assert rc == 0
        heads = np.zeros(32 * k, dtype=np.uint8)
        l.oracle_sc_heads(h, _ptr(heads))
        return rounds, rs, [int.from_bytes(heads[32 * i:32 * i + 32].tobytes(), "little") for i in range(k)]
    finally:
        l.oracle_sc_free(h)
