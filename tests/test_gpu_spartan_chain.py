"""End-to-end property of the N4 row: the GPU half of `compress` for ONE relaxed R1CS instance -- the control flow of
Arecibo's RelaxedR1CSSNARK::prove over the C-ABI primitives (lurk-beta_b200/spartan.py: RelaxedR1CSProver) followed by the two HyperKZG
openings -- is accepted by the verifier's algebra (oracle/spartan.py + oracle/kzg.py with a key of known beta), for a running instance
that really went through two Nova folds (u != 1, E != 0); a tampered error vector is rejected.
Reference: src/proof/nova.rs:341-373 (compress / verify).  Challenges: a sha256 stand-in for Keccak256Transcript."""
import hashlib

import numpy as np
import pytest

from oracle import kzg, nifs, spartan as ospartan, sumcheck as sc
from util import ints, pack

pytestmark = pytest.mark.gpu
CURVE, FIELD = 0, 0


def challenge(label, data):
    return int.from_bytes(hashlib.sha256(repr((label, data)).encode()).digest() + hashlib.sha256(repr((data, label)).encode()).digest(), "little")


def to_device(L, field, canon_buf):
    import torch
    import ctypes as C
    t = torch.from_numpy(np.ascontiguousarray(canon_buf, dtype=np.uint8)).cuda()
    L._capi.check(L._capi.lib().lurk_convert_dev(field, C.c_void_p(t.data_ptr()), t.numel() // 32, L.FMT_MONTGOMERY, C.c_void_p(t.data_ptr()), None))
    return t


def rows_of(mat):
    rp, col, val = mat
    v = ints(val)
    return [[(int(col[k]), v[k]) for k in range(int(rp[i]), int(rp[i + 1]))] for i in range(len(rp) - 1)]


def folded_instance(oracle, spec, rng, frames=3, slot_elems=10, glue=5, lin=4):
    p = spec.FIELD_MODULUS[FIELD]
    mats, n_w, glue_fn = nifs.synthetic_step_circuit(rng, frames, slot_elems, glue, lin)
    rows = len(mats[0][0]) - 1
    bases = oracle.gen_bases(CURVE, max(n_w, rows))
    o = nifs.NovaOracle(CURVE, bases, mats, n_w, 2, pp_digest=11)
    for step in range(3):
        W = [int(x) % p for x in rng.integers(0, 2**62, size=n_w)]
        for dst, v in glue_fn(W, p).items():
            W[dst] = v
        X2 = [int(rng.integers(1, 2**60)), int(rng.integers(1, 2**60))]
        if step == 0:
            o.init_running(nifs.pack(W), X2)
        else:
            o.prove_step(nifs.pack(W), X2)
    assert o.bad_rows() == 0 and o.u != 1 and any(ints(o.E))
    return mats, n_w, o


def open_and_check(L, spec, ck, g, beta, d_poly, poly_ints, point, claimed_eval):
    """HyperKZG opening on the GPU + the verifier's algebra with the commitments' discrete logs (key beta^i g)"""
    pb, p = spec.FIELD_MODULUS[spec.CURVES[CURVE]["base"]], spec.FIELD_MODULUS[FIELD]
    log = []

    def cb(rnd, msg):
        log.append(bytes(msg))
        return challenge("pcs", (rnd, bytes(msg))) % p
    com, v, w = L.spartan.hyperkzg_prove(CURVE, ck, d_poly.data_ptr(), point, cb)
    polys = kzg.fold_chain(poly_ints, point, p)
    at_beta = [kzg.poly_eval(f, beta, p) for f in polys]
    assert com == [spec.ec_mul(s, g, pb) for s in at_beta[1:]]
    r, q = challenge("pcs", (0, log[0])) % p, challenge("pcs", (1, log[1])) % p
    u = [r, (-r) % p, r * r % p]
    Bbeta = sum(pow(q, j, p) * s for j, s in enumerate(at_beta)) % p
    w_scalars = [(Bbeta - sum(pow(q, j, p) * v[t][j] for j in range(len(point))) % p) * pow(beta - u[t], -1, p) % p for t in range(3)]
    assert w == [spec.ec_mul(s, g, pb) for s in w_scalars]
    return kzg.verify_known_beta(CURVE, g, beta, at_beta[0], point, claimed_eval, at_beta[1:], v, w_scalars, r, q)


def test_compress_chain_is_accepted_by_the_verifier(L, oracle, spec):
    import torch
    p = spec.FIELD_MODULUS[FIELD]
    pb = spec.FIELD_MODULUS[spec.CURVES[CURVE]["base"]]
    rng = np.random.default_rng(31)
    mats, n_w, o = folded_instance(oracle, spec, rng)
    prover = L.spartan.RelaxedR1CSProver(FIELD, mats, n_w, 2)
    rows_lists = [rows_of(m) for m in mats]
    dW, dE = to_device(L, FIELD, o.W), to_device(L, FIELD, o.E)
    z = prover.pad_z(dW, o.u, o.X)
    timings = {}
    proof = prover.prove(z, dE, o.u, challenge, timings)
    ok, rx, ry = ospartan.verify(rows_lists, n_w, prover.num_vars, prover.log_rows, o.u, o.X, proof, challenge, p)
    assert ok and rx == proof["rx"] and ry == proof["ry"]
    # the claimed evaluations are the multilinear extensions of the padded vectors
    Wp = ints(o.W) + [0] * (prover.num_vars - n_w)
    Ep = ints(o.E) + [0] * ((1 << prover.log_rows) - prover.rows)
    assert proof["eval_W"] == sc.mle_eval(Wp, ry[1:], p) and proof["claims"][3] == sc.mle_eval(Ep, rx, p)
    # polynomial-commitment openings of W at ry[1:] and of E at rx under a powers-of-tau key of known beta
    g = spec.ec_mul(4242, spec.CURVES[CURVE]["gen"], pb)
    beta = 0x1234567890abcdef1234567890abcdef % p
    ck = L.CommitmentKey.powers_of_tau(CURVE, g, beta, max(prover.num_vars, 1 << prover.log_rows))
    assert open_and_check(L, spec, ck, g, beta, z[:prover.num_vars * 32].clone(), Wp, ry[1:], proof["eval_W"])
    assert open_and_check(L, spec, ck, g, beta, proof["E_padded"], Ep, rx, proof["claims"][3])
    # a wrong evaluation claim is not accepted by the opening check
    assert not open_and_check(L, spec, ck, g, beta, proof["E_padded"], Ep, rx, (proof["claims"][3] + 1) % p)
    # soundness smoke test: one tampered row of E (the instance no longer satisfies the relaxed R1CS) -> the verifier rejects
    bad_E = o.E.copy()
    bad_E[0] ^= 1
    proof_bad = prover.prove(z, to_device(L, FIELD, bad_E), o.u, challenge)
    assert not ospartan.verify(rows_lists, n_w, prover.num_vars, prover.log_rows, o.u, o.X, proof_bad, challenge, p)[0]
    # and a transcript replayed against another instance (u changed) is rejected as well
    assert not ospartan.verify(rows_lists, n_w, prover.num_vars, prover.log_rows, (o.u + 1) % p, o.X, proof, challenge, p)[0]
    assert set(timings) >= {"outer sum-check", "inner sum-check"}


def test_spmv_with_very_long_rows(L, oracle, spec):
    """the transposed R1CS matrices have a few rows with 10^4..10^5 entries (the columns of u and X): lurk_spmv_csr_dev hands rows
    longer than 1024 non-zeros to whole CTAs; result equals the oracle's SpMV"""
    import ctypes as C
    import torch
    from util import random_elements
    rng = np.random.default_rng(3)
    ncols = 3000
    lens = [2, 0, 5000, 1, 1024, 1025, 3, 70000, 0, 2]
    rp = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    col = rng.integers(0, ncols, size=int(rp[-1])).astype(np.uint32)
    val = random_elements(FIELD, int(rp[-1]), seed=1)
    zv = random_elements(FIELD, ncols, seed=2)
    want = oracle.spmv(FIELD, rp, col, val, zv, nthreads=4)
    M = L.spartan.DeviceCSR(FIELD, len(lens), rp, col, val)
    dz = to_device(L, FIELD, zv)
    y = torch.zeros(len(lens) * 32, dtype=torch.uint8, device="cuda")
    M.mv(FIELD, dz.data_ptr(), y.data_ptr())
    L._capi.check(L._capi.lib().lurk_convert_dev(FIELD, C.c_void_p(y.data_ptr()), len(lens), L.FMT_CANONICAL, C.c_void_p(y.data_ptr()), None))
    torch.cuda.synchronize()
    assert np.array_equal(y.cpu().numpy(), want)
