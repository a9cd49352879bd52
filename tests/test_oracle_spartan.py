"""The oracle's Spartan verifier (oracle/spartan.py) against a pure-Python prover that follows the same flow as the GPU prover of
tests/test_gpu_spartan_chain.py: accepts a genuinely folded relaxed-R1CS instance, rejects a tampered one.  CPU only."""
import numpy as np

from oracle import spartan as osp, sumcheck as sc
from test_gpu_spartan_chain import challenge, folded_instance, rows_of
from util import ints


def python_prover(R, n_w, nv, s, rows, W, E, u, X, p):
    z = W + [0] * (nv - n_w) + [u] + X + [0] * (nv - 1 - len(X))

    def mv(rowsl):
        return [sum(v * z[osp.col_map(c, n_w, nv)] for c, v in r) % p for r in rowsl] + [0] * ((1 << s) - rows)
    Az, Bz, Cz = [mv(r) for r in R]
    Ep = E + [0] * ((1 << s) - rows)
    uCzE = [(u * c + e) % p for c, e in zip(Cz, Ep)]
    tau = [challenge("tau", i) % p for i in range(s)]
    outer = sc.prove([sc.eq_evals(tau, p), Az, Bz, uCzE], "cubic", 0, lambda i, ev: challenge("outer", (i, ev)), p)
    rx = outer[1]
    eqrx = sc.eq_evals(rx, p)
    claims = (outer[2][1], outer[2][2], sc.inner_product(Cz, eqrx, p), sc.inner_product(Ep, eqrx, p))
    r = challenge("inner_r", claims) % p
    abc = [0] * (2 * nv)
    for k, rowsl in enumerate(R):
        for i, row in enumerate(rowsl):
            for c, v in row:
                j = osp.col_map(c, n_w, nv)
                abc[j] = (abc[j] + pow(r, k, p) * eqrx[i] * v) % p
    joint = (claims[0] + r * claims[1] + r * r * claims[2]) % p
    inner = sc.prove([abc, z], "quad", joint, lambda i, ev: challenge("inner", (i, ev)), p)
    eval_W = sc.mle_eval(W + [0] * (nv - n_w), inner[1][1:], p)
    return dict(outer_rounds=outer[0], inner_rounds=inner[0], claims=claims, eval_W=eval_W)


def test_verifier_accepts_folded_instance_and_rejects_tampering(oracle, spec):
    p = spec.FIELD_MODULUS[0]
    mats, n_w, o = folded_instance(oracle, spec, np.random.default_rng(31))
    rows = len(mats[0][0]) - 1
    s, nv = max(1, (rows - 1).bit_length()), 1 << max(1, (max(n_w, 3) - 1).bit_length())
    R = [rows_of(m) for m in mats]
    W, E = ints(o.W), ints(o.E)
    good = python_prover(R, n_w, nv, s, rows, W, E, o.u, o.X, p)
    assert osp.verify(R, n_w, nv, s, o.u, o.X, good, challenge, p)[0]
    E2 = list(E)
    E2[0] = (E2[0] + 1) % p
    assert not osp.verify(R, n_w, nv, s, o.u, o.X, python_prover(R, n_w, nv, s, rows, W, E2, o.u, o.X, p), challenge, p)[0]
    W2 = list(W)
    W2[3] = (W2[3] + 1) % p
    assert not osp.verify(R, n_w, nv, s, o.u, o.X, python_prover(R, n_w, nv, s, rows, W2, E, o.u, o.X, p), challenge, p)[0]
    assert not osp.verify(R, n_w, nv, s, (o.u + 1) % p, o.X, good, challenge, p)[0]
    lie = dict(good, eval_W=(good["eval_W"] + 1) % p)
    assert not osp.verify(R, n_w, nv, s, o.u, o.X, lie, challenge, p)[0]


def test_host_setup_of_padded_and_transposed_matrices(spec):
    """lurk-beta_b200/spartan.py: padded_and_transposed (numpy, no GPU): the re-based columns follow oracle col_map and the transposed CSR
    holds exactly the entries of the forward one"""
    import lurk_beta_b200 as L
    from oracle import nifs
    rng = np.random.default_rng(4)
    mats, n_w, _ = nifs.synthetic_step_circuit(rng, 2, 6, 3, 3)
    nv = 1 << max(1, (max(n_w, 3) - 1).bit_length())
    fwd, tr = L.spartan.padded_and_transposed(mats, n_w, nv, 64)
    for (rp, col, val), (nrows, frp, fcol, fval), (trows, trp, tcol, tval) in zip(mats, fwd, tr):
        assert nrows == len(rp) - 1 and trows == 2 * nv and int(trp[-1]) == int(rp[-1]) == len(fcol)
        assert [int(c) for c in fcol] == [osp.col_map(int(c), n_w, nv) for c in col]
        v = ints(val)
        entries = sorted((i, int(fcol[k]), v[k]) for i in range(nrows) for k in range(int(rp[i]), int(rp[i + 1])))
        tv = ints(tval)
        t_entries = sorted((int(tcol[k]), j, tv[k]) for j in range(trows) for k in range(int(trp[j]), int(trp[j + 1])))
        assert entries == t_entries
