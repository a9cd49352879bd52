"""CPU checks of the oracle's Nova folding restatement (oracle/nifs.py): the protocol-level property the reference's own
tests rely on (prove -> verify round trips, src/proof/tests/mod.rs:184-201): a folded relaxed-R1CS instance stays
satisfiable and its commitments stay consistent; and the pieces of the random oracle (SAFE IO tag, width-25 permutation,
optimised == textbook schedule)."""
import numpy as np

from oracle import capi, nifs, spec

CURVE = 0


def _fresh(rng, field, n_w, slot_cols, glue_fn, frames_per):
    p = spec.FIELD_MODULUS[field]
    W = [0] * n_w
    for c in slot_cols:
        W[c] = int(rng.integers(0, 2**62)) * int(rng.integers(0, 2**62)) % p
    for dst, v in glue_fn(W, p).items():
        W[dst] = v
    X = [int(rng.integers(1, 2**60)) for _ in range(2)]
    return nifs.pack(W), X


def test_sponge_io_tag_and_permutation_shapes():
    # [Absorb(9), Squeeze(1)]: tag < 2^128, depends on the pattern
    t9, t5 = spec.sponge_io_tag(9), spec.sponge_io_tag(5)
    assert t9 != t5 and t9 < 1 << 128
    # width 25: R_F = 8, R_P = 59 from Neptune's round-number search
    P = spec.params(1, 24)
    assert (P["t"], P["rf"], P["rp"]) == (25, 8, 59)
    # the optimised schedule (what the device sponge runs) equals the textbook permutation at width 25 too
    rng = np.random.default_rng(5)
    pre = [int(rng.integers(0, 2**62)) for _ in range(24)]
    assert spec.hash_optimised(1, pre) == spec.hash_correct(1, pre)
    c, h = spec.ro_squeeze(1, pre[:9])
    assert c == h & ((1 << 128) - 1) and h < spec.FIELD_MODULUS[1]


def test_three_folds_keep_the_relaxed_instance_satisfiable():
    rng = np.random.default_rng(11)
    frames, slot_elems, glue, lin = 2, 12, 5, 7
    mats, n_w, glue_fn = nifs.synthetic_step_circuit(rng, frames, slot_elems, glue, lin)
    rows = len(mats[0][0]) - 1
    bases = capi.gen_bases(CURVE, max(n_w, rows))
    o = nifs.NovaOracle(CURVE, bases, mats, n_w, 2, pp_digest=12345)
    per = slot_elems + glue
    slot_cols = [f * per + j for f in range(frames) for j in range(slot_elems)]
    W2, X2 = _fresh(rng, 0, n_w, slot_cols, glue_fn, per)
    o.init_running(W2, X2)
    assert o.bad_rows() == 0
    for _ in range(3):
        W2, X2 = _fresh(rng, 0, n_w, slot_cols, glue_fn, per)
        # the fresh instance alone satisfies the strict R1CS
        assert o.bad_rows(W2, np.zeros(rows * 32, dtype=np.uint8), 1, X2) == 0
        rec = o.prove_step(W2, X2)
        assert 0 < rec["r"] < 1 << 128
        assert o.bad_rows() == 0
        assert o.commitments_consistent() == (True, True)
    # a tampered witness is caught
    W = o.W.copy()
    W[0] ^= 1
    assert o.bad_rows(W=W) > 0 or o.commitments_consistent(W=W)[0] is False
