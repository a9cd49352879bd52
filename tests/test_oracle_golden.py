"""Pins the oracle (oracle/spec.py and oracle/oracle.c) on every golden vector the reference's own tests hold for
the Poseidon path and on the sizes it pins (witness sizes, bit-decomposition sizes)."""
import hashlib

import numpy as np
import pytest

from util import GOLDEN, TAG_CHAR, TAG_NUM, TAG_STR, TAG_SYM, ints, pack, random_elements

BN = 0


def h(oracle, field, pre, mode=1):
    return ints(oracle.poseidon_hash_batch(field, len(pre), pack(pre), mode=mode))[0]


@pytest.mark.parametrize("mode", [0, 1])
def test_trie_goldens(oracle, spec, mode):
    d = h(oracle, BN, [0] * 8, mode)
    assert d == GOLDEN["G1"] == spec.hash_correct(BN, [0] * 8)
    seq = [d]
    for _ in range(84):
        seq.append(h(oracle, BN, [seq[-1]] * 8, mode))
    assert seq[1] == GOLDEN["G2"] and seq[2] == GOLDEN["G3"] and seq[3] == GOLDEN["G4"]
    assert seq[84] == GOLDEN["G5"]          # empty StandardTrie root = H8 iterated 85 times from 0


def test_commitment_goldens(oracle, spec):
    assert h(oracle, BN, [0, TAG_NUM, 0]) == GOLDEN["G6"] == spec.hash_optimised(BN, [0, TAG_NUM, 0])
    assert h(oracle, BN, [0, TAG_NUM, 123]) == GOLDEN["G7"]


def lurk_str(oracle, s):
    acc = 0
    for ch in reversed(s):
        acc = h(oracle, BN, [TAG_CHAR, ord(ch), TAG_STR, acc])     # src/lem/store.rs:1368-1386
    return acc


def lurk_sym(oracle, path):
    acc = 0
    for name in path:
        acc = h(oracle, BN, [TAG_STR, lurk_str(oracle, name), TAG_SYM, acc])   # src/lem/store.rs:1389-1412
    return acc


def test_nil_commitment_golden_arity4_chain(oracle):
    nil = lurk_sym(oracle, ["lurk", "nil"])
    assert h(oracle, BN, [0, 0, nil]) == GOLDEN["G8"]             # (commit nil): Nil tag = 0


def test_structural_zero_tuples(oracle, spec):
    # G12: hash_ptr of tuple2/3/4 of zero atoms equals hash4/6/8 of zeros (src/lem/store.rs:1305-1338)
    for a in (4, 6, 8):
        assert h(oracle, BN, [0] * a) == spec.hash_correct(BN, [0] * a)


def test_witness_sizes_pinned_by_reference(oracle, spec):
    from util import REFERENCE
    sz = REFERENCE["sizes"]
    assert [oracle.witness_block(BN, a) for a in (4, 6, 8, 3)] == [sz["slot_witness_elements_bn256"][k] for k in ("Hash4", "Hash6", "Hash8", "Commitment")]
    assert [oracle.bitdecomp_size(f) for f in (2, 3, 0, 1)] == [sz["bit_decomp_witness_elements"][k] for k in ("pallas", "vesta", "bn256", "grumpkin")]
    # src/lem/multiframe.rs:991-1016 / src/lem/eval.rs:1960-1967: 14*293 + 6*396 + 268 + 3*354 = 7808 (BN256)
    assert [oracle.witness_block(BN, a) for a in (4, 6, 8, 3)] == [293, 343, 396, 268]
    assert 14 * 293 + 6 * 396 + 268 + 3 * oracle.bitdecomp_size(BN) == 7808
    # BIT_DECOMP_*_WITNESS_SIZE, src/lem/multiframe.rs:495-498 (Pallas, Vesta, BN256, Grumpkin)
    assert [oracle.bitdecomp_size(f) for f in (2, 3, 0, 1)] == [298, 301, 354, 364]
    assert [len(spec.bitdecomp_witness(f, 5)[0]) for f in (2, 3, 0, 1)] == [298, 301, 354, 364]


def test_survey_fingerprints(spec):
    # digest + sha256 over the aux values of the surveyor's independent probe
    fp = {
        (0, (0,) * 8): "1cefe00ba3c2b6ff56990084b0e3c943f0243034f65c1e9a125735d76b0a73f9",
        (0, (0, 4, 0)): "9f6f2d39663d2a4b585cc22e5274e4a5350a887e5cce2d8f2a6f438fe87bbe6f",
        (0, (1, 2, 3, 4)): "cd51b69c4f4de87998fd42e74a84fd9ab1aaffd0a7586133f2ee91ee21b1b5cd",
        (2, (0,) * 8): "2987485991fa415a02592a614b658cb4f604cb0ecec2f0584a3c12f6e64f6ec1",
        (2, (1, 2, 3, 4)): "dd957ded87f617df3015ec08a4937f40085748d7a72026f09a15de9068cd1530",
    }
    for (f, pre), want in fp.items():
        _, aux = spec.hash_optimised(f, list(pre), True)
        assert hashlib.sha256(b"".join(spec.fe_to_bytes(a) for a in aux)).hexdigest() == want
    assert spec.hash_correct(2, [0] * 8) == 0x0ef417527046e53c528056fe84bb984683b4610d346c2a33a81830c13a876b1c


@pytest.mark.parametrize("field", [0, 1, 2, 3])
@pytest.mark.parametrize("arity", [3, 4, 6, 8])
def test_c_oracle_matches_python_spec(oracle, spec, field, arity):
    pre = random_elements(field, 3 * arity, seed=100 + 10 * field + arity)
    rows = [ints(pre)[i * arity:(i + 1) * arity] for i in range(3)]
    d0 = ints(oracle.poseidon_hash_batch(field, arity, pre, mode=0))
    d1 = ints(oracle.poseidon_hash_batch(field, arity, pre, mode=1))
    w = ints(oracle.poseidon_witness_batch(field, arity, pre))
    blk = oracle.witness_block(field, arity)
    for i, r in enumerate(rows):
        assert d0[i] == d1[i] == spec.hash_correct(field, r)
        assert w[i * blk:(i + 1) * blk] == spec.slot_witness(field, r)


@pytest.mark.parametrize("field", [0, 1, 2, 3])
def test_bitdecomp_oracle(oracle, spec, field):
    p = spec.FIELD_MODULUS[field]
    vals = [0, 1, 2, p - 1, p // 2, (1 << 64) - 1] + ints(random_elements(field, 4, seed=7))
    w = ints(oracle.bitdecomp_witness_batch(field, pack(vals)))
    size = oracle.bitdecomp_size(field)
    for i, v in enumerate(vals):
        aux, bits = spec.bitdecomp_witness(field, v)
        assert w[i * size:(i + 1) * size] == aux
        assert sum(b << k for k, b in enumerate(bits)) == v


def test_noncanonical_rejected(oracle, spec):
    bad = pack([spec.FIELD_MODULUS[0]] + [0] * 7)
    with pytest.raises(ValueError):
        oracle.poseidon_hash_batch(0, 8, bad)


@pytest.mark.parametrize("curve", [0, 1, 2, 3])
def test_msm_oracle(oracle, spec, curve):
    C = spec.CURVES[curve]
    pb, q = spec.FIELD_MODULUS[C["base"]], spec.FIELD_MODULUS[C["scalar"]]
    assert spec.on_curve(curve, C["gen"])
    n = 24
    bases = oracle.gen_bases(curve, n)
    pts = list(zip(ints(bases)[0::2], ints(bases)[1::2]))
    assert pts[6] == spec.ec_mul(7, C["gen"], pb)
    sc = ints(random_elements(C["scalar"], n, seed=3))
    sc[0], sc[1], sc[2] = 0, 1, q - 1
    want = spec.msm_naive(curve, pts, sc)
    for naive in (True, False):
        o = oracle.msm(curve, bases, pack(sc), nthreads=2, naive=naive)
        assert (ints(o[:64])[0], ints(o[:64])[1]) == want and o[64] == 1
    # identity result
    o = oracle.msm(curve, bases, pack([0] * n))
    assert not o.any()


@pytest.mark.parametrize("field", [0, 2, 3])
def test_ntt_oracle(oracle, spec, field):
    a = ints(random_elements(field, 16, seed=9))
    got = oracle.ntt(field, pack(a))
    assert ints(got) == spec.ntt_naive(field, a)
    assert ints(oracle.ntt(field, got, inverse=True)) == a


def test_fold_helpers_oracle(oracle, spec):
    f, p = 0, spec.FIELD_MODULUS[0]
    a, b = ints(random_elements(f, 8, 1)), ints(random_elements(f, 8, 2))
    r = 0x1234567890abcdef1234567890abcdef
    assert ints(oracle.axpy(f, pack(a), pack(b), pack([r]))) == [(x + r * y) % p for x, y in zip(a, b)]
    row_ptr, col, val = [0, 2, 2, 5], [0, 3, 1, 2, 7], [3, p - 1, 5, 7, 11]
    y = ints(oracle.spmv(f, row_ptr, col, pack(val), pack(a)))
    assert y == [(3 * a[0] + (p - 1) * a[3]) % p, 0, (5 * a[1] + 7 * a[2] + 11 * a[7]) % p]
    v = [ints(random_elements(f, 4, 10 + k)) for k in range(6)]
    u1, u2 = 77, 1
    t = ints(oracle.cross_term(f, *[pack(x) for x in v], pack([u1]), pack([u2])))
    assert t == [(v[0][i] * v[4][i] + v[3][i] * v[1][i] - u1 * v[5][i] - u2 * v[2][i]) % p for i in range(4)]


def test_dag_oracle_matches_flat_hashes(oracle, spec):
    # tuple2 of (atom0, atom1) then a compact and a commitment over it (src/lem/store.rs:29-78 layouts)
    atoms = pack([11, 22, 33])
    nodes = np.zeros(3, dtype=oracle.DAG_NODE)
    nodes[0] = (2, 0, [5, 6, 0, 0], [0, 1, 0, 0])
    nodes[1] = (5, 0, [9, 7, 9, 9], [2, 3, 0, 0])      # compact: children atom2, node0, atom0; tags 0 and 2 dropped
    nodes[2] = (6, 0, [0, 4, 0, 0], [1, 4, 0, 0])      # commitment: secret atom1, payload (tag 4, node1)
    out = ints(oracle.dag_hash(0, nodes, atoms))
    d0 = spec.hash_correct(0, [5, 11, 6, 22])
    d1 = spec.hash_correct(0, [33, 7, d0, 11])
    assert out == [d0, d1, spec.hash_correct(0, [22, 4, d1])]


def test_trie_insert_golden(oracle):
    """G10: root of the StandardTrie (arity 8, height 85) after inserting 123 -> 456 (src/lem/tests/eval_tests.rs:3904,
    src/proof/tests/nova_tests.rs:4351); path = 3-bit chunks of the key, most significant first (trie/mod.rs:589-609,
    test_path: path(500) at height 3 = [7, 6, 4])."""
    height, arity, key, value = 85, 8, 123, 456
    path = lambda k, hgt: [(k >> (3 * (hgt - 1 - i))) & 7 for i in range(hgt)]
    assert path(500, 3) == [7, 6, 4]
    empty = [0]
    for _ in range(height):
        empty.append(h(oracle, BN, [empty[-1]] * arity))
    assert empty[85] == GOLDEN["G5"]
    # on the empty trie the node at depth d on any path has 8 children equal to empty[height - d - 1]
    v = value
    for depth in reversed(range(height)):
        pre = [empty[height - depth - 1]] * arity
        pre[path(key, height)[depth]] = v
        v = h(oracle, BN, pre)
    assert v == GOLDEN["G10"]


def test_lambda_commitment_golden_tuple4(oracle):
    """G9: (commit (lambda (x) x)).  Fun = tuple4 [vars, body, env, dummy] (src/lem/eval.rs:1158-1186: cons4(vars, body, env,
    foo); src/lem/store.rs:623-626) -> arity 8 with mixed tags; vars = ((x)), body = x, x = .lurk.user.x, empty env = (Env, 0)."""
    NIL, CONS, SYM, FUN, ENV = 0, 1, 2, 3, 12
    x = lurk_sym(oracle, ["lurk", "user", "x"])
    nil = lurk_sym(oracle, ["lurk", "nil"])
    vars_ = h(oracle, BN, [SYM, x, NIL, nil])
    fun = h(oracle, BN, [CONS, vars_, SYM, x, ENV, 0, NIL, 0])
    assert h(oracle, BN, [0, FUN, fun]) == GOLDEN["G9"]


class _Exprs:
    """Minimal expression builder over the oracle: z-pointers (tag, digest) of the shapes the reference's reader and REPL
    produce (src/lem/store.rs:481-505 symbols, :603-605 keywords, list = right fold of Cons over nil)."""
    NIL, CONS, FUN, NUM, KEY, ENV = 0, 1, 3, 4, 10, 12
    OUTERMOST, TERMINAL = 0x1000, 0x100E                           # src/tag.rs:126-146

    def __init__(self, oracle):
        self.o = oracle
        self.nil = (self.NIL, lurk_sym(oracle, ["lurk", "nil"]))
        self.env0 = (self.ENV, 0)

    def cons(self, a, b): return (self.CONS, h(self.o, BN, [a[0], a[1], b[0], b[1]]))
    def num(self, v): return (self.NUM, v)
    def sym(self, *path): return (TAG_SYM, lurk_sym(self.o, list(path)))
    def key(self, name): return (self.KEY, lurk_sym(self.o, [name]))

    def lst(self, items):
        acc = self.nil
        for it in reversed(items):
            acc = self.cons(it, acc)
        return acc

    def commit(self, secret, ptr): return h(self.o, BN, [secret, ptr[0], ptr[1]])

    def claim_hash(self, expr, env, expr_out, env_out):
        """src/cli/repl/mod.rs:263-296,332: list of keyword/value pairs, continuations as (Num tag . Num hash); Outermost and
        Terminal are continuation atoms whose hash is H8(0^8) (src/lem/eval.rs:1415-1418)"""
        cont = self.cons(self.num(self.OUTERMOST), self.num(GOLDEN["G1"]))
        cont_out = self.cons(self.num(self.TERMINAL), self.num(GOLDEN["G1"]))
        claim = self.lst([self.key("expr"), expr, self.key("env"), env, self.key("cont"), cont,
                          self.key("expr-out"), expr_out, self.key("env-out"), env_out, self.key("cont-out"), cont_out])
        return self.commit(0, claim)


def test_proof_claim_golden(oracle):
    """G11: the claim hash inside the proof key the CLI test expects for `!(prove (+ 1 1))` (tests/lurk-cli-tests.rs:58);
    the final environment is the empty one (erased, src/lem/eval.rs:1417)."""
    e = _Exprs(oracle)
    expr = e.lst([e.sym("lurk", "+"), e.num(1), e.num(1)])
    assert e.claim_hash(expr, e.env0, e.num(2), e.env0) == GOLDEN["G11"]


def test_documented_commitment_examples(oracle):
    """G13/G14: `!(commit '(13 . 21))` and `!(hide 12345 '(13 . 21))` (src/cli/repl/meta_cmd.rs:246-247,262-264): the second
    is the only reference vector with a non-zero secret.  G15: proof key of `!(prove '(1 2 3))` (:359-361)."""
    e = _Exprs(oracle)
    pair = e.cons(e.num(13), e.num(21))
    assert e.commit(0, pair) == GOLDEN["G13"]
    assert e.commit(12345, pair) == GOLDEN["G14"]
    l123 = e.lst([e.num(1), e.num(2), e.num(3)])
    assert e.claim_hash(e.lst([e.sym("lurk", "quote"), l123]), e.env0, l123, e.env0) == GOLDEN["G15"]


def test_functional_commitment_demo(oracle):
    """G16/G17 (demo/functional-commitment.lurk): commitment to f = (lambda (x) (+ (* 3 (* x x)) (+ (* 9 x) 2))) defined at the
    top level (closed over the empty env), and the claim of the proof of `!(call <G16> 5)`: expr = ((open <G16>) 5)
    (src/cli/repl/meta_cmd.rs:530-548) evaluated in the REPL env {f -> Fun}; an Env is a *compact* node
    H4[sym digest, val tag, val digest, env digest] (src/lem/store.rs:333-338, src/lem/store_core.rs:235-241)."""
    e = _Exprs(oracle)
    x, plus, mul = e.sym("lurk", "user", "x"), e.sym("lurk", "+"), e.sym("lurk", "*")
    body = e.lst([plus, e.lst([mul, e.num(3), e.lst([mul, x, x])]), e.lst([plus, e.lst([mul, e.num(9), x]), e.num(2)])])
    vars_ = e.lst([x])
    fun = (e.FUN, h(oracle, BN, [*vars_, *body, *e.env0, e.NIL, 0]))
    comm = e.commit(0, fun)
    assert comm == GOLDEN["G16"]
    env = (e.ENV, h(oracle, BN, [e.sym("lurk", "user", "f")[1], fun[0], fun[1], e.env0[1]]))
    expr = e.lst([e.lst([e.sym("lurk", "open"), e.num(comm)]), e.num(5)])
    assert e.claim_hash(expr, env, e.num(122), e.env0) == GOLDEN["G17"]


def _chained_counter(e, oracle):
    """the closures of demo/chained-functional-commitment.lurk.  `(add c)` is an undersaturated call of the two-argument
    recursive closure: Fun((x), body, env) with env = {counter -> c} on top of {add -> Rec} (src/lem/eval.rs:1470-1500);
    looking `add` up turns the Rec (the Fun's digest under tag 13, :1511-1517) back into a Fun whose environment binds
    `add` again (:1084-1090).  Returns head(c) -> commitment digest."""
    REC, sym = 13, e.sym
    counter, x, add = sym("lurk", "user", "counter"), sym("lurk", "user", "x"), sym("lurk", "user", "add")
    let, plus, cons_, commit = sym("lurk", "let"), sym("lurk", "+"), sym("lurk", "cons"), sym("lurk", "commit")
    body = e.lst([let, e.lst([e.lst([counter, e.lst([plus, counter, x])])]),
                  e.lst([cons_, counter, e.lst([commit, e.lst([add, counter])])])])
    foo = (e.NIL, 0)
    cons4 = lambda a, b, c, d: h(oracle, BN, [*a, *b, *c, *d])
    push = lambda s, v, env: (e.ENV, h(oracle, BN, [s[1], v[0], v[1], env[1]]))         # compact node
    rec = (REC, cons4(e.lst([counter, x]), body, e.env0, foo))
    rec_env = push(add, rec, e.env0)
    return lambda c: e.commit(0, (e.FUN, cons4(e.lst([x]), body, push(counter, e.num(c), rec_env), foo)))


def test_chained_functional_commitment_demo(oracle):
    """G18..G23 (demo/chained-functional-commitment.lurk): three links of the chain and the claims of their proofs; the
    input expression of a chain/call claim is ((open <Num hash>) arg) whether the head was given as a number or as a
    (comm ..) literal (src/cli/repl/meta_cmd.rs:538-539), the output is (total . (comm <next head>)) with Comm = tag 8."""
    COMM = 8
    e = _Exprs(oracle)
    head = _chained_counter(e, oracle)
    c0, c1, c2, c3 = head(0), head(9), head(21), head(35)
    assert (c0, c1, c2) == (GOLDEN["G18"], GOLDEN["G19"], GOLDEN["G20"])

    def claim(cin, arg, total, cout):
        expr = e.lst([e.lst([e.sym("lurk", "open"), e.num(cin)]), e.num(arg)])
        return e.claim_hash(expr, e.env0, e.cons(e.num(total), (COMM, cout)), e.env0)

    assert claim(c0, 9, 9, c1) == GOLDEN["G21"]
    assert claim(c1, 12, 21, c2) == GOLDEN["G22"]
    assert claim(c2, 14, 35, c3) == GOLDEN["G23"]


def test_protocol_and_chain_server_goldens(oracle):
    """G24 (demo/protocol.lurk) and G25..G27 (chain-server/README.md): `letrec` with two bindings nests -- `add` closes over
    the environment that already binds `sum` to its Rec -- and the head after each call differs only in the counter."""
    REC = 13
    e = _Exprs(oracle)
    assert e.commit(0, e.cons(e.num(13), e.num(17))) == GOLDEN["G24"]
    U, K = (lambda n: e.sym("lurk", "user", n)), (lambda n: e.sym("lurk", n))
    xs, acc, sum_, add, counter = U("xs"), U("acc"), U("sum"), U("add"), U("counter")
    body1 = e.lst([K("if"), e.lst([K("eq"), xs, e.nil]), acc,
                   e.lst([sum_, e.lst([K("cdr"), xs]), e.lst([K("+"), acc, e.lst([K("car"), xs])])])])
    body2 = e.lst([K("let"), e.lst([e.lst([counter, e.lst([K("+"), counter, e.lst([sum_, xs, e.num(0)])])])]),
                   e.lst([K("cons"), counter, e.lst([K("commit"), e.lst([add, counter])])])])
    foo = (e.NIL, 0)
    cons4 = lambda a, b, c, d: h(oracle, BN, [*a, *b, *c, *d])
    push = lambda s, v, env: (e.ENV, h(oracle, BN, [s[1], v[0], v[1], env[1]]))
    env1 = push(sum_, (REC, cons4(e.lst([xs, acc]), body1, e.env0, foo)), e.env0)
    rec_env = push(add, (REC, cons4(e.lst([counter, xs]), body2, env1, foo)), env1)
    head = lambda c: e.commit(0, (e.FUN, cons4(e.lst([xs]), body2, push(counter, e.num(c), rec_env), foo)))
    assert (head(0), head(7), head(37)) == (GOLDEN["G25"], GOLDEN["G26"], GOLDEN["G27"])


# (1, 2) + (1, 2) on alt_bn128, the first addition vector of the Ethereum bn256Add precompile tests (EIP-196): an
# independent known answer for the BN254 G1 group law (the reference holds no golden commitment)
BN254_2G = (0x030644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd3,
            0x15ed738c0e0a7c92e7845f96b2ae9c0a68a6a449e3538fc7ff3ebf7a5a18a2c4)


def test_bn254_g1_doubling_known_answer(oracle, spec):
    p = spec.FIELD_MODULUS[spec.CURVES[0]["base"]]
    g = spec.CURVES[0]["gen"]
    assert g == (1, 2) and spec.ec_mul(2, g, p) == BN254_2G and spec.ec_add(g, g, p) == BN254_2G
    bases = oracle.gen_bases(0, 2)                               # [1]G, [2]G
    assert tuple(ints(bases[64:128])) == BN254_2G
    assert tuple(ints(oracle.msm(0, bases[:64], pack([2]), naive=True))[:2]) == BN254_2G
    assert tuple(ints(oracle.msm(0, bases, pack([0, 1])))[:2]) == BN254_2G


def test_curve_generators_are_the_published_ones(spec):
    """the synthetic commitment keys are multiples of the standard generators: BN254 G1 (1, 2); Grumpkin (1, sqrt(-16)) with the
    y published for the Aztec / halo2curves `grumpkin::G1::generator()`; Pallas and Vesta (-1, 2) (pasta_curves)"""
    assert spec.CURVES[0]["gen"] == (1, 2) and spec.CURVES[0]["b"] == 3
    assert spec.CURVES[1]["gen"] == (1, 17631683881184975370165255887551781615748388533673675138860)
    assert spec.CURVES[1]["b"] == spec.FIELD_MODULUS[spec.CURVES[1]["base"]] - 17
    for c in (2, 3):
        p = spec.FIELD_MODULUS[spec.CURVES[c]["base"]]
        assert spec.CURVES[c]["gen"] == (p - 1, 2) and spec.CURVES[c]["b"] == 5
    for c in range(4):
        assert spec.on_curve(c, spec.CURVES[c]["gen"])
