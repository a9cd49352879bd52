"""GPU parity for S2 (store hydration), S5 (fold helpers) and K6 (NTT) against the oracle."""
import numpy as np
import pytest

from util import GOLDEN, TAG_CHAR, TAG_NIL, TAG_STR, TAG_SYM, ints, pack, random_elements

pytestmark = pytest.mark.gpu


def test_store_symbol_hashing_golden(L):
    # (commit nil): nil = symbol .lurk.nil; strings are H4 chains of chars, symbols H4 chains of strings
    # (src/lem/store.rs:1368-1412); golden src/lem/tests/eval_tests.rs:1955
    s = L.StoreCore(L.FIELD_BN254_FR)
    zero_str = s.intern_atom(TAG_STR, 0)
    zero_sym = s.intern_atom(TAG_SYM, 0)

    def intern_str(text):
        ptr = zero_str
        for ch in reversed(text):
            ptr = s.intern_tuple2([s.intern_atom(TAG_CHAR, ord(ch)), ptr], TAG_STR)
        return ptr

    sym = zero_sym
    for name in ["lurk", "nil"]:
        sym = s.intern_tuple2([intern_str(name), sym], TAG_SYM)
    nil = (TAG_NIL, sym[1])
    assert len(s.dehydrated) == 9          # 7 string conses + 2 symbol conses, nothing hashed yet
    assert s.hide(0, nil) == GOLDEN["G8"]
    assert not s.dehydrated or all(v in s.z_cache for v in s.dehydrated)
    assert s.open(GOLDEN["G8"]) == (0, nil)


def test_store_basic_hashing_matches_flat_hashes(L):
    # mirror of test_basic_hashing (src/lem/store.rs:1305-1338)
    s = L.StoreCore(L.FIELD_BN254_FR)
    z = s.intern_atom(0, 0)
    t2 = s.intern_tuple2([z, z], 1)
    t3 = s.intern_tuple3([z, z, z], 1)
    t4 = s.intern_tuple4([z, z, z, z], 1)
    s.hydrate_z_cache()
    pc = s.hasher
    assert s.hash_ptr(t2)[1] == pc.hash4([0] * 4)
    assert s.hash_ptr(t3)[1] == pc.hash6([0] * 6)
    assert s.hash_ptr(t4)[1] == pc.hash8([0] * 8)
    env = s.intern_compact([z, t2, t3], 1)
    assert s.hash_ptr(env)[1] == pc.hash_compact(0, 1, s.hash_ptr_val(t2[1]), s.hash_ptr_val(t3[1]))


@pytest.mark.parametrize("field", [0, 2])
def test_random_dag_parity(L, oracle, field):
    rng = np.random.default_rng(12)
    n_atoms, n = 50, 3000
    atoms = random_elements(field, n_atoms, seed=3)
    nodes = np.zeros(n, dtype=oracle.DAG_NODE)
    kinds = [2, 3, 4, 5, 6]
    for i in range(n):
        k = kinds[rng.integers(0, 5)]
        nodes[i]["kind"] = k
        nodes[i]["tag"] = rng.integers(0, 0x3014, size=4)
        # mix of shallow and deep references; a long dependent chain every 7th node
        hi = n_atoms + i
        ch = rng.integers(0, hi, size=4)
        if i and i % 7 == 0:
            ch[0] = hi - 1
        nodes[i]["child"] = ch
    out = np.zeros(n * 32, dtype=np.uint8)
    L._capi.check(L._capi.lib().lurk_dag_hash(field, L._capi.np_ptr(nodes), n, L._capi.np_ptr(atoms), n_atoms, L._capi.np_ptr(out)))
    assert np.array_equal(out, oracle.dag_hash(field, nodes, atoms))
    # ordering violation is reported, not executed
    bad = nodes.copy()
    bad[0]["child"][0] = n_atoms + 5
    rc = L._capi.lib().lurk_dag_hash(field, L._capi.np_ptr(bad), n, L._capi.np_ptr(atoms), n_atoms, L._capi.np_ptr(out))
    assert rc == L._capi.ERR_ORDER


def dev(arr):
    import torch
    return torch.from_numpy(np.ascontiguousarray(arr)).cuda()


def mont(spec, field, buf):
    p = spec.FIELD_MODULUS[field]
    return pack([x * (1 << 256) % p for x in ints(buf)])


def unmont(spec, field, buf):
    p = spec.FIELD_MODULUS[field]
    rinv = pow(1 << 256, -1, p)
    return [x * rinv % p for x in ints(buf)]


@pytest.mark.parametrize("field", [0, 1, 2, 3])
def test_fold_helpers_parity(L, oracle, spec, field):
    import torch
    lib = L._capi.lib()
    n = 5000
    a, b = random_elements(field, n, 1, "witness"), random_elements(field, n, 2)
    r = random_elements(field, 1, 3)
    da, db = dev(mont(spec, field, a)), dev(mont(spec, field, b))
    out = torch.empty_like(da)
    L._capi.check(lib.lurk_axpy_dev(field, da.data_ptr(), db.data_ptr(), L._capi.np_ptr(mont(spec, field, r)), n, out.data_ptr(), None))
    assert unmont(spec, field, out.cpu().numpy()) == ints(oracle.axpy(field, a, b, r))
    # conversion kernel round trip
    canon = torch.empty_like(da)
    L._capi.check(lib.lurk_convert_dev(field, da.data_ptr(), n, L.FMT_CANONICAL, canon.data_ptr(), None))
    assert np.array_equal(canon.cpu().numpy(), a)
    # CSR SpMV with R1CS-like rows (0..6 non-zeros, small coefficients and -1)
    rng = np.random.default_rng(5)
    rows = 3000
    nnz_per = rng.integers(0, 7, size=rows)
    row_ptr = np.concatenate([[0], np.cumsum(nnz_per)]).astype(np.uint64)
    col = rng.integers(0, n, size=int(row_ptr[-1])).astype(np.uint32)
    p = spec.FIELD_MODULUS[field]
    val = pack([[1, p - 1, 2, 5, 7, (1 << 64) + 3][k] for k in rng.integers(0, 6, size=col.size)])
    y = torch.empty(rows * 32, dtype=torch.uint8, device="cuda")
    d_rp, d_col, d_val = dev(row_ptr), dev(col), dev(mont(spec, field, val))     # keep the device buffers alive
    L._capi.check(lib.lurk_spmv_csr_dev(field, d_rp.data_ptr(), d_col.data_ptr(), d_val.data_ptr(),
                                        rows, da.data_ptr(), y.data_ptr(), None))
    assert unmont(spec, field, y.cpu().numpy()) == ints(oracle.spmv(field, row_ptr, col, val, a))
    # cross term
    v = [random_elements(field, 1000, 20 + k) for k in range(6)]
    u1, u2 = random_elements(field, 1, 30), pack([1])
    dv = [dev(mont(spec, field, x)) for x in v]
    t = torch.empty_like(dv[0])
    L._capi.check(lib.lurk_cross_term_dev(field, *[x.data_ptr() for x in dv], L._capi.np_ptr(mont(spec, field, u1)),
                                          L._capi.np_ptr(mont(spec, field, u2)), 1000, t.data_ptr(), None))
    assert unmont(spec, field, t.cpu().numpy()) == ints(oracle.cross_term(field, *v, u1, u2))


@pytest.mark.parametrize("field", [0, 2, 3])
@pytest.mark.parametrize("log_n", [1, 4, 10, 11, 13, 16, 20, 21, 22])
def test_ntt_parity(L, oracle, spec, field, log_n):
    import torch
    lib = L._capi.lib()
    n = 1 << log_n
    a = random_elements(field, n, seed=log_n)
    d = dev(mont(spec, field, a))
    L._capi.check(lib.lurk_ntt_dev(field, d.data_ptr(), log_n, 0, None))
    got = unmont(spec, field, d.cpu().numpy())
    assert got == ints(oracle.ntt(field, a, nthreads=8))
    if log_n <= 4:
        assert got == spec.ntt_naive(field, ints(a))
    L._capi.check(lib.lurk_ntt_dev(field, d.data_ptr(), log_n, 1, None))     # inverse round trip
    assert unmont(spec, field, d.cpu().numpy()) == ints(a)


def test_ntt_rejects_unsupported_size(L):
    import torch
    d = torch.zeros(64, dtype=torch.uint8, device="cuda")
    assert L._capi.lib().lurk_ntt_dev(L.FIELD_BN254_FQ, d.data_ptr(), 2, 0, None) == L._capi.ERR_ARG   # 2-adicity 1


def test_trie_coprocessor_mirror_goldens(L, oracle):
    """trie coprocessor hashing through the CUDA path: empty roots (trie/mod.rs:925-1010), empty StandardTrie root,
    insert golden (eval_tests.rs:3868,3904), lookup, and the 85 batched lookup-circuit witnesses (a12)."""
    pc = L.PoseidonCache(L.FIELD_BN254_FR)
    small = L.Trie(pc, 8, 3)
    assert [small.empty_root_for_height(k) for k in (0, 1, 2, 3)] == [0, GOLDEN["G1"], GOLDEN["G2"], GOLDEN["G3"]]
    assert small.path(500) == [7, 6, 4]                                   # test_path
    assert L.Trie(pc, 8, 4).empty_root() == GOLDEN["G4"]
    t = L.StandardTrie(pc)
    assert t.root == t.empty_root() == GOLDEN["G5"]
    assert t.lookup(123) is None
    assert t.insert(123, 456) is True
    from util import GOLDEN as G
    assert t.root == G["G10"]
    assert t.lookup(123) == 456 and t.lookup(124) is None
    assert t.insert(123, 456) is False                                    # same value: root unchanged
    w = t.lookup_circuit_witnesses(123)
    assert w.size == 85 * 396 * 32                                        # 85 x hash8 slot blocks (a12: ~1.08 MB of aux)
    pre = pack([x for p in t.prove_lookup_at_path(t.path(123)) for x in p])
    assert np.array_equal(w, oracle.poseidon_witness_batch(0, 8, pre, nthreads=4))


def test_store_lambda_commitment_golden(L):
    """G9 through the store mirror on the GPU: a Fun is a tuple4 -> H8 (src/lem/store.rs:51-67,623-626)"""
    NIL, CONS, FUN, ENV = 0, 1, 3, 12
    s = L.StoreCore(L.FIELD_BN254_FR)
    zero_str, zero_sym = s.intern_atom(TAG_STR, 0), s.intern_atom(TAG_SYM, 0)

    def intern_sym(path):
        sym = zero_sym
        for name in path:
            st = zero_str
            for ch in reversed(name):
                st = s.intern_tuple2([s.intern_atom(TAG_CHAR, ord(ch)), st], TAG_STR)
            sym = s.intern_tuple2([st, sym], TAG_SYM)
        return sym

    x = intern_sym(["lurk", "user", "x"])
    nil = (NIL, intern_sym(["lurk", "nil"])[1])
    vars_ = s.intern_tuple2([x, nil], CONS)
    fun = s.intern_tuple4([vars_, x, s.intern_atom(ENV, 0), s.intern_atom(NIL, 0)], FUN)
    assert s.hide(0, fun) == GOLDEN["G9"]


class _StoreExprs:
    """expression builder over the store mirror (same shapes as tests/test_oracle_golden.py::_Exprs, but as store pointers
    hydrated on the GPU)"""
    NIL, CONS, FUN, NUM, KEY, ENV = 0, 1, 3, 4, 10, 12

    def __init__(self, L):
        self.s = s = L.StoreCore(L.FIELD_BN254_FR)
        self.zero_str, self.zero_sym = s.intern_atom(TAG_STR, 0), s.intern_atom(TAG_SYM, 0)
        self.nil = self.sym("lurk", "nil", tag=self.NIL)
        self.env0 = s.intern_atom(self.ENV, 0)

    def sym(self, *path, tag=TAG_SYM):
        s, sym = self.s, self.zero_sym
        for name in path:
            st = self.zero_str
            for ch in reversed(name):
                st = s.intern_tuple2([s.intern_atom(TAG_CHAR, ord(ch)), st], TAG_STR)
            sym = s.intern_tuple2([st, sym], TAG_SYM)
        return (tag, sym[1])

    def key(self, name): return self.sym(name, tag=self.KEY)
    def num(self, v): return self.s.intern_atom(self.NUM, v)
    def cons(self, a, b): return self.s.intern_tuple2([a, b], self.CONS)

    def lst(self, items):
        acc = self.nil
        for it in reversed(items):
            acc = self.cons(it, acc)
        return acc

    def claim(self, expr, env, expr_out, env_out):
        cont = self.cons(self.num(0x1000), self.num(GOLDEN["G1"]))
        cont_out = self.cons(self.num(0x100E), self.num(GOLDEN["G1"]))
        return self.lst([self.key("expr"), expr, self.key("env"), env, self.key("cont"), cont,
                         self.key("expr-out"), expr_out, self.key("env-out"), env_out, self.key("cont-out"), cont_out])


def test_store_proof_claim_golden(L):
    """G11 (tests/lurk-cli-tests.rs:58) through the store mirror on the GPU: the claim of `!(prove (+ 1 1))` is a DAG of
    H4 nodes (strings, symbol paths, keywords, conses) hydrated in one lurk_dag_hash call, then hidden with secret 0."""
    e = _StoreExprs(L)
    expr = e.lst([e.sym("lurk", "+"), e.num(1), e.num(1)])
    assert e.s.hide(0, e.claim(expr, e.env0, e.num(2), e.env0)) == GOLDEN["G11"]


def test_store_functional_commitment_goldens(L):
    """G13/G14 (hiding commitment with a non-zero secret) and G16/G17 (demo/functional-commitment.lurk): one DAG mixing
    tuple2 (H4), tuple4 (H8, the Fun) and a compact Env node (H4 with two tags dropped), hydrated on the GPU."""
    e = _StoreExprs(L)
    s = e.s
    pair = e.cons(e.num(13), e.num(21))
    assert s.hide(0, pair) == GOLDEN["G13"]
    assert s.hide(12345, pair) == GOLDEN["G14"]
    x, plus, mul = e.sym("lurk", "user", "x"), e.sym("lurk", "+"), e.sym("lurk", "*")
    body = e.lst([plus, e.lst([mul, e.num(3), e.lst([mul, x, x])]), e.lst([plus, e.lst([mul, e.num(9), x]), e.num(2)])])
    fun = s.intern_tuple4([e.lst([x]), body, e.env0, s.intern_atom(e.NIL, 0)], e.FUN)
    comm = s.hide(0, fun)
    assert comm == GOLDEN["G16"]
    env = s.intern_compact([e.sym("lurk", "user", "f"), fun, e.env0], e.ENV)
    expr = e.lst([e.lst([e.sym("lurk", "open"), e.num(comm)]), e.num(5)])
    assert s.hide(0, e.claim(expr, env, e.num(122), e.env0)) == GOLDEN["G17"]
