"""CPU tests of the store mirror's HOST logic (interning, children-first flattening into lurk_dag_node records, digest
tables, commitments) against the reference's goldens.  There is no GPU here, so the two device entry points the mirror
calls are replaced -- in this test module only -- by a stand-in that hands the very same buffers to the oracle; what
is under test is everything the mirror does before and after that call.  The same expressions run against the real
library in tests/test_gpu_dag_fold.py."""
import ctypes as C

import numpy as np
import pytest

from test_gpu_dag_fold import _StoreExprs
from util import GOLDEN


class _OracleBackedLib:
    """stand-in for liblurk_b200 exposing only lurk_dag_hash and lurk_poseidon_hash_batch with the C-ABI's argument order"""

    def __init__(self, oracle):
        self.o = oracle
        self.dag_calls = 0

    @staticmethod
    def _view(ptr, nbytes):
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(nbytes,)) if nbytes else np.zeros(0, np.uint8)

    def lurk_dag_hash(self, field, nodes, n, atoms, n_atoms, out):
        self.dag_calls += 1
        nodes_a = self._view(nodes, n * self.o.DAG_NODE.itemsize).view(self.o.DAG_NODE)
        self._view(out, n * 32)[:] = self.o.dag_hash(field, nodes_a, self._view(atoms, n_atoms * 32))
        return 0

    def lurk_poseidon_hash_batch(self, field, arity, pre, n, out):
        self._view(out, n * 32)[:] = self.o.poseidon_hash_batch(field, arity, self._view(pre, n * arity * 32))
        return 0


@pytest.fixture()
def HL(monkeypatch, oracle):
    import lurk_beta_b200 as L
    from lurk_beta_b200 import _capi
    fake = _OracleBackedLib(oracle)
    monkeypatch.setattr(_capi, "lib", lambda: fake)
    L._fake = fake
    return L


def test_claim_golden_through_store_flattening(HL):
    e = _StoreExprs(HL)
    expr = e.lst([e.sym("lurk", "+"), e.num(1), e.num(1)])
    claim = e.claim(expr, e.env0, e.num(2), e.env0)
    assert e.s.hide(0, claim) == GOLDEN["G11"]
    assert HL._fake.dag_calls == 1                      # the whole claim DAG went down in one call
    assert e.s.hide(0, claim) == GOLDEN["G11"] and HL._fake.dag_calls == 1   # z_cache hit, nothing re-hashed


def test_functional_commitment_goldens_through_store_flattening(HL):
    e = _StoreExprs(HL)
    s = e.s
    pair = e.cons(e.num(13), e.num(21))
    assert s.hide(0, pair) == GOLDEN["G13"] and s.hide(12345, pair) == GOLDEN["G14"]
    assert s.open(GOLDEN["G14"]) == (12345, pair)
    x, plus, mul = e.sym("lurk", "user", "x"), e.sym("lurk", "+"), e.sym("lurk", "*")
    body = e.lst([plus, e.lst([mul, e.num(3), e.lst([mul, x, x])]), e.lst([plus, e.lst([mul, e.num(9), x]), e.num(2)])])
    fun = s.intern_tuple4([e.lst([x]), body, e.env0, s.intern_atom(e.NIL, 0)], e.FUN)
    comm = s.hide(0, fun)
    assert comm == GOLDEN["G16"]
    env = s.intern_compact([e.sym("lurk", "user", "f"), fun, e.env0], e.ENV)
    expr = e.lst([e.lst([e.sym("lurk", "open"), e.num(comm)]), e.num(5)])
    # children hashed by the earlier calls (the Fun, symbol paths) travel as extra digests, not as nodes
    assert s.hide(0, e.claim(expr, env, e.num(122), e.env0)) == GOLDEN["G17"]


def test_hydrate_queue_matches_on_demand_hashing(HL):
    """hydrate_z_cache over the interning-ordered queue (store_core.rs:266-269) and on-demand hash_ptr give equal digests"""
    a, b = _StoreExprs(HL), _StoreExprs(HL)
    mk = lambda e: e.lst([e.sym("lurk", "user", "abc"), e.num(7), e.lst([e.num(1), e.key("k")])])
    pa, pb = mk(a), mk(b)
    a.s.hydrate_z_cache()
    assert not a.s.dehydrated and a.s.z_cache[pa[1]] == b.s.hash_ptr(pb)[1]
