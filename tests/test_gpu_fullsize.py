"""BASELINE.json full-size configurations on the GPU, checked through size-independent properties plus oracle parity on
sampled units (the oracle cannot finish these sizes in seconds):
  config 2: Poseidon batch hash of 2^22 preimages (arity 8 and arity 4 -- a cons cell is a Tuple2 -> arity 4)
            and hydration of a 2^22-node store DAG;
  config 3: Pedersen MSM over 2^24 Pallas bases."""
import numpy as np
import pytest

from util import ints, pack, random_elements

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("field,arity", [(0, 8), (0, 4), (2, 8)])
def test_poseidon_2_22(L, oracle, field, arity):
    n = 1 << 22
    pre = random_elements(field, n * arity, seed=2200 + field + arity, shape="lem")
    pc = L.PoseidonCache(field)
    dig = pc.hash_batch_bytes(arity, pre)
    # (a) oracle parity on a strided sample of 4096 hashes
    idx = np.arange(0, n, n // 4096)
    sample = pre.reshape(n, arity * 32)[idx].reshape(-1)
    assert np.array_equal(dig.reshape(n, 32)[idx].reshape(-1), oracle.poseidon_hash_batch(field, arity, sample, nthreads=8))
    # (b) permutation equivariance: hashing the reversed batch gives the reversed digests
    rev = np.ascontiguousarray(pre.reshape(n, arity * 32)[::-1]).reshape(-1)
    assert np.array_equal(pc.hash_batch_bytes(arity, rev).reshape(n, 32)[::-1], dig.reshape(n, 32))
    # (c) a small batch (latency launch shape) agrees with the persistent launch shape
    assert np.array_equal(pc.hash_batch_bytes(arity, pre[:1000 * arity * 32]), dig[:1000 * 32])


def test_dag_hydration_2_22_nodes(L, oracle):
    """2^22 cons cells: a forest of balanced binary trees over 2^16 atoms (22 levels deep at the tail chain)"""
    field = 0
    n_atoms, n = 1 << 16, 1 << 22
    rng = np.random.default_rng(5)
    atoms = random_elements(field, n_atoms, seed=6)
    nodes = np.zeros(n, dtype=oracle.DAG_NODE)
    nodes["kind"] = 2
    nodes["tag"][:, :2] = rng.integers(0, 16, size=(n, 2))
    hi = n_atoms + np.arange(n, dtype=np.int64)
    # children: uniformly among everything that precedes the node, biased to recent nodes (deep chains)
    back = rng.integers(1, 1 << 14, size=(n, 2))
    ch = np.maximum(hi[:, None] - back, 0)
    ch[:1024] = rng.integers(0, n_atoms, size=(1024, 2))
    nodes["child"][:, :2] = ch
    out = np.zeros(n * 32, dtype=np.uint8)
    L._capi.check(L._capi.lib().lurk_dag_hash(field, L._capi.np_ptr(nodes), n, L._capi.np_ptr(atoms), n_atoms, L._capi.np_ptr(out)))
    # every sampled node's digest must be H4(tag0, d(child0), tag1, d(child1)) of the digests the GPU produced
    table = np.concatenate([atoms, out]).reshape(-1, 32)
    idx = np.concatenate([np.arange(0, n, n // 4096), np.arange(n - 64, n)])
    pre = np.zeros((idx.size, 4, 32), dtype=np.uint8)
    pre[:, 0, :2] = nodes["tag"][idx, 0].astype("<u2").view(np.uint8).reshape(-1, 2)
    pre[:, 2, :2] = nodes["tag"][idx, 1].astype("<u2").view(np.uint8).reshape(-1, 2)
    pre[:, 1] = table[nodes["child"][idx, 0]]
    pre[:, 3] = table[nodes["child"][idx, 1]]
    want = oracle.poseidon_hash_batch(field, 4, pre.reshape(-1), nthreads=8)
    assert np.array_equal(out.reshape(n, 32)[idx].reshape(-1), want)
    # the first 2000 nodes against the sequential oracle walk
    assert np.array_equal(out[:2000 * 32], oracle.dag_hash(field, nodes[:2000], atoms))


def test_msm_2_24_pallas(L, oracle, spec):
    curve, n = 2, 1 << 24
    sf = spec.CURVES[curve]["scalar"]
    bases = L.synthetic_bases(curve, n)
    ck = L.CommitmentKey(curve, bases)
    a = random_elements(sf, n, seed=31, shape="witness")
    b = random_elements(sf, n, seed=32, shape="uniform")
    ab = oracle.axpy(sf, a, b, pack([1]), nthreads=8)
    ca, cb, cab = ck.commit(a), ck.commit(b), ck.commit(ab)
    assert ca[64] == 1 and cb[64] == 1
    assert spec.on_curve(curve, tuple(ints(ca[:64])))
    assert np.array_equal(L.point_sum(curve, np.concatenate([ca, cb])), cab)          # linearity
    # scaling: commit(3*a) = 3 * commit(a)
    a3 = oracle.axpy(sf, a, a, pack([2]), nthreads=8)
    assert np.array_equal(ck.commit(a3), L.point_sum(curve, np.concatenate([ca, ca, ca])))
    # prefix parity against the oracle (2^15 terms), and tail-only scalars (bases near the end of the key)
    m = 1 << 15
    assert np.array_equal(ck.commit(a[:32 * m]), oracle.msm(curve, bases[:64 * m], a[:32 * m], nthreads=8))
    tail = np.zeros(n * 32, dtype=np.uint8)
    tail[-32 * 100:] = b[:32 * 100]
    assert np.array_equal(ck.commit(tail), oracle.msm(curve, bases[-64 * 100:], b[:32 * 100]))
