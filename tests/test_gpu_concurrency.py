"""Thread-safety of the C ABI.  The reference calls these seams concurrently -- rayon workers hash store nodes and build slot
witnesses while the fold thread commits (src/proof/nova.rs:297-326, src/lem/multiframe.rs:579-584) -- so
every entry point must give the oracle's bytes when several host threads use it at once: the shared pinned staging pool
of the host-buffer Poseidon calls, one commitment context shared by several threads (internally serialised), private
contexts on different curves, store hydration, and transforms of the same size on different streams."""
import ctypes as C
import threading

import numpy as np
import pytest

from test_gpu_dag_fold import dev, mont, unmont
from util import ints, random_elements

pytestmark = pytest.mark.gpu


def run_threads(fns):
    errs = []

    def wrap(f):
        try:
            f()
        except BaseException as e:      # noqa: BLE001 - re-raised in the main thread
            errs.append(e)

    ts = [threading.Thread(target=wrap, args=(f,)) for f in fns]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    if errs:
        raise errs[0]


def test_concurrent_poseidon_host_calls(L, oracle):
    field = 0
    pc = L.PoseidonCache(field)
    jobs = []
    for t in range(8):
        arity = (3, 4, 6, 8)[t % 4]
        n = 2000 + 3500 * t if t < 4 else 200
        pre = random_elements(field, n * arity, seed=100 + t, shape="lem" if t & 1 else "uniform")
        if t < 4:
            jobs.append(("hash", arity, pre, oracle.poseidon_hash_batch(field, arity, pre, nthreads=4)))
        else:
            pre = pre[:200 * arity * 32]
            jobs.append(("witness", arity, pre, oracle.poseidon_witness_batch(field, arity, pre, nthreads=4)))
    got = [None] * len(jobs)
    st = {3: L.SlotType.Commitment, 4: L.SlotType.Hash4, 6: L.SlotType.Hash6, 8: L.SlotType.Hash8}

    def worker(i):
        kind, arity, pre, _ = jobs[i]
        for _rep in range(3):
            got[i] = pc.hash_batch_bytes(arity, pre) if kind == "hash" else L.slot_witness_batch_bytes(field, st[arity], pre)

    run_threads([lambda i=i: worker(i) for i in range(len(jobs))])
    for i, job in enumerate(jobs):
        assert np.array_equal(got[i], job[3]), job[:2]


def test_concurrent_commitments_shared_and_private_contexts(L, oracle):
    n = 3000
    shared_curve = 0
    bases = oracle.gen_bases(shared_curve, n)
    shared = L.CommitmentKey(shared_curve, bases)
    jobs = []
    for t in range(4):                                   # four threads on ONE context
        m = n - 173 * t
        sc = random_elements(0, m, seed=300 + t, shape="witness")      # scalar field of BN254 G1 = field 0
        jobs.append((shared, sc, oracle.msm(shared_curve, bases[:64 * m], sc, nthreads=4)))
    scalar_field = {1: 1, 2: 2, 3: 3}
    for curve in (1, 2, 3):                              # three threads with private contexts on other curves
        b = oracle.gen_bases(curve, 1500, start=curve)
        sc = random_elements(scalar_field[curve], 1500, seed=400 + curve)
        jobs.append((L.CommitmentKey(curve, b), sc, oracle.msm(curve, b, sc, nthreads=4)))
    got = [None] * len(jobs)

    def worker(i):
        ck, sc, _ = jobs[i]
        for _rep in range(3):
            got[i] = ck.commit(sc)

    run_threads([lambda i=i: worker(i) for i in range(len(jobs))])
    for i, job in enumerate(jobs):
        assert np.array_equal(got[i], job[2]), i


def test_concurrent_store_hydration(L, oracle):
    """independent stores hydrated from four threads (lurk_dag_hash allocates per call; the Poseidon constant cache is shared)"""
    def build(seed):
        import random
        rng = random.Random(seed)
        s = L.StoreCore(L.FIELD_BN254_FR)
        ptrs = [s.intern_atom(rng.randrange(15), rng.randrange(1 << 200)) for _ in range(10)]
        for _ in range(300):
            kind, k = rng.choice([("tuple2", 2), ("tuple3", 3), ("tuple4", 4), ("compact", 3)])
            ptrs.append(getattr(s, "intern_" + kind)([rng.choice(ptrs) for _ in range(k)], rng.randrange(15)))
        return s, ptrs[-1]

    stores = [build(7 + t) for t in range(4)]
    got = [None] * 4

    def worker(i):
        s, root = stores[i]
        got[i] = s.hash_ptr(root)

    run_threads([lambda i=i: worker(i) for i in range(4)])
    for i in range(4):                                   # same DAG hashed alone afterwards gives the same digest
        s2, root2 = build(7 + i)
        assert s2.hash_ptr(root2) == got[i]


def test_concurrent_ntt_same_size_different_streams(L, oracle, spec):
    """two transforms of one size in flight on different streams must not share scratch (stream-ordered allocation)"""
    import torch
    lib, field, log_n = L._capi.lib(), 0, 14
    n = 1 << log_n
    ins = [random_elements(field, n, seed=500 + t) for t in range(4)]
    want = [ints(oracle.ntt(field, a, nthreads=4)) for a in ins]
    bufs = [dev(mont(spec, field, a)) for a in ins]
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream() for _ in range(4)]

    def worker(i):
        for _rep in range(2):                            # forward, inverse, forward: ends as the forward transform
            L._capi.check(lib.lurk_ntt_dev(field, bufs[i].data_ptr(), log_n, 0, C.c_void_p(streams[i].cuda_stream)))
            L._capi.check(lib.lurk_ntt_dev(field, bufs[i].data_ptr(), log_n, 1, C.c_void_p(streams[i].cuda_stream)))
        L._capi.check(lib.lurk_ntt_dev(field, bufs[i].data_ptr(), log_n, 0, C.c_void_p(streams[i].cuda_stream)))
        streams[i].synchronize()

    run_threads([lambda i=i: worker(i) for i in range(4)])
    for i in range(4):
        assert unmont(spec, field, bufs[i].cpu().numpy()) == want[i], i
