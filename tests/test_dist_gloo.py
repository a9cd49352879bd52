"""world_size-2 (gloo, CPU) test of the N-GPU commitment combine: contiguous sharding of the key and the
all-gather + local point sum that stands in for an all-reduce (there is no reduction op for elliptic-curve
addition).  The per-rank partial commitments are produced by the oracle here (no GPU); the combine itself is the
product's code path (ShardedCommitmentKey.combine -> lurk_point_sum)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n, curve, ret, key_kind="synthetic"):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import lurk_beta_b200 as L
    from lurk_beta_b200.commit import ShardedCommitmentKey
    from oracle import capi as oracle
    from util import random_elements
    from oracle import spec
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        # "from_label": the reference's hash-to-curve key (N3), each rank owning the slice lurk_ck_generate_range_dev would generate for it
        bases = oracle.gen_bases(curve, n) if key_kind == "synthetic" else oracle.from_label(curve, b"ck", n, nthreads=2)
        scalars = random_elements(spec.CURVES[curve]["scalar"], n, seed=99, shape="witness")
        lo, hi = L.shard_bounds(n, world, rank)
        key = ShardedCommitmentKey.__new__(ShardedCommitmentKey)      # no GPU here: skip the device upload
        key.dist, key.group, key.curve_id, key.n_total = dist, None, curve, n
        key.rank, key.world, key.lo, key.hi = rank, world, lo, hi
        partial = oracle.msm(curve, bases[64 * lo:64 * hi], scalars[32 * lo:32 * hi])
        total = key.combine(partial)
        want = oracle.msm(curve, bases, scalars)
        ret[rank] = bool(np.array_equal(total, want))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [1001, 2])
def test_sharded_commit_combine_world2(n):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n, 2, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


def test_sharded_commit_over_a_from_label_key_world2():
    """the same combine over slices of the hash-to-curve key (oracle's C port stands in for the per-rank GPU generation)"""
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), 301, 0, ret, "from_label"), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


def test_shard_bounds_cover_exactly():
    sys.path.insert(0, ROOT)
    import lurk_beta_b200 as L
    for n in (0, 1, 7, 8, 1 << 21, (1 << 21) + 5):
        for w in (1, 2, 4, 8):
            spans = [L.shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
