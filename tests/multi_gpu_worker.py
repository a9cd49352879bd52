"""Worker of tests/test_gpu_multi.py: one process per GPU (torchrun, NCCL for the setup only).

Every rank owns a contiguous range of frames of ONE step circuit: its slice of W, of the rows of A, B, C / E / T and of the
commitment key (reference layout: W index = frame * per + j, T index = frame * rows_per_frame + j), exactly how
north_star shards `commit` across GPUs.  The ranks fold an IVC chain through lurk_fold_ctx_* with world = N; the partial
commitments are exchanged inside the challenge kernel over peer memory.  Checked on every rank:
  * comm_W, comm_T, the challenge and the folded commitments of every step equal the oracle's UNSHARDED Nova fold;
  * the rank's slice of the folded W / E equals the oracle's slice, u and X agree;
  * check_running() (collective) reports a satisfied relaxed R1CS with consistent commitments;
  * rank 0 also runs the same chain on one GPU (world = 1) and gets identical records;
  * ShardedCommitmentKey.commit over NCCL equals the oracle's commitment;
  * a commitment key generated slice by slice on the ranks' GPUs (N3, lurk_ck_generate_range_dev) commits like the oracle's whole key.
Exit code 0 = all assertions passed on this rank."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def shard_circuit(mats, n_w, per, rows_per_frame, f0, f1):
    """rows of frames [f0, f1) with columns re-based to the rank's slice of W (the (u, X) tail follows the slice)"""
    out = []
    lo, hi = f0 * per, f1 * per
    for rp, col, val in mats:
        r0, r1 = f0 * rows_per_frame, f1 * rows_per_frame
        k0, k1 = int(rp[r0]), int(rp[r1])
        c = col[k0:k1].astype(np.int64)
        tail = c >= n_w
        assert np.all(tail | ((c >= lo) & (c < hi))), "frame-local columns expected"
        c = np.where(tail, c - n_w + (hi - lo), c - lo)
        out.append(((rp[r0:r1 + 1] - rp[r0]).astype(np.uint64), c.astype(np.uint32), val[k0 * 32:k1 * 32]))
    return out


def main():
    import torch
    import torch.distributed as dist
    import lurk_beta_b200 as L
    from oracle import capi as oracle, nifs, spec
    import test_gpu_fold_pipeline as T

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    CURVE, FIELD = 0, 0
    frames, glue, lin = 2 * world, 20, 16
    rng = np.random.default_rng(2024)                       # identical on every rank
    lay = T._layout(oracle, frames, glue)
    mats, n_w, glue_fn = nifs.synthetic_step_circuit(rng, frames, lay["slot_elems"], glue, lin)
    rows = len(mats[0][0]) - 1
    rpf = rows // frames
    bases = oracle.gen_bases(CURVE, max(n_w, rows))
    pp = 424242
    steps = [T._step_inputs(oracle, nifs, spec, lay, glue_fn, 200 + s, rng) for s in range(4)]
    o = nifs.NovaOracle(CURVE, bases, mats, n_w, 2, nthreads=8, pp_digest=pp)

    # ---- this rank's share
    f0, f1 = 2 * rank, 2 * rank + 2
    per = lay["per"]
    lmats = shard_circuit(mats, n_w, per, rpf, f0, f1)
    ln_w, lrows = (f1 - f0) * per, (f1 - f0) * rpf
    ck_w = L.CommitmentKey(CURVE, bases[f0 * per * 64:f1 * per * 64])
    ck_t = L.CommitmentKey(CURVE, bases[f0 * rpf * 64:f1 * rpf * 64])
    ctx = L.NovaFoldContext(CURVE, ck_w, ln_w, 2, lmats, depth=2, fmt=L.FMT_CANONICAL, ck_t=ck_t, world=world, rank=rank)
    llay = T._layout(oracle, f1 - f0, glue)
    bi = {a: ctx.add_slot_batch(a, llay["offs"][a]) for a, _ in llay["slots"]}
    bi[0] = ctx.add_slot_batch(0, llay["offs"][0])
    ctx.set_spans([(llay["slot_elems"], glue, per, f1 - f0)])
    ctx.connect()

    def local_inputs(st):
        """the rank's frames of a step's inputs (slot batches are frame-major)"""
        out = dict(pre={}, X2=st["X2"])
        for a, n in lay["slots"]:
            k = n // frames
            out["pre"][a] = st["pre"][a].reshape(n, a * 32)[f0 * k:f1 * k].reshape(-1)
        kb = lay["nbd"] // frames
        out["bd"] = st["bd"].reshape(lay["nbd"], 32)[f0 * kb:f1 * kb].reshape(-1)
        out["glue"] = st["glue"].reshape(frames, glue * 32)[f0:f1].reshape(-1)
        return out

    def run_chain(c, layc, bic, pick):
        recs = []
        T._fill(c, 0, layc, pick(steps[0]), pp, bic)
        c.stage_a(0)
        c.init_running(0)
        T._fill(c, 1, layc, pick(steps[1]), pp, bic)
        c.stage_a(1)
        recs.append(c.collect(0))
        for s in range(1, 4):
            b = s & 1
            c.stage_b_launch(b)
            recs.append(c.collect(b))
            if s + 1 < 4:
                T._fill(c, b ^ 1, layc, pick(steps[s + 1]), pp, bic)
                c.stage_a(b ^ 1)
        return recs

    recs = run_chain(ctx, llay, bi, local_inputs)
    want = [o.init_running(steps[0]["W2"], steps[0]["X2"])]
    assert T._same_point(nifs, recs[0].comm_W, want[0]["comm_W"]), "init comm_W"
    for s in range(1, 4):
        w = o.prove_step(steps[s]["W2"], steps[s]["X2"])
        r = recs[s]
        assert T._same_point(nifs, r.comm_W, w["comm_W"]), f"step {s} comm_W"
        assert T._same_point(nifs, r.comm_T, w["comm_T"]), f"step {s} comm_T"
        assert int.from_bytes(r.r.tobytes(), "little") == w["r"], f"step {s} challenge"
        if s == 3:
            assert T._same_point(nifs, r.running_comm_W, o.comm_W) and T._same_point(nifs, r.running_comm_E, o.comm_E), "folded commitments"
    run = ctx.get_running()
    assert np.array_equal(run["W"], o.W[f0 * per * 32:f1 * per * 32]), "W slice"
    assert np.array_equal(run["E"], o.E[f0 * rpf * 32:f1 * rpf * 32]), "E slice"
    assert nifs.ints(run["u"]) == [o.u] and nifs.ints(run["X"]) == o.X
    assert ctx.check_running() == (0, True, True), "device-side relaxed R1CS check (collective)"
    assert o.bad_rows() == 0

    # ---- one GPU, same chain: identical records (rank 0 only; the others wait at the barrier)
    if rank == 0:
        ck = L.CommitmentKey(CURVE, bases)
        c1 = L.NovaFoldContext(CURVE, ck, n_w, 2, mats, depth=2, fmt=L.FMT_CANONICAL)
        bi1 = {a: c1.add_slot_batch(a, lay["offs"][a]) for a, _ in lay["slots"]}
        bi1[0] = c1.add_slot_batch(0, lay["offs"][0])
        c1.set_spans([(lay["slot_elems"], glue, per, frames)])
        recs1 = run_chain(c1, lay, bi1, lambda st: st)
        for a, b in zip(recs, recs1):
            for k in ("comm_W", "comm_T", "r", "running_comm_W", "running_comm_E"):
                assert np.array_equal(getattr(a, k), getattr(b, k)), k

    # ---- the sharded commitment key over NCCL (commit.py), against the oracle
    sc = T.random_elements(FIELD, n_w, seed=77, shape="witness")
    lo, hi = L.shard_bounds(n_w, world, rank)
    sk = L.ShardedCommitmentKey(CURVE, bases[lo * 64:hi * 64], n_w)
    got = sk.commit(sc[lo * 32:hi * 32])
    assert np.array_equal(got, oracle.msm(CURVE, bases, sc, nthreads=8)), "ShardedCommitmentKey.commit"
    # ---- N3 on N GPUs: every rank GENERATES its slice of the reference's key (DlogGroup::from_label, points lo..hi) on its own GPU;
    # the sharded commitment over the generated slices equals the oracle's commitment over the oracle's restatement of the whole key
    from oracle import h2c
    import ctypes as C
    nk = 96 * world + 5
    klo, khi = L.shard_bounds(nk, world, rank)
    mine = L.CommitmentKey.setup(CURVE, b"ck", khi - klo, first=klo)
    whole = np.frombuffer(h2c.from_label_bytes(CURVE, b"ck", nk), dtype=np.uint8)
    ksc = T.random_elements(FIELD, nk, seed=78, shape="uniform")
    part = torch.from_numpy(mine.commit(ksc[klo * 32:khi * 32])).cuda()
    allp = torch.empty(96 * world, dtype=torch.uint8, device="cuda")
    dist.all_gather_into_tensor(allp, part)
    assert np.array_equal(L.point_sum(CURVE, allp.cpu().numpy()), oracle.msm(CURVE, whole, ksc, nthreads=8)), "sharded from_label key"
    dist.barrier()
    print(f"rank {rank}: multi-GPU fold parity ok", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
