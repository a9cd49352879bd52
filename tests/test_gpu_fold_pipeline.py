"""GPU parity of the device-resident fold pipeline (lurk_beta_b200/fold.py, the GPU half of prove_recursively,
src/proof/nova.rs:260-339) against the oracle, step by step: slot witnesses written in place into W2, commit(W2),
Az/Bz/Cz for both instances, cross term T, commit(T), and the fold W1 <- W1 + r W2, E1 <- E1 + r T -- over two
consecutive steps so that the prefetch (stage A one step ahead) and the double buffering are exercised."""
import hashlib

import numpy as np
import pytest

from util import ints, pack, random_elements

pytestmark = pytest.mark.gpu
FIELD, CURVE = 0, 0
R = 1 << 256


def mont(spec, buf):
    p = spec.FIELD_MODULUS[FIELD]
    return pack([x * R % p for x in ints(buf)])


def unmont(spec, buf):
    p = spec.FIELD_MODULUS[FIELD]
    rinv = pow(R, -1, p)
    return pack([x * rinv % p for x in ints(buf)])


def test_two_pipelined_folds_match_oracle(L, oracle, spec):
    import torch
    from lurk_beta_b200.fold import NovaFoldPipeline, SlotBatch
    rng = np.random.default_rng(42)
    p = spec.FIELD_MODULUS[FIELD]
    frames = 3
    slots = [(4, 14 * frames), (8, 6 * frames), (3, 1 * frames)]
    nbd = 3 * frames
    blocks = {a: oracle.witness_block(FIELD, a) for a, _ in slots}
    bd_block = oracle.bitdecomp_size(FIELD)
    slot_region = sum(n * blocks[a] for a, n in slots) + nbd * bd_block
    assert slot_region == 7808 * frames
    glue = 500
    n_w, n_t = slot_region + glue, 3000
    ncols = n_w + 3
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()

    # commitment key and R1CS (canonical on the host for the oracle, Montgomery on the device)
    bases = oracle.gen_bases(CURVE, max(n_w, n_t))
    ck = L.CommitmentKey(CURVE, bases).precompute()
    mats = []
    for seed in (1, 2, 3):
        nnz_per = rng.integers(1, 4, size=n_t)
        row_ptr = np.concatenate([[0], np.cumsum(nnz_per)]).astype(np.uint64)
        col = rng.integers(0, ncols, size=int(row_ptr[-1])).astype(np.uint32)
        val = pack([[1, p - 1, 2, 7][k] for k in rng.integers(0, 4, size=col.size)])
        mats.append((row_ptr, col, val))
    d_mats = [(dev(rp), dev(col), dev(mont(spec, val))) for rp, col, val in mats]
    u1, u2 = random_elements(FIELD, 1, 5), pack([1])
    tail = random_elements(FIELD, 3, 6)
    W1 = random_elements(FIELD, n_w, 7)
    E1 = random_elements(FIELD, n_t, 8)
    z1 = dev(mont(spec, np.concatenate([W1, tail])))
    z2 = [dev(mont(spec, np.concatenate([np.zeros(n_w * 32, dtype=np.uint8), tail]))) for _ in range(2)]
    dE1 = dev(mont(spec, E1))
    pipe = NovaFoldPipeline(torch, FIELD, CURVE, ck, n_w, n_t, d_mats, mont(spec, u1), mont(spec, u2), z1, dE1, z2)

    def challenge(cw, ct):
        r = np.zeros(32, dtype=np.uint8)
        r[:16] = np.frombuffer(hashlib.sha256(cw.tobytes() + ct.tobytes()).digest()[:16], dtype=np.uint8)
        return r                                   # used as a Montgomery-form scalar by the pipeline

    steps = []
    for s in range(2):                             # per-step inputs: slot preimages (some dummy) and glue aux
        pre = {}
        for a, n in slots:
            x = random_elements(FIELD, n * a, seed=100 * s + a, shape="lem").reshape(n, a * 32)
            x[rng.random(n) < 0.6] = 0
            pre[a] = x.reshape(-1)
        steps.append(dict(pre=pre, bd=random_elements(FIELD, nbd, seed=50 + s, shape="witness"),
                          glue=random_elements(FIELD, glue, seed=60 + s, shape="witness")))

    d_pre = {a: torch.empty(n * a * 32, dtype=torch.uint8, device="cuda") for a, n in slots}
    d_bd = torch.empty(nbd * 32, dtype=torch.uint8, device="cuda")
    batches, off = [], 0
    for a, n in slots:
        batches.append(SlotBatch(a, n, off, d_pre[a]))
        off += n * blocks[a]
    batches.append(SlotBatch(0, nbd, off, d_bd))

    def stage_inputs_for(step):
        def before(b):
            for a, _ in slots:
                d_pre[a].copy_(dev(mont(spec, step["pre"][a])))
            d_bd.copy_(dev(mont(spec, step["bd"])))
            pipe.W2[b][slot_region * 32:].copy_(dev(mont(spec, step["glue"])))
        return before

    # oracle state
    oW1, oE1 = W1.copy(), E1.copy()
    pipe.stage_a(0, batches, stage_inputs_for(steps[0]))
    for s in range(2):
        b = s & 1
        if s + 1 < 2:
            torch.cuda.synchronize()               # inputs are staged through shared device buffers in this test
            pipe.stage_a(b ^ 1, batches, stage_inputs_for(steps[s + 1]))
        cw, ct = pipe.stage_b(b, challenge)
        torch.cuda.synchronize()
        # ---- the same step on the oracle
        st = steps[s]
        parts = [oracle.poseidon_witness_batch(FIELD, a, st["pre"][a], nthreads=4) for a, _ in slots]
        parts.append(oracle.bitdecomp_witness_batch(FIELD, st["bd"]))
        oW2 = np.concatenate(parts + [st["glue"]])
        assert np.array_equal(unmont(spec, pipe.W2[b].cpu().numpy()), oW2), f"step {s}: W2"
        want_cw = oracle.msm(CURVE, bases, oW2, nthreads=8)
        oz1, oz2 = np.concatenate([oW1, tail]), np.concatenate([oW2, tail])
        mv = [oracle.spmv(FIELD, rp, col, val, z) for (rp, col, val) in mats for z in (oz1, oz2)]
        az1, az2, bz1, bz2, cz1, cz2 = mv
        oT = oracle.cross_term(FIELD, az1, bz1, cz1, az2, bz2, cz2, u1, u2)
        assert np.array_equal(unmont(spec, pipe.T.cpu().numpy()), oT), f"step {s}: T"
        want_ct = oracle.msm(CURVE, bases, oT, nthreads=8)
        pb = spec.FIELD_MODULUS[spec.CURVES[CURVE]["base"]]
        rinv = pow(R, -1, pb)
        assert [v * rinv % pb for v in ints(cw[:64])] == ints(want_cw[:64]), f"step {s}: comm_W"
        assert [v * rinv % pb for v in ints(ct[:64])] == ints(want_ct[:64]), f"step {s}: comm_T"
        r_canon = pack([ints(challenge(cw, ct))[0] * pow(R, -1, p) % p])
        oW1 = oracle.axpy(FIELD, oW1, oW2, r_canon)
        oE1 = oracle.axpy(FIELD, oE1, oT, r_canon)
        assert np.array_equal(unmont(spec, pipe.W1.cpu().numpy()), oW1), f"step {s}: folded W"
        assert np.array_equal(unmont(spec, pipe.E1.cpu().numpy()), oE1), f"step {s}: folded E"


def test_supernova_running_instances(L, oracle, spec):
    """NIVC: two circuits of different shapes, steps [0, 1, 0]; each fold must land in its own running instance and
    leave the other untouched (src/proof/supernova.rs:207-291)."""
    import torch
    from lurk_beta_b200.fold import NovaFoldPipeline, SlotBatch, SuperNovaFoldPipeline
    rng = np.random.default_rng(7)
    p = spec.FIELD_MODULUS[FIELD]
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    shapes = [(4, 5, 700, 900), (8, 3, 1500, 1100)]       # (slot arity, slots, glue, constraints) per circuit
    n_key = 4096
    bases = oracle.gen_bases(CURVE, n_key)
    ck = L.CommitmentKey(CURVE, bases)
    circuits, state = [], []
    for k, (arity, nslots, glue, n_t) in enumerate(shapes):
        blk = oracle.witness_block(FIELD, arity)
        n_w = nslots * blk + glue
        ncols = n_w + 3
        nnz_per = rng.integers(1, 3, size=n_t)
        row_ptr = np.concatenate([[0], np.cumsum(nnz_per)]).astype(np.uint64)
        col = rng.integers(0, ncols, size=int(row_ptr[-1])).astype(np.uint32)
        val = pack([[1, p - 1, 3][i] for i in rng.integers(0, 3, size=col.size)])
        mats = [(row_ptr, col, val)] * 3
        d_mats = [(dev(row_ptr), dev(col), dev(mont(spec, val)))] * 3
        W1, E1, tail = random_elements(FIELD, n_w, 10 + k), random_elements(FIELD, n_t, 20 + k), random_elements(FIELD, 3, 30 + k)
        z1 = dev(mont(spec, np.concatenate([W1, tail])))
        z2 = [dev(mont(spec, np.concatenate([np.zeros(n_w * 32, dtype=np.uint8), tail]))) for _ in range(2)]
        u1, u2 = random_elements(FIELD, 1, 40 + k), pack([1])
        pipe = NovaFoldPipeline(torch, FIELD, CURVE, ck.clone(), n_w, n_t, d_mats, mont(spec, u1), mont(spec, u2), z1,
                                dev(mont(spec, E1)), z2)
        circuits.append(pipe)
        state.append(dict(arity=arity, nslots=nslots, blk=blk, glue=glue, n_w=n_w, n_t=n_t, mats=mats, W1=W1, E1=E1, tail=tail, u1=u1, u2=u2))
    nivc = SuperNovaFoldPipeline(circuits)

    def challenge(cw, ct):
        r = np.zeros(32, dtype=np.uint8)
        r[:16] = np.frombuffer(hashlib.sha256(cw.tobytes() + ct.tobytes()).digest()[:16], dtype=np.uint8)
        return r

    for step, k in enumerate([0, 1, 0]):
        st, pipe = state[k], circuits[k]
        pre = random_elements(FIELD, st["nslots"] * st["arity"], seed=200 + step, shape="lem")
        glue = random_elements(FIELD, st["glue"], seed=300 + step, shape="witness")
        d_pre = dev(mont(spec, pre))
        batches = [SlotBatch(st["arity"], st["nslots"], 0, d_pre)]

        def before(b, pipe=pipe, st=st, glue=glue):
            pipe.W2[b][st["nslots"] * st["blk"] * 32:].copy_(dev(mont(spec, glue)))
        b = nivc.stage_a(k, batches, before)
        other_before = unmont(spec, circuits[1 - k].W1.cpu().numpy())
        cw, ct = nivc.stage_b(k, b, challenge)
        torch.cuda.synchronize()
        oW2 = np.concatenate([oracle.poseidon_witness_batch(FIELD, st["arity"], pre), glue])
        oz1, oz2 = np.concatenate([st["W1"], st["tail"]]), np.concatenate([oW2, st["tail"]])
        az1, az2, bz1, bz2, cz1, cz2 = [oracle.spmv(FIELD, rp, col, val, z) for (rp, col, val) in st["mats"] for z in (oz1, oz2)]
        oT = oracle.cross_term(FIELD, az1, bz1, cz1, az2, bz2, cz2, st["u1"], st["u2"])
        pb = spec.FIELD_MODULUS[spec.CURVES[CURVE]["base"]]
        rinv = pow(R, -1, pb)
        assert [v * rinv % pb for v in ints(cw[:64])] == ints(oracle.msm(CURVE, bases, oW2, nthreads=4)[:64])
        assert [v * rinv % pb for v in ints(ct[:64])] == ints(oracle.msm(CURVE, bases, oT, nthreads=4)[:64])
        r_canon = pack([ints(challenge(cw, ct))[0] * pow(R, -1, p) % p])
        st["W1"] = oracle.axpy(FIELD, st["W1"], oW2, r_canon)
        st["E1"] = oracle.axpy(FIELD, st["E1"], oT, r_canon)
        assert np.array_equal(unmont(spec, pipe.W1.cpu().numpy()), st["W1"])
        assert np.array_equal(unmont(spec, pipe.E1.cpu().numpy()), st["E1"])
        assert np.array_equal(unmont(spec, circuits[1 - k].W1.cpu().numpy()), other_before)     # the other instance is untouched
