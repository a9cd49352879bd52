"""GPU parity of the fold context (lurk_fold_ctx_*, csrc/foldctx_impl.cuh) -- the GPU half of prove_recursively
(src/proof/nova.rs:260-339, supernova.rs:207-291) -- against the oracle's Nova folding (oracle/nifs.py), step by step:
slot witnesses written in place into W2, commit(W2), cross term, commit(T), the random-oracle challenge, the fold of
(W, u, X), E and of the commitments; then the protocol-level property: the folded instance read back from the device
satisfies the relaxed R1CS on the oracle, with consistent commitments.  Everything goes through the C ABI."""
import numpy as np
import pytest

from util import random_elements

pytestmark = pytest.mark.gpu
CURVE = 0            # BN254 G1: witness field Fr (0), commitment coordinates / RO field Fq (1)
FIELD = 0
R = 1 << 256


def _layout(oracle, frames, glue, slots_per_frame=((4, 14), (8, 6), (3, 1)), bd_per_frame=3):
    """the reference's frame layout (src/lem/multiframe.rs:635-712): per frame [slot blocks in slot order | body aux]"""
    blocks = {a: oracle.witness_block(FIELD, a) for a, _ in slots_per_frame}
    bd_block = oracle.bitdecomp_size(FIELD)
    slot_elems = sum(n * blocks[a] for a, n in slots_per_frame) + bd_per_frame * bd_block
    per = slot_elems + glue
    offs, cur = {}, 0
    for a, n in slots_per_frame:
        offs[a] = np.array([f * per + cur + k * blocks[a] for f in range(frames) for k in range(n)], dtype=np.uint64)
        cur += n * blocks[a]
    offs[0] = np.array([f * per + cur + k * bd_block for f in range(frames) for k in range(bd_per_frame)], dtype=np.uint64)
    return dict(blocks=blocks, bd_block=bd_block, slot_elems=slot_elems, per=per, offs=offs, frames=frames, glue=glue,
                slots=[(a, n * frames) for a, n in slots_per_frame], nbd=bd_per_frame * frames)


def _step_inputs(oracle, nifs, spec, lay, glue_fn, seed, rng):
    """slot preimages of one step + the fresh witness the reference would assemble from them (oracle side)"""
    p = spec.FIELD_MODULUS[FIELD]
    pre = {}
    for a, n in lay["slots"]:
        x = random_elements(FIELD, n * a, seed=100 * seed + a, shape="lem").reshape(n, a * 32)
        x[rng.random(n) < 0.6] = 0                       # most slots are dummies (multiframe.rs:553-577)
        pre[a] = x.reshape(-1)
    bd = random_elements(FIELD, lay["nbd"], seed=50 + seed, shape="witness")
    n_w = lay["frames"] * lay["per"]
    W = np.zeros(n_w * 32, dtype=np.uint8)
    Wv = W.reshape(n_w, 32)
    for a, n in lay["slots"]:
        blk = lay["blocks"][a]
        wit = oracle.poseidon_witness_batch(FIELD, a, pre[a], nthreads=4).reshape(n, blk, 32)
        for k, off in enumerate(lay["offs"][a]):
            Wv[int(off):int(off) + blk] = wit[k]
    wit = oracle.bitdecomp_witness_batch(FIELD, bd).reshape(lay["nbd"], lay["bd_block"], 32)
    for k, off in enumerate(lay["offs"][0]):
        Wv[int(off):int(off) + lay["bd_block"]] = wit[k]
    Wi = nifs.ints(W)
    glue_vals = glue_fn(Wi, p)
    for dst, v in glue_vals.items():
        Wv[dst] = np.frombuffer(int(v).to_bytes(32, "little"), dtype=np.uint8)
    glue_dense = nifs.pack([glue_vals[f * lay["per"] + lay["slot_elems"] + g] for f in range(lay["frames"]) for g in range(lay["glue"])])
    X2 = [int(rng.integers(1, 2**62)) * int(rng.integers(1, 2**62)) for _ in range(2)]
    return dict(pre=pre, bd=bd, W2=W, glue=glue_dense, X2=X2)


def _fill(ctx, b, lay, st, pp_digest, batch_index, mont=None):
    """what the CPU witness generator does: write this step's inputs into the context's pinned buffers"""
    conv = (lambda x: x) if mont is None else mont
    for a, _ in lay["slots"]:
        ctx.host_buffer(b, batch_index[a])[:] = conv(st["pre"][a])
    ctx.host_buffer(b, batch_index[0])[:] = conv(st["bd"])
    ctx.host_buffer(b, -1)[:] = conv(st["glue"])
    ctx.host_buffer(b, -2)[:] = conv(np.frombuffer(b"".join(int(x).to_bytes(32, "little") for x in st["X2"]), dtype=np.uint8))
    ro = np.zeros(24 * 32, dtype=np.uint8).reshape(24, 32)
    for pos, v in ((0, pp_digest), (4, st["X2"][0]), (5, st["X2"][1])):
        ro[pos] = np.frombuffer(int(v).to_bytes(32, "little"), dtype=np.uint8)
    ctx.host_buffer(b, -3)[:] = ro.reshape(-1) if mont is None else mont(ro.reshape(-1), base=True)


def _build(L, oracle, nifs, spec, rng, frames, glue, lin_rows, depth=2, fmt=None, ck=None, bases=None, **kw):
    lay = _layout(oracle, frames, glue)
    mats, n_w, glue_fn = nifs.synthetic_step_circuit(rng, frames, lay["slot_elems"], glue, lin_rows)
    rows = len(mats[0][0]) - 1
    if bases is None:
        bases = oracle.gen_bases(CURVE, max(n_w, rows))
        ck = L.CommitmentKey(CURVE, bases)
    ctx = L.NovaFoldContext(CURVE, ck, n_w, 2, mats, depth=depth, fmt=L.FMT_CANONICAL, **kw)
    batch_index = {a: ctx.add_slot_batch(a, lay["offs"][a]) for a, _ in lay["slots"]}
    batch_index[0] = ctx.add_slot_batch(0, lay["offs"][0])
    ctx.set_spans([(lay["slot_elems"], glue, lay["per"], frames)])
    return ctx, lay, mats, n_w, rows, glue_fn, bases, ck, batch_index


def _same_point(nifs, buf96, P):
    return np.array_equal(buf96, nifs.point_bytes(P))


def test_ivc_chain_matches_oracle_and_stays_satisfiable(L, oracle, spec):
    from oracle import nifs
    rng = np.random.default_rng(42)
    ctx, lay, mats, n_w, rows, glue_fn, bases, ck, bi = _build(L, oracle, nifs, spec, rng, frames=2, glue=40, lin_rows=60)
    pp_digest = 0x1234567890abcdef1122334455667788
    o = nifs.NovaOracle(CURVE, bases, mats, n_w, 2, nthreads=8, pp_digest=pp_digest)
    steps = [_step_inputs(oracle, nifs, spec, lay, glue_fn, s, rng) for s in range(4)]
    # step 0: RecursiveSNARK::new; stage A of step 1 is enqueued before step 0 is collected (one step ahead)
    _fill(ctx, 0, lay, steps[0], pp_digest, bi)
    ctx.stage_a(0)
    ctx.init_running(0)
    _fill(ctx, 1, lay, steps[1], pp_digest, bi)
    ctx.stage_a(1)
    rec = ctx.collect(0)
    assert np.array_equal(ctx.read_device(0, -4)[:n_w * 32], _mont_bytes(spec, steps[0]["W2"])), "W2 of step 0"
    want = o.init_running(steps[0]["W2"], steps[0]["X2"])
    assert _same_point(nifs, rec.comm_W, want["comm_W"]) and _same_point(nifs, rec.running_comm_W, want["comm_W"])
    assert not rec.running_comm_E.any() and not rec.comm_T.any()
    for s in range(1, 4):
        b = s & 1
        ctx.stage_b_launch(b)
        rec = ctx.collect(b)
        if s + 1 < 4:                                   # the other buffer is free again: prefetch the next step
            _fill(ctx, b ^ 1, lay, steps[s + 1], pp_digest, bi)
            ctx.stage_a(b ^ 1)
        want = o.prove_step(steps[s]["W2"], steps[s]["X2"])
        assert _same_point(nifs, rec.comm_W, want["comm_W"]), f"step {s}: comm_W"
        assert _same_point(nifs, rec.comm_T, want["comm_T"]), f"step {s}: comm_T"
        assert int.from_bytes(rec.ro_hash.tobytes(), "little") == want["hash"], f"step {s}: sponge output"
        assert int.from_bytes(rec.r.tobytes(), "little") == want["r"], f"step {s}: challenge"
        assert _same_point(nifs, rec.running_comm_W, o.comm_W) and _same_point(nifs, rec.running_comm_E, o.comm_E), f"step {s}: folded commitments"
    run = ctx.get_running()
    assert np.array_equal(run["W"], o.W) and np.array_equal(run["E"], o.E)
    assert nifs.ints(run["u"]) == [o.u] and nifs.ints(run["X"]) == o.X
    assert _same_point(nifs, run["comm_W"], o.comm_W) and _same_point(nifs, run["comm_E"], o.comm_E)
    # the verifier's view, from what the device holds: relaxed R1CS satisfied, commitments open to the vectors
    assert o.bad_rows(run["W"], run["E"], nifs.ints(run["u"])[0], nifs.ints(run["X"])) == 0
    assert o.commitments_consistent(run["W"], run["E"], nifs.point_of(run["comm_W"]), nifs.point_of(run["comm_E"])) == (True, True)
    assert ctx.check_running() == (0, True, True)
    st = ctx.stats()
    assert st["launches_a"] > 10 and st["launches_b"] > 10


def _mont_bytes(spec, buf, field=FIELD):
    from oracle import nifs
    p = spec.FIELD_MODULUS[field]
    return nifs.pack([x * R % p for x in nifs.ints(buf)])


def test_checkpoint_resume_and_montgomery_inputs(L, oracle, spec):
    """prove_recursively(.., init: Some(snark)) (src/proof/mod.rs:107-115): the running instance read back from one
    context and installed into a fresh one continues to the same result; the second context takes Montgomery inputs"""
    from oracle import nifs
    rng = np.random.default_rng(7)
    ctx, lay, mats, n_w, rows, glue_fn, bases, ck, bi = _build(L, oracle, nifs, spec, rng, frames=1, glue=24, lin_rows=30)
    pp = 99
    steps = [_step_inputs(oracle, nifs, spec, lay, glue_fn, 10 + s, rng) for s in range(3)]
    _fill(ctx, 0, lay, steps[0], pp, bi); ctx.stage_a(0); ctx.init_running(0); ctx.collect(0)
    _fill(ctx, 1, lay, steps[1], pp, bi); ctx.stage_a(1); ctx.stage_b_launch(1); ctx.collect(1)
    snap = ctx.get_running()
    _fill(ctx, 0, lay, steps[2], pp, bi); ctx.stage_a(0); ctx.stage_b_launch(0); last = ctx.collect(0)
    final = ctx.get_running()

    ctx2 = L.NovaFoldContext(CURVE, ck, n_w, 2, mats, depth=1, fmt=L.FMT_CANONICAL)
    bi2 = {a: ctx2.add_slot_batch(a, lay["offs"][a]) for a, _ in lay["slots"]}
    bi2[0] = ctx2.add_slot_batch(0, lay["offs"][0])
    ctx2.set_spans([(lay["slot_elems"], lay["glue"], lay["per"], lay["frames"])])
    ctx2.set_running(snap["W"], snap["E"], snap["u"], snap["X"], snap["comm_W"], snap["comm_E"])

    def mont(x, base=False):
        return _mont_bytes(spec, x, field=1 if base else 0)
    _fill(ctx2, 0, lay, steps[2], pp, bi2, mont=mont)
    ctx2.stage_a(0, fmt=L.FMT_MONTGOMERY)
    ctx2.stage_b_launch(0)
    with pytest.raises(L.LurkError):      # misuse is reported, not executed: the buffer's result has not been collected
        ctx2.stage_a(0, fmt=L.FMT_MONTGOMERY)
    rec = ctx2.collect(0)
    assert np.array_equal(rec.r, last.r) and np.array_equal(rec.comm_T, last.comm_T)
    again = ctx2.get_running()
    for k in ("W", "E", "u", "X", "comm_W", "comm_E"):
        assert np.array_equal(again[k], final[k]), k
    assert ctx2.check_running() == (0, True, True)


def test_nivc_two_circuits_share_one_key(L, oracle, spec):
    """SuperNova (src/proof/supernova.rs:207-291): each circuit index folds into its own running instance; the Lurk step
    circuit and a coprocessor circuit of another shape share the device-resident commitment key"""
    from oracle import nifs
    rng = np.random.default_rng(3)
    ctx0, lay0, mats0, n_w0, rows0, glue0, bases, ck, bi0 = _build(L, oracle, nifs, spec, rng, frames=1, glue=30, lin_rows=25)
    # coprocessor circuit (trie lookup shape, src/coprocessor/trie/mod.rs:592-640): only arity-8 slots + glue
    blk8 = oracle.witness_block(FIELD, 8)
    nslots, glue1 = 5, 17
    slot_elems1 = nslots * blk8
    mats1, n_w1, glue_fn1 = nifs.synthetic_step_circuit(rng, 1, slot_elems1, glue1, 20)
    assert max(n_w1, len(mats1[0][0]) - 1) <= len(bases) // 64
    ctx1 = L.NovaFoldContext(CURVE, ck, n_w1, 2, mats1, depth=2, fmt=L.FMT_CANONICAL)
    offs8 = np.arange(nslots, dtype=np.uint64) * blk8
    b8 = ctx1.add_slot_batch(8, offs8)
    ctx1.set_spans([(slot_elems1, glue1, slot_elems1 + glue1, 1)])
    nivc = L.SuperNovaFoldContext([ctx0, ctx1])
    o = [nifs.NovaOracle(CURVE, bases, mats0, n_w0, 2, nthreads=8, pp_digest=5), nifs.NovaOracle(CURVE, bases, mats1, n_w1, 2, nthreads=8, pp_digest=5)]
    started = [False, False]
    p = spec.FIELD_MODULUS[FIELD]
    for s, ci in enumerate([0, 1, 0, 1, 1, 0]):
        if ci == 0:
            st = _step_inputs(oracle, nifs, spec, lay0, glue0, 30 + s, rng)
            b = nivc._next[0]
            _fill(ctx0, b, lay0, st, 5, bi0)
        else:
            pre = random_elements(FIELD, nslots * 8, seed=70 + s, shape="lem")
            wit = oracle.poseidon_witness_batch(FIELD, 8, pre, nthreads=4)
            W = np.zeros(n_w1 * 32, dtype=np.uint8)
            W[:slot_elems1 * 32] = wit
            gv = glue_fn1(nifs.ints(W), p)
            glue_dense = nifs.pack([gv[slot_elems1 + g] for g in range(glue1)])
            W[slot_elems1 * 32:] = glue_dense
            st = dict(W2=W, X2=[int(rng.integers(1, 2**61)) for _ in range(2)])
            b = nivc._next[1]
            ctx1.host_buffer(b, b8)[:] = pre
            ctx1.host_buffer(b, -1)[:] = glue_dense
            ctx1.host_buffer(b, -2)[:] = nifs.pack(st["X2"])
            ro = np.zeros((24, 32), dtype=np.uint8)
            for pos, v in ((0, 5), (4, st["X2"][0]), (5, st["X2"][1])):
                ro[pos] = np.frombuffer(int(v).to_bytes(32, "little"), dtype=np.uint8)
            ctx1.host_buffer(b, -3)[:] = ro.reshape(-1)
        assert nivc.stage_a(ci) == b
        nivc.fold(ci, b)
        rec = nivc.collect(ci, b)
        if not started[ci]:
            want = o[ci].init_running(st["W2"], st["X2"])
            started[ci] = True
        else:
            want = o[ci].prove_step(st["W2"], st["X2"])
            assert int.from_bytes(rec.r.tobytes(), "little") == want["r"], f"step {s}"
            assert _same_point(nifs, rec.comm_T, want["comm_T"]), f"step {s}"
        assert _same_point(nifs, rec.comm_W, want["comm_W"]), f"step {s}"
    for ci, c in enumerate((ctx0, ctx1)):
        run = c.get_running()
        assert np.array_equal(run["W"], o[ci].W) and np.array_equal(run["E"], o[ci].E)
        assert o[ci].bad_rows(run["W"], run["E"], nifs.ints(run["u"])[0], nifs.ints(run["X"])) == 0
        assert c.check_running() == (0, True, True)


def test_secondary_curve_instance_grumpkin(L, oracle, spec):
    """the secondary circuit of the cycle (Grumpkin, witness field = BN254 Fq) folds through the same context type"""
    from oracle import nifs
    rng = np.random.default_rng(9)
    curve = 1
    field = spec.CURVES[curve]["scalar"]
    p = spec.FIELD_MODULUS[field]
    mats, n_w, glue_fn = nifs.synthetic_step_circuit(rng, 1, 300, 40, 80)
    rows = len(mats[0][0]) - 1
    bases = oracle.gen_bases(curve, max(n_w, rows))
    ck = L.CommitmentKey(curve, bases)
    ctx = L.NovaFoldContext(curve, ck, n_w, 2, mats, depth=1, fmt=L.FMT_CANONICAL)
    ctx.set_spans([(0, n_w, n_w, 1)])                      # no slots: the whole witness comes from the host
    o = nifs.NovaOracle(curve, bases, mats, n_w, 2, nthreads=4, pp_digest=77)
    for s in range(3):
        W = [int(rng.integers(0, 2**62)) * int(rng.integers(0, 2**62)) % p for _ in range(n_w)]
        for dst, v in glue_fn(W, p).items():
            W[dst] = v
        Wb, X2 = nifs.pack(W), [int(rng.integers(1, 2**60)) for _ in range(2)]
        ctx.host_buffer(0, -1)[:] = Wb
        ctx.host_buffer(0, -2)[:] = nifs.pack(X2)
        ro = np.zeros((24, 32), dtype=np.uint8)
        for pos, v in ((0, 77), (4, X2[0]), (5, X2[1])):
            ro[pos] = np.frombuffer(int(v).to_bytes(32, "little"), dtype=np.uint8)
        ctx.host_buffer(0, -3)[:] = ro.reshape(-1)
        ctx.stage_a(0)
        if s == 0:
            ctx.init_running(0)
            want = o.init_running(Wb, X2)
        else:
            ctx.stage_b_launch(0)
            want = o.prove_step(Wb, X2)
        rec = ctx.collect(0)
        assert _same_point(nifs, rec.comm_W, want["comm_W"])
        if s:
            assert int.from_bytes(rec.r.tobytes(), "little") == want["r"]
    run = ctx.get_running()
    assert o.bad_rows(run["W"], run["E"], nifs.ints(run["u"])[0], nifs.ints(run["X"])) == 0
    assert ctx.check_running() == (0, True, True)
