"""Reference-side traces: every `slots_*.bin` / `commit_*.bin` under tests/golden/traces/ is replayed
through the oracle (CPU) and through the CUDA library (-m gpu) and must match byte for byte.  Files written by a lurk-beta
built with integration/rust/trace_export.patch pin the Neptune / bellpepper aux ORDER and Arecibo's commitment bytes --
today "parity unpinned" (DESIGN.md section 2).  Until such a file is dropped in, the committed synthetic traces (header
flag bit 0; written by tools/make_synthetic_trace.py from the ORACLE) keep the format, the loader and both replay paths
exercised; they pin nothing about the reference and the tests say so in their ids."""
import glob
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TRACES = os.path.join(ROOT, "tests", "golden", "traces")


def _files(pattern):
    return sorted(glob.glob(os.path.join(TRACES, pattern)))


def _id(path):
    import lurk_beta_b200.trace as T
    b = os.path.basename(path)
    t = T.read_slots(path) if b.startswith("slots") else T.read_key(path) if b.startswith("ck_") else T.read_commit(path)
    return os.path.basename(path) + (" [synthetic]" if t.synthetic else " [reference]")


def test_writer_reader_round_trip(tmp_path):
    import lurk_beta_b200.trace as T
    rng = np.random.default_rng(0)
    slots = [T.Slot("Hash4", False, rng.integers(0, 256, 293 * 32, dtype=np.uint8)), T.Slot("BitDecomp", True, np.zeros(354 * 32, dtype=np.uint8))]
    p = str(tmp_path / "slots_00000.bin")
    T.write_slots(p, 0, slots)
    back = T.read_slots(p)
    assert back.field_id == 0 and back.synthetic and len(back.slots) == 2
    assert all(a.slot_type == b.slot_type and a.is_dummy == b.is_dummy and np.array_equal(a.witness, b.witness) for a, b in zip(slots, back.slots))
    with open(p, "ab") as f:
        f.write(b"x")
    with pytest.raises(ValueError):
        T.read_slots(p)


@pytest.mark.parametrize("path", _files("slots_*.bin"), ids=_id)
def test_neptune_witness_trace_on_the_oracle(path, oracle):
    import lurk_beta_b200.trace as T
    tr = T.read_slots(path)
    for typ, (pre, wit, _idx) in T.slot_batches(tr).items():
        a = T.SLOT_ARITY[typ]
        got = oracle.poseidon_witness_batch(tr.field_id, a, pre) if a else oracle.bitdecomp_witness_batch(tr.field_id, pre)
        assert np.array_equal(got, wit), f"{typ}: oracle aux differs from the trace"


@pytest.mark.gpu
@pytest.mark.parametrize("path", _files("slots_*.bin"), ids=_id)
def test_neptune_witness_trace_on_the_gpu(path, L):
    import lurk_beta_b200.trace as T
    tr = T.read_slots(path)
    for typ, (pre, wit, _idx) in T.slot_batches(tr).items():
        got = L.slot_witness_batch_bytes(tr.field_id, getattr(L.SlotType, typ), pre)
        assert np.array_equal(got, wit), f"{typ}: CUDA aux differs from the trace"


@pytest.mark.parametrize("path", _files("commit_*.bin"), ids=_id)
def test_arecibo_commit_trace_on_the_oracle(path, oracle):
    import lurk_beta_b200.trace as T
    tr = T.read_commit(path)
    got = oracle.msm(tr.curve_id, tr.bases, tr.scalars)
    assert bool(got[64:].any()) != tr.is_identity
    assert np.array_equal(got[:64], tr.result)


@pytest.mark.gpu
@pytest.mark.parametrize("path", _files("commit_*.bin"), ids=_id)
def test_arecibo_commit_trace_on_the_gpu(path, L):
    import lurk_beta_b200.trace as T
    tr = T.read_commit(path)
    got = L.CommitmentKey(tr.curve_id, tr.bases).commit(tr.scalars)
    assert bool(got[64:].any()) != tr.is_identity
    assert np.array_equal(got[:64], tr.result)


def _key_points_from_oracle(tr):
    from oracle import h2c, kzg
    if tr.kind == 0:
        return h2c.from_label_bytes(tr.curve_id, tr.label, tr.points.size // 64)
    g = (int.from_bytes(tr.label[:32], "little"), int.from_bytes(tr.label[32:64], "little"))
    beta = int.from_bytes(tr.label[64:96], "little")
    pts = kzg.powers_of_tau(tr.curve_id, g, beta, tr.points.size // 64)
    return b"".join((0).to_bytes(64, "little") if P is None else P[0].to_bytes(32, "little") + P[1].to_bytes(32, "little") for P in pts)


@pytest.mark.parametrize("path", _files("ck_*.bin"), ids=_id)
def test_commitment_key_trace_on_the_oracle(path):
    """the head of the reference's commitment key: a reference-written file pins from_label / powers of tau"""
    import lurk_beta_b200.trace as T
    tr = T.read_key(path)
    assert _key_points_from_oracle(tr) == tr.points.tobytes()


@pytest.mark.gpu
@pytest.mark.parametrize("path", _files("ck_*.bin"), ids=_id)
def test_commitment_key_trace_on_the_gpu(path, L):
    import lurk_beta_b200.trace as T
    tr = T.read_key(path)
    n = tr.points.size // 64
    if tr.kind == 0:
        got = L.from_label(tr.curve_id, tr.label, n).tobytes()
    else:
        import ctypes as C
        import torch
        buf = torch.empty(n * 64, dtype=torch.uint8, device="cuda")
        lab = np.frombuffer(tr.label, dtype=np.uint8).copy()
        base_field = {0: 1, 1: 0, 2: 3, 3: 2}[tr.curve_id]
        L._capi.check(L._capi.lib().lurk_ck_powers_dev(tr.curve_id, L._capi.np_ptr(lab[:64]), L._capi.np_ptr(lab[64:96]), n, C.c_void_p(buf.data_ptr()),
                                                       L.FMT_CANONICAL, None))
        L._capi.check(L._capi.lib().lurk_convert_dev(base_field, C.c_void_p(buf.data_ptr()), 2 * n, L.FMT_CANONICAL, C.c_void_p(buf.data_ptr()), None))
        torch.cuda.synchronize()
        got = buf.cpu().numpy().tobytes()
    assert got == tr.points.tobytes()


def test_key_trace_round_trip(tmp_path):
    import lurk_beta_b200.trace as T
    p = str(tmp_path / "ck_x.bin")
    pts = np.arange(128, dtype=np.uint8)
    T.write_key(p, 2, 0, b"label", pts)
    back = T.read_key(p)
    assert (back.curve_id, back.kind, back.label, back.synthetic) == (2, 0, b"label", True) and np.array_equal(back.points, pts)
    with open(p, "ab") as f:
        f.write(b"x")
    with pytest.raises(ValueError):
        T.read_key(p)
