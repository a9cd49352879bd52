"""Slot witnesses: the batch body of generate_slots_witnesses (reference src/lem/multiframe.rs:520-592).

SlotType and preimage sizes follow src/lem/slot.rs:280-296; witness sizes follow compute_witness_size
(src/lem/multiframe.rs:503-516).  A frame's slots are laid out in the reference's order: hash4s, hash6s, hash8s,
commitments, bit_decomps (multiframe.rs:528-536); `None` slots are all-zero preimages (circuit.rs:301-313).
"""
import enum

import numpy as np

from . import _capi
from .field import pack


class SlotType(enum.Enum):
    Hash4 = "Hash4"
    Hash6 = "Hash6"
    Hash8 = "Hash8"
    Commitment = "Commitment"
    BitDecomp = "BitDecomp"

    def preimg_size(self):
        return {"Hash4": 4, "Hash6": 6, "Hash8": 8, "Commitment": 3, "BitDecomp": 1}[self.value]


def compute_witness_size(slot_type, field_id):
    """field elements per slot witness block (multiframe.rs:503-516)"""
    lib = _capi.lib()
    if slot_type is SlotType.BitDecomp:
        return lib.lurk_bitdecomp_witness_block(field_id)
    return lib.lurk_poseidon_witness_block(field_id, slot_type.preimg_size())


def slot_witness_batch_bytes(field_id, slot_type, preimages, fmt=_capi.FMT_CANONICAL):
    """preimages: uint8 array of n * preimg_size elements -> uint8 array of n witness blocks"""
    lib = _capi.lib()
    pre = np.ascontiguousarray(preimages, dtype=np.uint8).reshape(-1)
    a = slot_type.preimg_size()
    if pre.size % (32 * a):
        raise ValueError("preimage buffer is not a whole number of slot preimages")
    n = pre.size // (32 * a)
    out = np.zeros(n * compute_witness_size(slot_type, field_id) * 32, dtype=np.uint8)
    if slot_type is SlotType.BitDecomp:
        _capi.check(lib.lurk_bitdecomp_witness_batch(field_id, _capi.np_ptr(pre), n, _capi.np_ptr(out), fmt))
    else:
        _capi.check(lib.lurk_poseidon_witness_batch(field_id, a, _capi.np_ptr(pre), n, _capi.np_ptr(out), fmt))
    return out


def generate_slots_witnesses(field_id, slots):
    """slots: list of (SlotType, preimage ints or None) in frame order -> list of uint8 witness blocks in the same
    order.  One launch per slot type; dummy (None) slots share one cached witness per type like the reference
    (multiframe.rs:553-577)."""
    out = [None] * len(slots)
    by_type = {}
    for i, (st, pre) in enumerate(slots):
        if pre is not None and len(pre) != st.preimg_size():
            raise ValueError(f"slot {i}: {len(pre)} preimage elements for {st.value}")   # is_compatible, slot.rs:298-300
        by_type.setdefault(st, []).append(i)
    for st, idxs in by_type.items():
        a = st.preimg_size()
        live = [i for i in idxs if slots[i][1] is not None]
        rows = [[0] * a] + [list(slots[i][1]) for i in live]      # row 0 = the shared dummy witness
        blocks = slot_witness_batch_bytes(field_id, st, pack([x for r in rows for x in r]))
        size = compute_witness_size(st, field_id) * 32
        for k, i in enumerate(live):
            out[i] = blocks[(k + 1) * size:(k + 2) * size]
        for i in idxs:
            if slots[i][1] is None:
                out[i] = blocks[:size]
    return out
