"""Host-side mirror of the reference's Poseidon interface (src/hash.rs), backed by the CUDA kernels.

`PoseidonCache` keeps the reference's method names and memoisation semantics (hash3/hash4/hash6/hash8,
compute_hash: src/hash.rs:97-113,180-203) and adds the batch entry point the GPU wants (`hash_batch`).
`HashConstants` exposes what `PoseidonConstants::new()` carries (src/hash.rs:41-84).
"""
import ctypes as C

import numpy as np

from . import _capi
from .field import pack, unpack

HASH_ARITIES = (3, 4, 6, 8)   # HashArity::{A3,A4,A6,A8}, src/hash.rs:11-29


class HashConstants:
    """Round numbers, round constants and MDS matrix of the Poseidon instance of one (field, arity)."""

    def __init__(self, field_id):
        self.field_id = field_id
        self._c = {}

    def constants(self, arity):
        if arity not in HASH_ARITIES:
            raise ValueError(f"unsupported arity: {arity}")   # reference panics the same way (src/hash.rs:26)
        if arity not in self._c:
            lib = _capi.lib()
            rf, rp = C.c_int(), C.c_int()
            _capi.check(lib.lurk_poseidon_constants(self.field_id, arity, C.byref(rf), C.byref(rp), None, None))
            t = arity + 1
            rc = np.zeros(t * (rf.value + rp.value) * 32, dtype=np.uint8)
            mds = np.zeros(t * t * 32, dtype=np.uint8)
            _capi.check(lib.lurk_poseidon_constants(self.field_id, arity, C.byref(rf), C.byref(rp), _capi.np_ptr(rc), _capi.np_ptr(mds)))
            m = unpack(mds)
            self._c[arity] = dict(full_rounds=rf.value, partial_rounds=rp.value, round_constants=unpack(rc),
                                  mds=[m[i * t:(i + 1) * t] for i in range(t)])
        return self._c[arity]

    def c3(self): return self.constants(3)
    def c4(self): return self.constants(4)
    def c6(self): return self.constants(6)
    def c8(self): return self.constants(8)


class PoseidonCache:
    """Memoised Poseidon digests of fixed-arity preimages of field elements (ints)."""

    def __init__(self, field_id=_capi.FIELD_BN254_FR):
        self.field_id = field_id
        self.constants = HashConstants(field_id)
        self._memo = {a: {} for a in HASH_ARITIES}

    # -- batch path (S1)
    def hash_batch_bytes(self, arity, preimages):
        """preimages: uint8 array of n*arity canonical elements -> uint8 array of n digests"""
        if arity not in HASH_ARITIES:
            raise ValueError(f"unsupported arity: {arity}")
        pre = np.ascontiguousarray(preimages, dtype=np.uint8).reshape(-1)
        if pre.size % (32 * arity):
            raise ValueError("preimage buffer is not a whole number of preimages")
        n = pre.size // (32 * arity)
        out = np.zeros(n * 32, dtype=np.uint8)
        _capi.check(_capi.lib().lurk_poseidon_hash_batch(self.field_id, arity, _capi.np_ptr(pre), n, _capi.np_ptr(out)))
        return out

    def hash_batch(self, arity, preimages):
        """preimages: list of arity-tuples of ints -> list of int digests (memoised)"""
        if arity not in HASH_ARITIES:
            raise ValueError(f"unsupported arity: {arity}")   # reference: panic!("unsupported arity"), src/hash.rs:26
        memo = self._memo[arity]
        keys, todo, seen = [], [], set()
        for p in preimages:
            key = tuple(int(x) for x in p)
            if len(key) != arity:
                raise ValueError(f"preimage of length {len(key)} for arity {arity}")
            keys.append(key)
            if key not in memo and key not in seen:
                seen.add(key)
                todo.append(key)
        if todo:
            digests = unpack(self.hash_batch_bytes(arity, pack([x for k in todo for x in k])))
            memo.update(zip(todo, digests))
        return [memo[k] for k in keys]

    # -- the reference's single-hash surface
    def compute_hash(self, preimage):
        return self.hash_batch(len(preimage), [preimage])[0]

    def hash3(self, preimage): return self.hash_batch(3, [preimage])[0]
    def hash4(self, preimage): return self.hash_batch(4, [preimage])[0]
    def hash6(self, preimage): return self.hash_batch(6, [preimage])[0]
    def hash8(self, preimage): return self.hash_batch(8, [preimage])[0]

    # -- impl StoreHasher for PoseidonCache (src/lem/store.rs:29-78); ptrs are (tag, digest) pairs
    def hash_ptrs(self, ptrs):
        if len(ptrs) not in (2, 3, 4):
            raise NotImplementedError("hash_ptrs takes 2, 3 or 4 pointers")   # unimplemented!() in the reference
        return self.compute_hash([x for tag, h in ptrs for x in (tag, h)])

    def hash_commitment(self, secret, payload):
        tag, h = payload
        return self.hash3([secret, tag, h])

    def hash_compact(self, d1, t2, d2, d3):
        return self.hash4([d1, t2, d2, d3])
