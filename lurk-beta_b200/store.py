"""Host-side mirror of the reference's hash-consed store core (src/lem/store_core.rs) for the hydration path.

Same surface, trimmed to what the hot path needs: intern_atom / intern_tuple2/3/4 / intern_compact, the
`dehydrated` queue, `hydrate_z_cache` (one GPU launch sequence instead of rayon over chunks of 256,
store_core.rs:256-269), `hash_ptr_val` / `hash_ptr`, commitments (`hide`, src/lem/store.rs:70-73).
Pointers are (tag: int u16, val) with val = ("atom"|"tuple2"|"tuple3"|"tuple4"|"compact", index).
"""
import numpy as np

from . import _capi
from .field import pack, unpack
from .hash import PoseidonCache

_KIND = {"tuple2": 2, "tuple3": 3, "tuple4": 4, "compact": 5}
# struct lurk_dag_node (include/lurk_b200.h), C layout: 28 bytes
DAG_NODE = np.dtype([("kind", "u1"), ("reserved", "u1"), ("tag", "<u2", (4,)), ("child", "<u4", (4,))], align=True)
assert DAG_NODE.itemsize == 28


class StoreCore:
    def __init__(self, field_id=_capi.FIELD_BN254_FR):
        self.field_id = field_id
        self.hasher = PoseidonCache(field_id)
        self._atoms, self._atom_ix = [], {}
        self._tuples = {k: ([], {}) for k in _KIND}
        self.dehydrated = []          # compound vals not hashed yet, in interning order (children first)
        self.z_cache = {}             # val -> digest
        self.inverse_z_cache = {}     # digest -> val
        self.comms = {}               # digest -> (secret, payload ptr)

    # -- interning (store_core.rs:108-172): index interning only, no hashing
    def intern_atom(self, tag, digest):
        d = int(digest)
        if d not in self._atom_ix:
            self._atom_ix[d] = len(self._atoms)
            self._atoms.append(d)
        return (tag, ("atom", self._atom_ix[d]))

    def _intern(self, kind, ptrs, tag):
        rows, index = self._tuples[kind]
        key = tuple(ptrs)
        if key not in index:
            index[key] = len(rows)
            rows.append(key)
            self.dehydrated.append((kind, index[key]))
        return (tag, (kind, index[key]))

    def intern_tuple2(self, ptrs, tag): assert len(ptrs) == 2; return self._intern("tuple2", ptrs, tag)
    def intern_tuple3(self, ptrs, tag): assert len(ptrs) == 3; return self._intern("tuple3", ptrs, tag)
    def intern_tuple4(self, ptrs, tag): assert len(ptrs) == 4; return self._intern("tuple4", ptrs, tag)
    def intern_compact(self, ptrs, tag): assert len(ptrs) == 3; return self._intern("compact", ptrs, tag)

    def fetch_digest(self, idx): return self._atoms[idx]
    def expect_children(self, val): return self._tuples[val[0]][0][val[1]]

    # -- hydration (S2)
    def _collect(self, roots):
        """post-order list of compound vals reachable from `roots` that are not in z_cache (iterative, like the
        safe variant store_core.rs:287-330)"""
        order, seen, stack = [], set(), [(v, False) for v in roots]
        while stack:
            v, done = stack.pop()
            if v[0] == "atom" or v in self.z_cache:
                continue
            if done:
                order.append(v)
                continue
            if v in seen:
                continue
            seen.add(v)
            stack.append((v, True))
            for (_t, c) in self.expect_children(v):
                stack.append((c, False))
        return order

    def _build_nodes(self, vals):
        """node list + digest table handed to the library: [atoms | digests of already-hashed compound children | vals]"""
        n_atoms = len(self._atoms)
        slot = {v: i for i, v in enumerate(vals)}
        extra, extra_ix = [], {}
        for v in vals:
            for (_t, c) in self.expect_children(v):
                if c[0] != "atom" and c not in slot:
                    d = self.z_cache[c]
                    if d not in extra_ix:
                        extra_ix[d] = len(extra)
                        extra.append(d)
        base = n_atoms + len(extra)
        nodes = np.zeros(len(vals), dtype=DAG_NODE)
        for i, v in enumerate(vals):
            nodes[i]["kind"] = _KIND[v[0]]
            for j, (t, c) in enumerate(self.expect_children(v)):
                nodes[i]["tag"][j] = t
                if c[0] == "atom":
                    nodes[i]["child"][j] = c[1]
                elif c in slot:
                    nodes[i]["child"][j] = base + slot[c]
                else:
                    nodes[i]["child"][j] = n_atoms + extra_ix[self.z_cache[c]]
        return nodes, pack(self._atoms + extra), base

    def _hash_vals(self, vals):
        if not vals:
            return
        nodes, atoms, base = self._build_nodes(vals)
        out = np.zeros(len(vals) * 32, dtype=np.uint8)
        _capi.check(_capi.lib().lurk_dag_hash(self.field_id, _capi.np_ptr(nodes), len(vals), _capi.np_ptr(atoms),
                                              base, _capi.np_ptr(out)))
        for v, d in zip(vals, unpack(out)):
            self.z_cache[v] = d
            self.inverse_z_cache[d] = v

    def hydration_plan(self):
        """lurk_dag_hash_plan for the queued nodes (host only).  The caller -- the reference's StoreCore -- keeps deep,
        narrow hydrations on its own CPU path (hash_ptr_val_unsafe, store_core.rs:199-248) when `use_gpu` is 0."""
        import ctypes as C
        todo = [v for v in self.dehydrated if v not in self.z_cache]
        plan = _capi.DagPlan()
        if todo:
            nodes, _atoms, base = self._build_nodes(todo)
            _capi.check(_capi.lib().lurk_dag_hash_plan(_capi.np_ptr(nodes), len(todo), base, C.byref(plan)))
        return plan

    def hydrate_z_cache(self):
        """hash every queued compound node (store_core.rs:266-269)"""
        todo = [v for v in self.dehydrated if v not in self.z_cache]
        self.dehydrated = []
        self._hash_vals(todo)   # interning order is already children-first

    def hash_ptr_val(self, val):
        if val[0] == "atom":
            return self._atoms[val[1]]
        if val not in self.z_cache:
            self._hash_vals(self._collect([val]))
        return self.z_cache[val]

    def hash_ptr(self, ptr):
        """Ptr -> ZPtr = (tag, digest) (store_core.rs:333-335)"""
        return (ptr[0], self.hash_ptr_val(ptr[1]))

    # -- commitments (src/lem/store.rs:70-73, hide / open)
    def hide(self, secret, payload_ptr):
        z = self.hash_ptr(payload_ptr)
        digest = self.hasher.hash_commitment(int(secret), z)
        self.comms[digest] = (int(secret), payload_ptr)
        return digest

    def open(self, digest):
        return self.comms.get(int(digest))
