"""Host-side mirror of the sparse Poseidon trie of the trie coprocessor (reference src/coprocessor/trie/mod.rs), backed by
the CUDA Poseidon kernels through `PoseidonCache`.

`StandardTrie = Trie<F, 8, 85>` (trie/mod.rs:43): arity 8, height 85 (3 x 85 = 255 key bits), empty element 0.  Same
method names as the reference: `path` (mod.rs:589-609), `empty_root` (485-497), `lookup` / `lookup_aux` (635-652),
`insert` (745-776), `prove_lookup_at_path` (725-743).  The evaluation side hashes through
`PoseidonCache::compute_hash` (hash.rs:97-113) and records preimages in an inverse cache, exactly like the reference;
`lookup_circuit_witnesses` produces the 85 arity-8 slot witnesses the lookup circuit allocates in
ONE batched launch -- the 85 hashes of a path are independent once the preimages are known."""
from .hash import PoseidonCache
from .slots import SlotType, slot_witness_batch_bytes
from .field import pack


class Trie:
    def __init__(self, poseidon_cache=None, arity=8, height=85, root=None, inverse_cache=None):
        if arity & (arity - 1):
            raise ValueError("ARITY must be a power of two")          # checked in new_aux()
        self.hash_cache = poseidon_cache or PoseidonCache()
        self.arity, self.height = arity, height
        self.children = inverse_cache if inverse_cache is not None else {}   # digest -> preimage (InversePoseidonCache)
        self.empty_roots = []
        cur = self.empty_element()
        for _ in range(height):
            cur = self.register_hash([cur] * arity)
            self.empty_roots.append(cur)
        self.root = self.empty_roots[height - 1] if root is None else int(root)

    @staticmethod
    def empty_element():
        return 0

    def register_hash(self, preimage):
        d = self.hash_cache.compute_hash(list(preimage))
        self.children[d] = tuple(int(x) for x in preimage)
        return d

    def empty_root_for_height(self, height):
        return self.empty_element() if height == 0 else self.empty_roots[height - 1]

    def empty_root(self):
        return self.empty_root_for_height(self.height)

    def arity_bits(self):
        return self.arity.bit_length() - 1

    def path(self, key):
        """most-significant chunk first, `arity_bits` bits per level (mod.rs:589-609)"""
        ab, n = self.arity_bits(), self.height
        key = int(key)
        return [(key >> (ab * (n - 1 - i))) & (self.arity - 1) for i in range(n)]

    def prove_lookup_at_path(self, path):
        preimages, nxt = [], self.root
        for k in path:
            if nxt not in self.children:
                raise KeyError(f"MissingPreimage({hex(nxt)})")
            pre = self.children[nxt]
            preimages.append(pre)
            nxt = pre[k]
        return preimages

    def lookup_aux(self, key):
        path = self.path(key)
        return self.prove_lookup_at_path(path)[-1][path[-1]]

    def lookup(self, key):
        v = self.lookup_aux(key)
        return None if v == self.empty_element() else v

    def insert(self, key, value):
        path = self.path(key)
        old = self.prove_lookup_at_path(path)
        value = int(value)
        for k, existing in zip(reversed(path), reversed(old)):
            new_pre = list(existing)
            new_pre[k] = value
            value = self.register_hash(new_pre)
        inserted = value != self.root
        self.root = value
        return inserted

    def lookup_circuit_witnesses(self, key, fmt=0):
        """the HEIGHT arity-8 Poseidon witnesses of synthesize_lookup (mod.rs:654-724) as one batch: uint8 array of
        HEIGHT slot blocks, root level first"""
        preimages = self.prove_lookup_at_path(self.path(key))
        st = {4: SlotType.Hash4, 8: SlotType.Hash8}[self.arity]
        return slot_witness_batch_bytes(self.hash_cache.field_id, st, pack([x for p in preimages for x in p]), fmt)


def StandardTrie(poseidon_cache=None, root=None, inverse_cache=None):
    return Trie(poseidon_cache, 8, 85, root, inverse_cache)
