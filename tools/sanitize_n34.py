#!/usr/bin/env python3
"""Small invocations of every N3 / N4 entry point for compute-sanitizer (memcheck / racecheck): sizes chosen so that ragged tails,
multi-level sweeps and multi-CTA reductions are all exercised but a 50x slowdown stays within a minute."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lurk_beta_b200 as L

lib = L._capi.lib()
chk = L._capi.check


def rand_mont(n, seed):
    rng = np.random.default_rng(seed)
    raw = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    raw[:, 31] &= 0x1f
    return torch.from_numpy(raw.reshape(-1)).cuda()


def chal(rnd, msg):
    return 3 + rnd + (msg[0] if msg else 0)


for curve in range(4):
    L.from_label(curve, b"ck", 300)
    L.hash_to_curve_batch(curve, "x", np.arange(33 * 7, dtype=np.uint8), 7)
ck = L.CommitmentKey.setup(0, b"ck", (1 << 12) + 3)
g = L.synthetic_bases(0, 1, start=4)
gi = [int.from_bytes(g[:32].tobytes(), "little"), int.from_bytes(g[32:].tobytes(), "little")]
kz = L.CommitmentKey.powers_of_tau(0, gi, 123456789, (1 << 11) + 1)
for l in (0, 1, 5, 11):
    for kind, k in ((L.spartan.QUAD, 2), (L.spartan.CUBIC, 4)):
        polys = [rand_mont(1 << l, 10 * l + i) for i in range(k)]
        L.spartan.sumcheck_prove(0, kind, [p.data_ptr() for p in polys], l, 0, chal)
insts = [([rand_mont(1 << l, 7 * l + i).data_ptr() for i in range(4)], l) for l in (6, 2, 0, 6)]
keep = [rand_mont(1 << 6, i) for i in range(16)]
insts = [([keep[4 * j + i][:32 << l].data_ptr() for i in range(4)], l) for j, l in enumerate((6, 2, 0, 6))]
L.spartan.sumcheck_prove_batch(0, L.spartan.CUBIC, insts, [1, 2, 3, 4], [5, 6, 7, 8], chal)
out = torch.empty((1 << 9) * 32, dtype=torch.uint8, device="cuda")
for l in (0, 3, 9):
    L.spartan.eq_evals(0, [5 + i for i in range(l)], out.data_ptr())
a, b = rand_mont(70001, 1), rand_mont(70001, 2)
L.spartan.inner_product(0, a.data_ptr(), b.data_ptr(), 70001)
for curve, logn in ((1, 6), (2, 5)):
    n = 1 << logn
    bases = L.synthetic_bases(curve, n + 1)
    ckk = L.CommitmentKey(curve, bases[:64 * n])
    gc = (int.from_bytes(bases[64 * n:64 * n + 32].tobytes(), "little"), int.from_bytes(bases[64 * n + 32:64 * n + 64].tobytes(), "little"))
    L.spartan.ipa_prove(curve, ckk, gc, rand_mont(n, 3).data_ptr(), rand_mont(n, 4).data_ptr(), logn, chal)
    d = torch.from_numpy(L.synthetic_bases(curve, 16, fmt=L.FMT_MONTGOMERY)).cuda()
    L.spartan.ipa_fold_bases(curve, d.data_ptr(), 16, 12345, 67890)
for l in (1, 6, 11):
    L.spartan.hyperkzg_prove(0, kz if l <= 11 else ck, rand_mont(1 << l, 9).data_ptr(), [3 + i for i in range(l)], chal)
torch.cuda.synchronize()
print("sanitize_n34 done")
