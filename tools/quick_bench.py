"""Ad-hoc device timings of the individual kernels (development aid; bench.py is the contract)."""
import sys
import time

import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import lurk_beta_b200 as L
from util import random_elements

lib = L._capi.lib()


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), sorted(ts)[len(ts) // 2]


def poseidon(field, arity, logn):
    n = 1 << logn
    pre = torch.from_numpy(random_elements(field, n * arity, seed=1)).cuda()
    out = torch.empty(n * 32, dtype=torch.uint8, device="cuda")
    fn = lambda: L._capi.check(lib.lurk_poseidon_hash_batch_dev(field, arity, pre.data_ptr(), n, out.data_ptr(), 0, None))
    best, med = timeit(fn)
    print(f"poseidon field={field} arity={arity} n=2^{logn}: best {best:.2f} ms med {med:.2f} ms -> {n / best / 1e3:.2f} Mhash/s, "
          f"{n * (arity + 1) * 32 / best / 1e6:.1f} GB/s algorithmic", flush=True)


def msm(curve, logn, shape, fixed=False):
    n = 1 << logn
    t0 = time.time()
    bases = L.synthetic_bases(curve, n)
    sc = random_elements([0, 1, 2, 3][curve], n, seed=2, shape=shape)
    ck = L.CommitmentKey(curve, bases)
    if fixed:
        ck.precompute()
    d_sc = torch.from_numpy(sc).cuda()
    out = None

    def fn():
        nonlocal out
        out = ck.commit_device(d_sc.data_ptr(), n, fmt=0)
    best, med = timeit(fn, reps=5, warm=2)
    print(f"msm curve={curve} n=2^{logn} {shape}{' fixed-base' if fixed else ''}: best {best:.2f} ms med {med:.2f} ms -> {n / best / 1e3:.2f} Mterm/s, "
          f"{n * 96 / best / 1e6:.1f} GB/s algorithmic (setup {time.time() - t0:.1f}s)", flush=True)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "poseidon"):
        for f, a, ln in [(0, 8, 20), (0, 8, 22), (0, 4, 22), (2, 8, 22), (0, 3, 20), (0, 6, 20)]:
            poseidon(f, a, ln)
        for ln in (8, 11):
            poseidon(0, 4, ln)
    if which == "poseidon8":
        poseidon(0, 8, 20)
    if which == "msm21":
        msm(0, 21, "witness")
    if which in ("all", "msm"):
        for c, ln, sh in [(0, 16, "uniform"), (0, 20, "uniform"), (0, 20, "witness"), (2, 20, "uniform"), (0, 21, "witness")]:
            msm(c, ln, sh)
        for c, ln, sh in [(0, 20, "uniform"), (0, 21, "witness")]:
            msm(c, ln, sh, fixed=True)
