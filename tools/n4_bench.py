#!/usr/bin/env python3
"""Measurements of the N3 / N4 rows (one JSON object per line): powers-of-tau key, eq table, inner product, sum-check (cubic =
outer, quad = inner), IPA rounds and the HyperKZG opening at the sizes of a fib rc = 100 step circuit (2^21 rows).
The challenge callback is a C function (built here with gcc) so that no Python runs inside the timed calls; wall-clock around whole
prover calls (they synchronise internally once per round), CUDA events around single kernels.
    python tools/n4_bench.py [--logn 21] [--only sumcheck|ipa|kzg|small]"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lurk_beta_b200 as L

lib = L._capi.lib()
chk = L._capi.check
PEAK = 6650.0
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass

CB_SRC = r"""
#include <stdint.h>
#include <string.h>
/* stand-in verifier: 120-bit challenge mixed from the message (never zero, always < p) */
int challenge(void *user, int round, const uint8_t *msg, size_t len, uint8_t out[32]) {
    uint64_t h = 1469598103934665603ull ^ (uint64_t)round;
    for (size_t i = 0; i < len; i++) { h ^= msg[i]; h *= 1099511628211ull; }
    memset(out, 0, 32);
    memcpy(out, &h, 8);
    h = h * 6364136223846793005ull + 1442695040888963407ull;
    memcpy(out + 8, &h, 7);
    out[0] |= 1;
    (*(int *)user)++;
    return 0;
}
"""


def c_callback():
    d = tempfile.mkdtemp()
    src, so = os.path.join(d, "cb.c"), os.path.join(d, "libcb.so")
    open(src, "w").write(CB_SRC)
    subprocess.check_call(["/usr/bin/gcc", "-O2", "-shared", "-fPIC", src, "-o", so])
    cl = C.CDLL(so)
    return cl, C.cast(cl.challenge, L._capi.CHALLENGE_FN)


def rand_mont(n, seed):
    """n uniformly random reduced elements, used directly as Montgomery-form device data"""
    rng = np.random.default_rng(seed)
    raw = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    raw[:, 31] &= 0x1f
    return torch.from_numpy(raw.reshape(-1)).cuda()


def emit(**kw):
    print(json.dumps(kw), flush=True)


def wall(fn, reps=3):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    return sorted(ts)[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--logn", type=int, default=21)
    ap.add_argument("--only", default="all")
    a = ap.parse_args()
    cl, cb = c_callback()
    calls = C.c_int(0)
    user = C.cast(C.pointer(calls), C.c_void_p)
    l = a.logn
    n = 1 << l
    zero32 = np.zeros(32, dtype=np.uint8)

    if a.only in ("all", "small"):
        tau = rand_mont(l, 1).cpu().numpy()
        out = torch.empty(n * 32, dtype=torch.uint8, device="cuda")
        ms = wall(lambda: chk(lib.lurk_eq_evals_dev(0, L._capi.np_ptr(tau), l, C.c_void_p(out.data_ptr()), 1, None)))
        emit(kernel="eq_kernel (EqPolynomial::evals)", log_n=l, ms=round(ms, 3), gb_s=round(n * 32 / ms / 1e6, 1), hbm_frac=round(n * 32 / ms / 1e6 / PEAK, 4))
        x, y = rand_mont(n, 2), rand_mont(n, 3)
        r = np.zeros(32, dtype=np.uint8)
        ms = wall(lambda: chk(lib.lurk_inner_product_dev(0, C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr()), n, L._capi.np_ptr(r), 1, None)))
        emit(kernel="dot_kernel (inner product, whole call incl. scratch + D2H)", log_n=l, ms=round(ms, 3), gb_s=round(n * 64 / ms / 1e6, 1),
             hbm_frac=round(n * 64 / ms / 1e6 / PEAK, 4))

    if a.only in ("all", "sumcheck"):
        for kind, k, name in ((1, 4, "cubic with additive term (outer sum-check: eq, Az, Bz, uCz+E)"), (0, 2, "quadratic (inner sum-check)")):
            src = [rand_mont(n, 10 + i) for i in range(k)]
            work = [s.clone() for s in src]
            ptrs = (C.c_void_p * k)(*[C.c_void_p(w.data_ptr()) for w in work])

            def run():
                for w, s in zip(work, src):
                    w.copy_(s)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                chk(lib.lurk_sumcheck_prove_dev(0, kind, ptrs, l, L._capi.np_ptr(zero32), cb, user, None, None, None, 1, None))
                return (time.perf_counter() - t0) * 1e3
            run()
            ts = sorted(run() for _ in range(5))
            ms = ts[2]
            # bytes: first pass reads k n elements; every later (fused) pass reads the current length and writes half of it
            alg = k * 32 * (n + sum((n >> j) + (n >> (j + 1)) for j in range(0, l)))
            emit(op="lurk_sumcheck_prove_dev", kind=name, log_n=l, rounds=l, polys=k, ms=round(ms, 3), ms_best=round(ts[0], 3),
                 algorithmic_gb_s=round(alg / ms / 1e6, 1), hbm_frac=round(alg / ms / 1e6 / PEAK, 4),
                 note="whole call: scratch set-up, one launch + one 64..96-byte D2H + C callback per round; wall-clock")

    if a.only in ("all", "ipa"):
        for curve, cname, ll in ((1, "grumpkin (secondary circuit, 2^14)", 14), (2, "pallas 2^18", 18)):
            m = 1 << ll
            bases = torch.from_numpy(L.synthetic_bases(curve, m + 1, fmt=L.FMT_MONTGOMERY)).cuda()
            gc = bases[64 * m:64 * (m + 1)].cpu().numpy()
            av, bv = rand_mont(m, 20), rand_mont(m, 21)

            ckk = L.CommitmentKey.from_device(curve, bases.data_ptr(), m)

            def run():
                aa, bb = av.clone(), bv.clone()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                chk(lib.lurk_ipa_prove_dev(curve, ckk._ctx, L._capi.np_ptr(gc), C.c_void_p(aa.data_ptr()), C.c_void_p(bb.data_ptr()), ll, cb, user,
                                           None, None, None, None, 1, None))
                return (time.perf_counter() - t0) * 1e3
            for fixed in (False, True):
                if fixed:
                    ckk.precompute()
                run()
                ts = sorted(run() for _ in range(3))
                emit(op="lurk_ipa_prove_dev", curve=cname, log_n=ll, fixed_base_table=fixed, ms=round(ts[1], 2), ms_best=round(ts[0], 2),
                     note="per round: 2 inner products, 2 Pippenger passes over the fixed key (n/2 non-zero weighted scalars each), folds of a and b, weight update")

    if a.only in ("all", "kzg"):
        curve = 0
        g = L.synthetic_bases(curve, 1, start=41, fmt=L.FMT_MONTGOMERY)
        beta = rand_mont(1, 30).cpu().numpy()
        key = torch.empty(n * 64, dtype=torch.uint8, device="cuda")
        ms = wall(lambda: chk(lib.lurk_ck_powers_dev(curve, L._capi.np_ptr(g), L._capi.np_ptr(beta), n, C.c_void_p(key.data_ptr()), 1, None)))
        emit(op="lurk_ck_powers_dev (powers-of-tau key, whole call incl. the host-built window table)", curve="bn254_g1", log_n=l, ms=round(ms, 2),
             mpoints_per_s=round(n / ms / 1e3, 2))
        ck = L.CommitmentKey.from_device(curve, key.data_ptr(), n)
        poly = rand_mont(n, 31)
        point = rand_mont(l, 32).cpu().numpy()

        def run():
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            chk(lib.lurk_hyperkzg_prove_dev(curve, ck._ctx, C.c_void_p(poly.data_ptr()), L._capi.np_ptr(point), l, cb, user, None, None, None, 1, None))
            return (time.perf_counter() - t0) * 1e3
        for fixed in (False, True):
            if fixed:
                ck.precompute()
            run()
            ts = sorted(run() for _ in range(3))
            emit(op="lurk_hyperkzg_prove_dev", curve="bn254_g1", log_n=l, fixed_base_table=fixed, ms=round(ts[1], 2), ms_best=round(ts[0], 2),
                 note="l - 1 folds + l - 1 commitments (n/2 .. 2 terms), 3 l evaluations, batched polynomial, 3 witness polynomials + 3 n-term commitments")
    try:
        q = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_throttle_reasons.active", "--format=csv,noheader", "-i", "0"],
                           capture_output=True, text=True, timeout=10).stdout.strip()
    except Exception as e:
        q = str(e)
    emit(callbacks=calls.value, nvidia_smi_after=q)


if __name__ == "__main__":
    main()
