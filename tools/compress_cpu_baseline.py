#!/usr/bin/env python3
"""CPU port of the `compress` chain timed beside the GPU numbers of tools/compress_bench.py: the oracle's C restatements (oracle/oracle.c:
sum-check rounds, SpMV, Pippenger MSM; OpenMP) on the host threads, at the fib rc = 100 shape.  The MSM legs are timed on a bounded sample
(2^18 terms) and scaled linearly to the terms one opening commits (4 n: the fold chain sums to n, plus three n-term witness commitments);
everything else runs at full size.  One JSON line.  Runs without a GPU."""
import argparse
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from oracle import capi as oracle  # noqa: E402
from util import random_elements  # noqa: E402


def chal(rnd, ev):
    return int.from_bytes(hashlib.sha256(repr((rnd, ev)).encode()).digest()[:30], "little")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rc", type=int, default=100)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    a = ap.parse_args()
    th = a.threads
    mats, n_w, rows, _ = bench.step_circuit(1, a.rc)
    s = max(1, (rows - 1).bit_length())
    t = max(1, (n_w - 1).bit_length())
    out = {"op": "compress, primary circuit, CPU port (oracle/oracle.c, OpenMP)", "rc": a.rc, "threads": th, "rows_padded_log2": s, "vars_padded_log2": t}
    z = random_elements(0, n_w + 3, seed=1)
    t0 = time.perf_counter()
    for rp, col, val in mats:
        oracle.spmv(0, rp, col, val, z, nthreads=th)
    out["multiply_vec_s"] = round(time.perf_counter() - t0, 3)
    out["eval_table_s_estimate"] = out["multiply_vec_s"]          # the transposed products move the same non-zeros
    for kind, k, l in (("cubic", 4, s), ("quad", 2, t + 1)):
        bufs = [random_elements(0, 1 << l, seed=10 + i) for i in range(k)]
        t0 = time.perf_counter()
        oracle.sumcheck_prove(0, kind, bufs, l, 0, chal, nthreads=th)
        out[f"sumcheck_{kind}_2^{l}_s"] = round(time.perf_counter() - t0, 3)
    m = 1 << 18
    bases = oracle.gen_bases(0, m)
    sc = random_elements(0, m, seed=3)
    t0 = time.perf_counter()
    oracle.msm(0, bases, sc, nthreads=th)
    per_term = (time.perf_counter() - t0) / m
    out["msm_sample_terms"] = m
    out["msm_us_per_term"] = round(per_term * 1e6, 3)
    out["open_W_s_estimate"] = round(per_term * 4 * (1 << t), 2)
    out["open_E_s_estimate"] = round(per_term * 4 * (1 << s), 2)
    out["total_s_estimate"] = round(out["multiply_vec_s"] * 2 + out[f"sumcheck_cubic_2^{s}_s"] + out[f"sumcheck_quad_2^{t + 1}_s"] + out["open_W_s_estimate"] +
                                    out["open_E_s_estimate"], 2)
    out["note"] = "sum-check legs include the conversion of their inputs into Montgomery form; upstream's CPU MSM uses hand-written asm and is faster than this port"
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
