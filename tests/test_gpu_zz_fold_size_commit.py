"""The commitment of the fold step at bench.py's exact size, on the GPU (kept in its own file, sorted last: it was added
after the round's GPU budget was spent, so its first run is the round-end suite)."""
import numpy as np
import pytest

from util import ints, pack, random_elements

pytestmark = pytest.mark.gpu


def test_msm_fixed_base_fold_size_bn254(L, oracle, spec):
    """the commitment of the fold step exactly as bench.py runs it: 2^21-point BN254 key with the fixed-base table (window
    c = 20, one shared set of 2^19 buckets, 32-slice bucket reduction), 911 900 witness-shaped scalars.  The table path
    must give the same point as the plain path on the same key, be linear, and match the oracle on a prefix."""
    curve, n_key, n = 0, 1 << 21, 911_900
    sf = spec.CURVES[curve]["scalar"]
    bases = L.synthetic_bases(curve, n_key)
    plain = L.CommitmentKey(curve, bases)
    fixed = L.CommitmentKey(curve, bases).precompute()
    a = random_elements(sf, n, seed=41, shape="witness")
    b = random_elements(sf, n, seed=42, shape="uniform")
    ca, cb = fixed.commit(a), fixed.commit(b)
    assert ca[64] == 1 and spec.on_curve(curve, tuple(ints(ca[:64])))
    assert np.array_equal(ca, plain.commit(a)) and np.array_equal(cb, plain.commit(b))
    ab = oracle.axpy(sf, a, b, pack([1]), nthreads=8)
    assert np.array_equal(L.point_sum(curve, np.concatenate([ca, cb])), fixed.commit(ab))          # linearity
    m = 1 << 14
    assert np.array_equal(fixed.commit(a[:32 * m]), oracle.msm(curve, bases[:64 * m], a[:32 * m], nthreads=8))
    full = random_elements(sf, n_key, seed=43, shape="uniform")                                    # every base of the key
    assert np.array_equal(fixed.commit(full), plain.commit(full))


def test_plain_c_client_on_the_gpu(tmp_path):
    """the gcc-built C client (tests/csrc/c_abi_client.c) run on the GPU box: golden G1 through a process that contains no
    Python, torch or C++ of ours -- only the C ABI"""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe, libdir = str(tmp_path / "c_abi_client"), os.path.join(root, "lurk-beta_b200")
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-I", os.path.join(root, "include"), os.path.join(root, "tests", "csrc", "c_abi_client.c"),
                           "-o", exe, "-L", libdir, "-llurk_b200", "-Wl,-rpath," + libdir])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    assert "c_abi_client ok" in out.stdout


def test_plain_c_fold_driver_on_the_gpu(tmp_path):
    """tests/csrc/fold_client.c on the GPU box: a four-step IVC chain through lurk_fold_ctx_* from plain C99 -- device-side
    relaxed-R1CS check, u = 1 + sum of the challenges, checkpoint / resume -- with no Python, torch or oracle in the process"""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe, libdir = str(tmp_path / "fold_client"), os.path.join(root, "lurk-beta_b200")
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-I", os.path.join(root, "include"), os.path.join(root, "tests", "csrc", "fold_client.c"),
                           "-o", exe, "-L", libdir, "-llurk_b200", "-Wl,-rpath," + libdir])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    assert out.stdout.strip() == "fold_client ok"
