"""Hardware parity of the N-GPU path (one process per GPU, torchrun): tests/multi_gpu_worker.py folds an IVC chain with the
step circuit, the witness and the commitment key sharded by frame over the ranks -- the partial commitments exchanged
through NVLink peer memory inside the challenge kernel -- and asserts every record against the oracle's unsharded fold, the
one-GPU result and the device-side relaxed-R1CS check.  Skipped on boxes with fewer than two GPUs."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_fold_matches_oracle_and_single_gpu(L, world):
    n = L._capi.lib().lurk_device_count()
    if n < world:
        pytest.skip(f"needs {world} GPUs, box has {n}")
    env = dict(os.environ, OMP_NUM_THREADS="8")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                          "--master-port", str(29650 + world), os.path.join(ROOT, "tests", "multi_gpu_worker.py")],
                         capture_output=True, text=True, timeout=900, cwd=ROOT, env=env)
    assert out.returncode == 0, (out.stdout[-3000:], out.stderr[-6000:])
    assert out.stdout.count("multi-GPU fold parity ok") == world
