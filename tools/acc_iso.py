"""The dominant kernel alone, as bench.py's `roofline` times it: commit of a dense 1 114 100-term vector (the size of the step's
error vector) with a resident 2^21-point BN254 key and its c = 20 fixed-base table.  Run under ncu by tools/profile_r2.sh."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import lurk_beta_b200 as L

n = 1_114_100
ck = L.CommitmentKey(0, L.synthetic_bases(0, 1 << 21, fmt=L.FMT_MONTGOMERY), fmt=L.FMT_MONTGOMERY).precompute()
ck.set_profiling(True)
rng = np.random.default_rng(1)
sc = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
sc[:, 31] &= 0x1f
d = torch.from_numpy(sc.reshape(-1)).cuda()
for _ in range(4):
    ck.commit_device(d.data_ptr(), n, fmt=L.FMT_MONTGOMERY)
    print("accumulate ms", ck.last_profile()[0], flush=True)
