"""Device-resident Nova fold pipeline: the GPU half of `Proof::prove_recursively` (reference src/proof/nova.rs:260-339).

The reference runs two threads over a bounded channel (nova.rs:297-326): a witness thread that calls
`step.cache_witness(store)` for later steps, and the fold thread that calls `RecursiveSNARK::prove_step`.  Here the same
split is two stages on CUDA streams, with the running instance (W1, E1) and the fresh witness resident in HBM between
steps:

  stage A (chain independent, runs one step ahead):   slot witnesses -> W2 (src/lem/multiframe.rs:520-592),
        comm_W = commit(W2) enqueued (Arecibo commit, src/proof/nova.rs:287,292), Az2, Bz2, Cz2
  stage B (the sequential fold chain):                Az1, Bz1, Cz1, cross term T, comm_T = commit(T), exchange of the
        partial commitments when the key is sharded, challenge r, W1 <- W1 + r W2, E1 <- E1 + r T   

What stays with the caller (CPU, out of scope here): the LEM body aux of every frame (the "glue" part of W2), the Nova
augmented-circuit part of the witness, and the random oracle that turns the commitments into the challenge `r` -- the
pipeline takes `challenge(comm_W, comm_T) -> 32 bytes (Montgomery)` as a callback.

All vectors are Montgomery-form device buffers passed as torch uint8 tensors; this module does no arithmetic itself.
"""
import ctypes as C

import numpy as np

from . import _capi
from .commit import point_sum


class SlotBatch:
    """slot preimages of one step for one slot type, and where their witness blocks go inside W"""

    def __init__(self, arity, count, offset_elems, d_preimages, d_offsets=None):
        """arity 0 = BitDecomp.  Blocks go to W contiguously from `offset_elems`, or -- the reference's real layout, every
        frame's aux = [its slot blocks | LEM body aux] (src/lem/multiframe.rs:635-712) -- block k to element offset
        d_offsets[k] (device tensor of u64)."""
        self.arity, self.count, self.offset, self.d_pre, self.d_offsets = arity, count, offset_elems, d_preimages, d_offsets


class NovaFoldPipeline:
    def __init__(self, torch, field_id, curve_id, ck, n_w, n_t, csr, u1, u2, z1, E1, z2_buffers, world=1, group=None):
        """ck: CommitmentKey (this rank's shard); csr: three (row_ptr, col, val) device CSR matrices (A, B, C) with n_t rows
        over z = (W, u, X); z1 / z2_buffers[b]: device vectors of len(z) elements whose first n_w elements are W1 / W2[b];
        E1: n_t elements; u1, u2: 32-byte Montgomery host arrays."""
        self.t, self.lib = torch, _capi.lib()
        self.field_id, self.curve_id = field_id, curve_id
        self.n_w, self.n_t = n_w, n_t
        self.csr, self.u1, self.u2 = csr, u1, u2
        self.z1, self.E1, self.z2 = z1, E1, z2_buffers
        self.W1 = z1[:n_w * 32]
        self.W2 = [z[:n_w * 32] for z in z2_buffers]
        self.world, self.group = world, group
        self.T = torch.empty(n_t * 32, dtype=torch.uint8, device="cuda")
        self.mv1 = [torch.empty(n_t * 32, dtype=torch.uint8, device="cuda") for _ in range(3)]
        nb = len(z2_buffers)                                    # stage A may run nb - 1 steps ahead of stage B
        self.mv2 = [[torch.empty(n_t * 32, dtype=torch.uint8, device="cuda") for _ in range(3)] for _ in range(nb)]
        # The fold chain (stage B) is the critical path: its stream gets the high CUDA priority so that the short,
        # latency-bound tail kernels of commit(T) are not queued behind the thousands of CTAs of a prefetched commit(W).
        self.sK = [torch.cuda.Stream(priority=0) for _ in range(3)]   # slot-witness kernels (one stream per slot type), commit(W)
        self.sA = torch.cuda.Stream(priority=0)                       # Az2, Bz2, Cz2 of the prefetched step
        self.sB = torch.cuda.Stream(priority=-1)                      # Az1.., cross term, commit(T), fold
        self.ckW = [ck] + [ck.clone() for _ in range(nb - 1)]
        self.ckT = ck.clone()
        for c in (*self.ckW, self.ckT):
            c.set_profiling(True)
        self.ev_fold = [None] * nb                             # fold that last read W2[b]
        self.ev_A = [None] * nb                                # stage A of buffer b complete (Az2.. ready)
        self.accumulate_ms = []                                # device time of the dominant kernel, per commitment
        self.launches_A = self.launches_B = 0

    # ---------------------------------------------------------------------------------------------- stage A
    def stage_a(self, b, slot_batches, before=None):
        """enqueue the chain-independent half for buffer b.  `before(b)`: optional hook run first on the current stream
        (e.g. the host->device copy of this step's preimages and glue aux)."""
        t, lib, chk = self.t, self.lib, _capi.check
        M = _capi.FMT_MONTGOMERY
        cur = t.cuda.current_stream()
        if before is not None:
            if self.ev_fold[b] is not None:
                cur.wait_event(self.ev_fold[b])                # W2[b] is still read by an earlier fold
            before(b)
        for st in (*self.sK, self.sA):
            st.wait_stream(cur)
            if self.ev_fold[b] is not None:
                st.wait_event(self.ev_fold[b])
        W2 = self.W2[b]
        k = 0
        for idx, sb in enumerate(slot_batches):
            st = self.sK[idx % len(self.sK)]
            dst = W2.data_ptr() + sb.offset * 32
            cs = C.c_void_p(st.cuda_stream)
            if sb.d_offsets is not None:
                if sb.arity:
                    chk(lib.lurk_poseidon_witness_scatter_dev(self.field_id, sb.arity, sb.d_pre.data_ptr(), sb.count, W2.data_ptr(),
                                                              sb.d_offsets.data_ptr(), M, cs))
                else:
                    chk(lib.lurk_bitdecomp_witness_scatter_dev(self.field_id, sb.d_pre.data_ptr(), sb.count, W2.data_ptr(),
                                                               sb.d_offsets.data_ptr(), M, cs))
            elif sb.arity:
                chk(lib.lurk_poseidon_witness_batch_dev(self.field_id, sb.arity, sb.d_pre.data_ptr(), sb.count, dst, M, cs))
            else:
                chk(lib.lurk_bitdecomp_witness_batch_dev(self.field_id, sb.d_pre.data_ptr(), sb.count, dst, M, cs))
            k += 1
        evs = []
        for st in self.sK:
            e = t.cuda.Event()
            e.record(st)
            evs.append(e)
        for e in evs[1:]:
            self.sK[0].wait_event(e)
        self.ckW[b].launch_device(W2.data_ptr(), self.n_w, fmt=M, stream=self.sK[0].cuda_stream)
        sa = C.c_void_p(self.sA.cuda_stream)
        for e in evs:
            self.sA.wait_event(e)
        for i, (rp, col, val) in enumerate(self.csr):
            chk(lib.lurk_spmv_csr_dev(self.field_id, rp.data_ptr(), col.data_ptr(), val.data_ptr(), self.n_t, self.z2[b].data_ptr(),
                                      self.mv2[b][i].data_ptr(), sa))
            k += 1
        self.ev_A[b] = t.cuda.Event()
        self.ev_A[b].record(self.sA)
        self.launches_A = k

    # ---------------------------------------------------------------------------------------------- stage B
    def stage_b_launch(self, b):
        """enqueue the chain-dependent kernels of the step on buffer b: Az1.., cross term, commit(T).  Call it right after
        the previous step's fold has been enqueued (stage_b_collect) so the GPU never waits for the host."""
        lib, chk = self.lib, _capi.check
        M = _capi.FMT_MONTGOMERY
        sb = C.c_void_p(self.sB.cuda_stream)
        k = 0
        for i, (rp, col, val) in enumerate(self.csr):
            chk(lib.lurk_spmv_csr_dev(self.field_id, rp.data_ptr(), col.data_ptr(), val.data_ptr(), self.n_t, self.z1.data_ptr(),
                                      self.mv1[i].data_ptr(), sb))
            k += 1
        self.sB.wait_event(self.ev_A[b])
        az1, bz1, cz1 = self.mv1
        az2, bz2, cz2 = self.mv2[b]
        chk(lib.lurk_cross_term_dev(self.field_id, az1.data_ptr(), bz1.data_ptr(), cz1.data_ptr(), az2.data_ptr(), bz2.data_ptr(),
                                    cz2.data_ptr(), _capi.np_ptr(self.u1), _capi.np_ptr(self.u2), self.n_t, self.T.data_ptr(), sb))
        k += 1
        self.ckT.launch_device(self.T.data_ptr(), self.n_t, fmt=M, stream=self.sB.cuda_stream)
        self._k_launch = k

    def stage_b_collect(self, b, challenge):
        """wait for commit(W2[b]) and commit(T), exchange partial commitments if the key is sharded, derive the challenge
        and enqueue the fold; returns (comm_W, comm_T) as 96-byte points (Montgomery)"""
        t, lib, chk = self.t, self.lib, _capi.check
        M = _capi.FMT_MONTGOMERY
        sb = C.c_void_p(self.sB.cuda_stream)
        k = self._k_launch
        cw = self.ckW[b].finish()
        ms, kl = self.ckW[b].last_profile()
        self.accumulate_ms.append(ms)
        k += kl
        ct = self.ckT.finish()
        ms, kl = self.ckT.last_profile()
        self.accumulate_ms.append(ms)
        k += kl
        if self.world > 1:
            # sharded key: all-gather the two 96-byte partial commitments, add them locally (no EC reduction op in NCCL)
            import torch.distributed as dist
            mine = t.from_numpy(np.concatenate([cw, ct])).cuda()
            allp = t.empty(192 * self.world, dtype=t.uint8, device="cuda")
            dist.all_gather_into_tensor(allp, mine, group=self.group)
            allp = allp.cpu().numpy().reshape(self.world, 2, 96)
            cw = point_sum(self.curve_id, allp[:, 0, :].reshape(-1), fmt=M)
            ct = point_sum(self.curve_id, allp[:, 1, :].reshape(-1), fmt=M)
        r = np.ascontiguousarray(challenge(cw, ct), dtype=np.uint8)
        chk(lib.lurk_axpy_dev(self.field_id, self.W1.data_ptr(), self.W2[b].data_ptr(), _capi.np_ptr(r), self.n_w, self.W1.data_ptr(), sb))
        chk(lib.lurk_axpy_dev(self.field_id, self.E1.data_ptr(), self.T.data_ptr(), _capi.np_ptr(r), self.n_t, self.E1.data_ptr(), sb))
        k += 2
        self.ev_fold[b] = t.cuda.Event()
        self.ev_fold[b].record(self.sB)
        self.launches_B = k
        return cw, ct

    def stage_b(self, b, challenge):
        """launch + collect in one call (no software pipelining of the host side)"""
        self.stage_b_launch(b)
        return self.stage_b_collect(b, challenge)

    def drain(self, b):
        """collect a prefetched commit(W) that will not be folded"""
        return self.ckW[b].finish()


class SuperNovaFoldPipeline:
    """NIVC form (reference src/proof/supernova.rs:207-291): one running instance per circuit -- the Lurk step circuit
    plus one per coprocessor -- and every step folds into the instance selected by `MultiFrame::circuit_index()`
    (src/lem/multiframe.rs:941).  Each circuit has its own R1CS shape, witness length and buffers; they share the
    device-resident commitment key (sized for the largest circuit).  Stage A of the next step may belong to a different
    circuit than the step being folded; the two never touch the same buffers."""

    def __init__(self, pipelines):
        self.pipelines = list(pipelines)          # NovaFoldPipeline per circuit index, all built on clones of one key
        self._buf = [0] * len(self.pipelines)     # next double-buffer slot per circuit

    def stage_a(self, circuit_index, slot_batches, before=None):
        p = self.pipelines[circuit_index]
        b = self._buf[circuit_index]
        p.stage_a(b, slot_batches, before)
        self._buf[circuit_index] ^= 1
        return b

    def stage_b(self, circuit_index, b, challenge):
        return self.pipelines[circuit_index].stage_b(b, challenge)
