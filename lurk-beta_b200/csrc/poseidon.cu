// C-ABI front end of the Poseidon digest / slot-witness kernels (S1, S3 in include/lurk_b200.h).
#include "poseidon_api.h"

#include <algorithm>
#include <mutex>

namespace lurk {

int bitdecomp_block_host(const uint32_t mod[8]) {
    uint32_t b[8];
    for (int i = 0; i < 8; i++) b[i] = mod[i];
    b[0] -= 1;
    int cnt = 1, run = 0;
    bool found = false, have_last = false;
    for (int i = 255; i >= 0; i--) {
        uint32_t bb = (b[i >> 5] >> (i & 31)) & 1;
        found |= bb != 0;
        if (!found) continue;
        if (bb) { cnt++; run++; }
        else {
            if (run) { cnt += run - 1 + (have_last ? 1 : 0); have_last = true; run = 0; }
            cnt++;
        }
    }
    return cnt;
}

static bool fmt_ok(int fmt) { return fmt == LURK_FMT_CANONICAL || fmt == LURK_FMT_MONTGOMERY; }
// device-pointer entry points: a null buffer with n > 0 is an argument error, not a kernel fault
static int need(size_t n, const void *a, const void *b) {
    if (n && (!a || !b)) { set_error("null buffer"); return LURK_ERR_ARG; }
    return LURK_OK;
}

// ---- host-buffer driver shared by the S1/S3 entry points.
// Small calls: one copy in, one launch, one copy out.  Large calls are cut into chunks that flow through three staging
// slots (pinned host + device buffers, one stream each): while chunk c is hashed, chunk c+1 is copied from the caller's
// pageable memory into pinned memory and uploaded, and chunk c-1 is downloaded -- the caller's buffers are touched by
// plain memcpy only.  The staging pool is created once per process and reused (guarded by a mutex).
struct StagingSlot {
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    void *h_in = nullptr, *h_out = nullptr, *d_in = nullptr, *d_out = nullptr;
    size_t in_cap = 0, out_cap = 0;
    size_t first = 0, count = 0;
    bool busy = false;
};
struct StagingPool {
    std::mutex mu;
    StagingSlot slot[3];
    int *d_bad = nullptr;
    int device = -1;
};
static StagingPool g_pool;

static int slot_reserve(StagingSlot &sl, size_t in_bytes, size_t out_bytes) {
    if (!sl.stream) {
        LURK_CUDA_TRY(cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
        LURK_CUDA_TRY(cudaEventCreateWithFlags(&sl.done, cudaEventDisableTiming));
    }
    if (sl.in_cap < in_bytes) {
        if (sl.h_in) cudaFreeHost(sl.h_in);
        if (sl.d_in) cudaFree(sl.d_in);
        sl.h_in = sl.d_in = nullptr; sl.in_cap = 0;
        LURK_CUDA_TRY(cudaMallocHost(&sl.h_in, in_bytes));
        LURK_CUDA_TRY(cudaMalloc(&sl.d_in, in_bytes));
        sl.in_cap = in_bytes;
    }
    if (sl.out_cap < out_bytes) {
        if (sl.h_out) cudaFreeHost(sl.h_out);
        if (sl.d_out) cudaFree(sl.d_out);
        sl.h_out = sl.d_out = nullptr; sl.out_cap = 0;
        LURK_CUDA_TRY(cudaMallocHost(&sl.h_out, out_bytes));
        LURK_CUDA_TRY(cudaMalloc(&sl.d_out, out_bytes));
        sl.out_cap = out_bytes;
    }
    return LURK_OK;
}

template <class F, bool WITNESS>
static int run_host(int arity, const uint8_t *pre, size_t n, uint8_t *out, int in_fmt, int out_fmt, size_t out_elems_per) {
    LURK_TRY(require_gpu());
    if (n == 0) return LURK_OK;
    if (!pre || !out) { set_error("null buffer"); return LURK_ERR_ARG; }
    if (arity != 3 && arity != 4 && arity != 6 && arity != 8) { set_error("unsupported Poseidon arity %d", arity); return LURK_ERR_ARG; }
    const size_t in_per = (size_t)arity * sizeof(F), out_per = out_elems_per * sizeof(F);
    const size_t budget = (size_t)48 << 20;                          // bytes of (in + out) per chunk
    size_t chunk = std::max<size_t>(1, budget / (in_per + out_per));
    if (chunk > n) chunk = n;
    std::lock_guard<std::mutex> g(g_pool.mu);
    int dev = 0;
    LURK_CUDA_TRY(cudaGetDevice(&dev));
    if (g_pool.device != dev) {                                      // staging buffers live on one device
        for (auto &sl : g_pool.slot) {
            if (sl.h_in) cudaFreeHost(sl.h_in);
            if (sl.h_out) cudaFreeHost(sl.h_out);
            if (sl.d_in) cudaFree(sl.d_in);
            if (sl.d_out) cudaFree(sl.d_out);
            if (sl.stream) cudaStreamDestroy(sl.stream);
            if (sl.done) cudaEventDestroy(sl.done);
            sl = StagingSlot();
        }
        if (g_pool.d_bad) cudaFree(g_pool.d_bad);
        g_pool.d_bad = nullptr;
        g_pool.device = dev;
    }
    if (!g_pool.d_bad) LURK_CUDA_TRY(cudaMalloc(&g_pool.d_bad, sizeof(int)));
    // the slot streams are non-blocking (they do not order against the legacy stream): clear the counter and wait for it
    // before any chunk's range check can add to it
    LURK_CUDA_TRY(cudaMemsetAsync(g_pool.d_bad, 0, sizeof(int), nullptr));
    LURK_CUDA_TRY(cudaStreamSynchronize(nullptr));
    const size_t nchunks = (n + chunk - 1) / chunk;
    const int nslots = nchunks >= 3 ? 3 : (int)nchunks;
    for (int k = 0; k < nslots; k++) LURK_TRY(slot_reserve(g_pool.slot[k], chunk * in_per, chunk * out_per));
    auto retire = [&](StagingSlot &sl) -> int {
        if (!sl.busy) return LURK_OK;
        LURK_CUDA_TRY(cudaEventSynchronize(sl.done));
        memcpy(out + sl.first * out_per, sl.h_out, sl.count * out_per);
        sl.busy = false;
        return LURK_OK;
    };
    int rc = LURK_OK;
    for (size_t c = 0; c < nchunks && rc == LURK_OK; c++) {
        StagingSlot &sl = g_pool.slot[c % nslots];
        rc = retire(sl);
        if (rc != LURK_OK) break;
        sl.first = c * chunk;
        sl.count = std::min(chunk, n - sl.first);
        memcpy(sl.h_in, pre + sl.first * in_per, sl.count * in_per);
        cudaError_t e = cudaMemcpyAsync(sl.d_in, sl.h_in, sl.count * in_per, cudaMemcpyHostToDevice, sl.stream);
        if (e != cudaSuccess) { set_error("upload failed: %s", cudaGetErrorString(e)); rc = LURK_ERR_CUDA; break; }
        rc = check_reduced_accumulate_dev<F>(sl.d_in, sl.count * arity, sl.stream, g_pool.d_bad);
        if (rc == LURK_OK) rc = launch_poseidon<F, WITNESS>(arity, sl.d_in, sl.count, sl.d_out, in_fmt, out_fmt, sl.stream);
        if (rc != LURK_OK) break;
        e = cudaMemcpyAsync(sl.h_out, sl.d_out, sl.count * out_per, cudaMemcpyDeviceToHost, sl.stream);
        if (e == cudaSuccess) e = cudaEventRecord(sl.done, sl.stream);
        if (e != cudaSuccess) { set_error("download failed: %s", cudaGetErrorString(e)); rc = LURK_ERR_CUDA; break; }
        sl.busy = true;
    }
    for (int k = 0; k < nslots; k++) {
        int r2 = retire(g_pool.slot[k]);
        if (rc == LURK_OK) rc = r2;
        g_pool.slot[k].busy = false;
    }
    if (rc != LURK_OK) { cudaDeviceSynchronize(); return rc; }
    int bad = 0;
    LURK_CUDA_TRY(cudaMemcpy(&bad, g_pool.d_bad, sizeof(int), cudaMemcpyDeviceToHost));
    if (bad) { set_error("%d input element(s) are not reduced below the field modulus", bad); return LURK_ERR_RANGE; }
    return LURK_OK;
}

}  // namespace lurk

using namespace lurk;

extern "C" {

int lurk_poseidon_hash_batch(int field_id, int arity, const uint8_t *preimages, size_t n, uint8_t *digests) {
    return dispatch_field(field_id, [&](auto f) {
        using F = decltype(f);
        return run_host<F, false>(arity, preimages, n, digests, LURK_FMT_CANONICAL, LURK_FMT_CANONICAL, 1);
    });
}
int lurk_poseidon_hash_batch_mont(int field_id, int arity, const uint8_t *preimages, size_t n, uint8_t *digests) {
    return dispatch_field(field_id, [&](auto f) {
        using F = decltype(f);
        return run_host<F, false>(arity, preimages, n, digests, LURK_FMT_MONTGOMERY, LURK_FMT_MONTGOMERY, 1);
    });
}
int lurk_poseidon_hash_batch_dev(int field_id, int arity, const void *d_preimages, size_t n, void *d_digests, int fmt,
                                 void *stream) {
    if (!fmt_ok(fmt)) { set_error("bad format %d", fmt); return LURK_ERR_ARG; }
    LURK_TRY(need(n, d_preimages, d_digests));
    LURK_TRY(require_gpu());
    return dispatch_field(field_id, [&](auto f) {
        using F = decltype(f);
        return launch_poseidon<F, false>(arity, d_preimages, n, d_digests, fmt, fmt, (cudaStream_t)stream);
    });
}

int lurk_poseidon_constants(int field_id, int arity, int *full_rounds, int *partial_rounds, uint8_t *round_constants,
                            uint8_t *mds) {
    if (arity != 3 && arity != 4 && arity != 6 && arity != 8) { set_error("unsupported arity %d", arity); return LURK_ERR_ARG; }
    return dispatch_field(field_id, [&](auto f) {
        using F = decltype(f);
        const PoseidonParams<F> *pp = nullptr;
        poseidon_instance_info<F>(arity, &pp, nullptr);
        const auto &p = *pp;
        if (full_rounds) *full_rounds = p.rf;
        if (partial_rounds) *partial_rounds = p.rp;
        if (round_constants)
            for (size_t i = 0; i < p.round_constants.size(); i++) { F c = p.round_constants[i].to_canonical(); memcpy(round_constants + 32 * i, c.v, 32); }
        if (mds)
            for (size_t i = 0; i < p.mds.size(); i++) { F c = p.mds[i].to_canonical(); memcpy(mds + 32 * i, c.v, 32); }
        return LURK_OK;
    });
}

size_t lurk_poseidon_witness_block(int field_id, int arity) {
    if (arity != 3 && arity != 4 && arity != 6 && arity != 8) return 0;
    size_t out = 0;
    dispatch_field(field_id, [&](auto f) {
        using F = decltype(f);
        PoseidonLayout L;
        poseidon_instance_info<F>(arity, nullptr, &L);
        out = (size_t)L.block_elems;
        return LURK_OK;
    });
    return out;
}
int lurk_poseidon_witness_batch(int field_id, int arity, const uint8_t *preimages, size_t n, uint8_t *blocks, int fmt) {
    if (!fmt_ok(fmt)) { set_error("bad format %d", fmt); return LURK_ERR_ARG; }
    size_t blk = lurk_poseidon_witness_block(field_id, arity);
    if (!blk) { set_error("unsupported field %d / arity %d", field_id, arity); return LURK_ERR_ARG; }
    return dispatch_field(field_id, [&](auto f) {
        using F = decltype(f);
        return run_host<F, true>(arity, preimages, n, blocks, fmt, fmt, blk);
    });
}
int lurk_poseidon_witness_batch_dev(int field_id, int arity, const void *d_preimages, size_t n, void *d_blocks, int fmt,
                                    void *stream) {
    if (!fmt_ok(fmt)) { set_error("bad format %d", fmt); return LURK_ERR_ARG; }
    LURK_TRY(need(n, d_preimages, d_blocks));
    LURK_TRY(require_gpu());
    return dispatch_field(field_id, [&](auto f) {
        using F = decltype(f);
        return launch_poseidon<F, true>(arity, d_preimages, n, d_blocks, fmt, fmt, (cudaStream_t)stream);
    });
}

int lurk_poseidon_witness_scatter_dev(int field_id, int arity, const void *d_preimages, size_t n, void *d_base, const void *d_offsets,
                                      int fmt, void *stream) {
    if (!fmt_ok(fmt)) { set_error("bad format %d", fmt); return LURK_ERR_ARG; }
    if (n && !d_offsets) { set_error("null offsets"); return LURK_ERR_ARG; }
    LURK_TRY(need(n, d_preimages, d_base));
    LURK_TRY(require_gpu());
    return dispatch_field(field_id, [&](auto f) {
        using F = decltype(f);
        return launch_poseidon<F, true>(arity, d_preimages, n, d_base, fmt, fmt, (cudaStream_t)stream, (const uint64_t *)d_offsets);
    });
}
int lurk_bitdecomp_witness_scatter_dev(int field_id, const void *d_values, size_t n, void *d_base, const void *d_offsets, int fmt,
                                       void *stream) {
    if (!fmt_ok(fmt)) { set_error("bad format %d", fmt); return LURK_ERR_ARG; }
    if (n && !d_offsets) { set_error("null offsets"); return LURK_ERR_ARG; }
    LURK_TRY(need(n, d_values, d_base));
    LURK_TRY(require_gpu());
    if (n == 0) return LURK_OK;
    int blk = (int)lurk_bitdecomp_witness_block(field_id);
    return dispatch_field(field_id, [&](auto f) {
        using F = decltype(f);
        return launch_bitdecomp<F>(d_values, n, d_base, blk, fmt, (cudaStream_t)stream, (const uint64_t *)d_offsets);
    });
}

size_t lurk_bitdecomp_witness_block(int field_id) {
    size_t out = 0;
    dispatch_field(field_id, [&](auto f) {
        using F = decltype(f);
        uint32_t m[8];
        for (int i = 0; i < 8; i++) m[i] = F::Params::MOD(i);
        out = (size_t)bitdecomp_block_host(m);
        return LURK_OK;
    });
    return out;
}
int lurk_bitdecomp_witness_batch_dev(int field_id, const void *d_values, size_t n, void *d_blocks, int fmt, void *stream) {
    if (!fmt_ok(fmt)) { set_error("bad format %d", fmt); return LURK_ERR_ARG; }
    LURK_TRY(need(n, d_values, d_blocks));
    LURK_TRY(require_gpu());
    if (n == 0) return LURK_OK;
    int blk = (int)lurk_bitdecomp_witness_block(field_id);
    return dispatch_field(field_id, [&](auto f) {
        using F = decltype(f);
        return launch_bitdecomp<F>(d_values, n, d_blocks, blk, fmt, (cudaStream_t)stream);
    });
}
int lurk_bitdecomp_witness_batch(int field_id, const uint8_t *values, size_t n, uint8_t *blocks, int fmt) {
    if (!fmt_ok(fmt)) { set_error("bad format %d", fmt); return LURK_ERR_ARG; }
    LURK_TRY(need(n, values, blocks));
    LURK_TRY(require_gpu());
    if (n == 0) return LURK_OK;
    size_t blk = lurk_bitdecomp_witness_block(field_id);
    return dispatch_field(field_id, [&](auto f) {
        using F = decltype(f);
        DevBuf din, dout;
        LURK_TRY(din.alloc(n * sizeof(F)));
        LURK_TRY(dout.alloc(n * blk * sizeof(F)));
        LURK_CUDA_TRY(cudaMemcpy(din.p, values, din.bytes, cudaMemcpyHostToDevice));
        int bad = 0;
        LURK_TRY(check_reduced_dev<F>(din.p, n, 0, &bad));
        if (bad) { set_error("%d input element(s) are not reduced below the field modulus", bad); return LURK_ERR_RANGE; }
        LURK_TRY(lurk_bitdecomp_witness_batch_dev(field_id, din.p, n, dout.p, fmt, 0));
        LURK_CUDA_TRY(cudaMemcpy(blocks, dout.p, dout.bytes, cudaMemcpyDeviceToHost));
        return LURK_OK;
    });
}

}  // extern "C"
