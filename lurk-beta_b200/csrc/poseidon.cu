// C-ABI front end of the Poseidon digest / slot-witness kernels (S1, S3 in include/lurk_b200.h).
#include "poseidon_api.h"

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

// host-buffer driver shared by the S1/S3 entry points
template <class F, bool WITNESS>
static int run_host(int arity, const uint8_t *pre, size_t n, uint8_t *out, int in_fmt, int out_fmt, size_t out_elems_per) {
    LURK_TRY(require_gpu());
    if (n == 0) return LURK_OK;
    if (!pre || !out) { set_error("null buffer"); return LURK_ERR_ARG; }
    DevBuf din, dout;
    LURK_TRY(din.alloc(n * arity * sizeof(F)));
    LURK_TRY(dout.alloc(n * out_elems_per * sizeof(F)));
    LURK_CUDA_TRY(cudaMemcpy(din.p, pre, din.bytes, cudaMemcpyHostToDevice));
    int bad = 0;
    LURK_TRY(check_reduced_dev<F>(din.p, n * arity, 0, &bad));
    if (bad) { set_error("%d input element(s) are not reduced below the field modulus", bad); return LURK_ERR_RANGE; }
    LURK_TRY((launch_poseidon<F, WITNESS>(arity, din.p, n, dout.p, in_fmt, out_fmt, 0)));
    LURK_CUDA_TRY(cudaMemcpy(out, dout.p, dout.bytes, cudaMemcpyDeviceToHost));
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
    LURK_TRY(require_gpu());
    return dispatch_field(field_id, [&](auto f) {
        using F = decltype(f);
        return launch_poseidon<F, true>(arity, d_preimages, n, d_blocks, fmt, fmt, (cudaStream_t)stream);
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
