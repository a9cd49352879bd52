// N4 -- the folding rounds of the inner-product argument behind the C ABI (include/lurk_b200.h): Arecibo
// provider::ipa_pc::InnerProductArgument::prove as `compress` reaches it for the secondary circuit (EE2) and for the Pasta cycle
// (reference src/proof/nova.rs:57-71, 341-356).  Split from sumcheck.cu to keep the two translation units' nvcc time apart.
//   ipa_fold_scalars / _bases    a' = x a_lo + y a_hi;  G' = x G_lo + y G_hi (interleaved double-and-add, uniform branches).
//   ipa_weighted / weights_update the prover's own path: commitments of a round as Pippenger passes over the fixed key.
#include "common.cuh"
#include "sumcheck.cuh"
#include "sc_scratch.cuh"

#include <algorithm>
#include <vector>

namespace lurk {

// ------------------------------------------------------------------------------------------------ IPA folds
template <class F>
__global__ void __launch_bounds__(256) ipa_fold_scalars_kernel(F *a, size_t half, F x, F y) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < half; i += (size_t)gridDim.x * blockDim.x)
        store_fe(a + i, ipa_fold_scalar(load_fe<F>(a + i), load_fe<F>(a + i + half), x, y));
}
struct Scalar256 { uint32_t w[8]; };
template <class F>
__global__ void __launch_bounds__(128) ipa_fold_bases_kernel(Affine<F> *g, size_t half, const __grid_constant__ Scalar256 x, const __grid_constant__ Scalar256 y) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < half; i += (size_t)gridDim.x * blockDim.x) {
        Affine<F> p, q;
        p.x = load_fe<F>(&g[i].x); p.y = load_fe<F>(&g[i].y);
        q.x = load_fe<F>(&g[i + half].x); q.y = load_fe<F>(&g[i + half].y);
        const Affine<F> r = ipa_fold_point(p, q, x.w, y.w);
        store_fe(&g[i].x, r.x);
        store_fe(&g[i].y, r.y);
    }
}

template <class Fb>
static void point_to_bytes_fmt(const XYZZ<Fb> &p, int fmt, uint8_t out[96]) {
    memset(out, 0, 96);
    if (p.is_identity()) return;
    Affine<Fb> a = p.to_affine();
    Fb one = Fb::one();
    if (fmt == LURK_FMT_CANONICAL) { a.x = a.x.to_canonical(); a.y = a.y.to_canonical(); one = one.to_canonical(); }
    memcpy(out, a.x.v, 32); memcpy(out + 32, a.y.v, 32); memcpy(out + 64, one.v, 32);
}

// Fixed-base multiplication of ck_c on the host (the c_L ck_c / c_R ck_c terms, 2 per round): 4-bit windows, 64 mixed additions
// per product instead of a 254-step double-and-add.
template <class Fb>
struct HostFixedBase {
    std::vector<Affine<Fb>> table;     // table[w * 15 + d - 1] = d 16^w P
    explicit HostFixedBase(const Affine<Fb> &p) : table(64 * 15) {
        std::vector<XYZZ<Fb>> pts(64 * 15);
        Affine<Fb> base = p;
        for (int w = 0; w < 64; w++) {
            XYZZ<Fb> acc = XYZZ<Fb>::identity();
            for (int d = 1; d <= 15; d++) { acc.add_affine(base); pts[w * 15 + d - 1] = acc; }
            XYZZ<Fb> nb = acc;
            nb.add_affine(base);
            base = nb.to_affine();
        }
        std::vector<Fb> pref(pts.size());
        Fb run = Fb::one();
        for (size_t i = 0; i < pts.size(); i++) { pref[i] = run; if (!pts[i].is_identity()) run = run * pts[i].zzz; }
        Fb inv = run.inv();
        for (size_t i = pts.size(); i-- > 0;) {
            if (pts[i].is_identity()) { table[i].x = Fb::zero(); table[i].y = Fb::zero(); continue; }
            const Fb zi = inv * pref[i];
            inv = inv * pts[i].zzz;
            const Fb zz_inv = (zi * pts[i].zz).sqr();
            table[i].x = pts[i].x * zz_inv;
            table[i].y = pts[i].y * zi;
        }
    }
    XYZZ<Fb> mul(const uint32_t k[8]) const {       // k canonical
        XYZZ<Fb> acc = XYZZ<Fb>::identity();
        for (int w = 0; w < 64; w++) {
            const uint32_t d = (k[w >> 3] >> (4 * (w & 7))) & 15u;
            if (d) acc.add_affine(table[w * 15 + d - 1]);
        }
        return acc;
    }
};

// The prover never needs the folded key itself, only commitments under it: with W_j[idx] = prod_{k < j} (bit_k(idx) ? r_k : 1 / r_k)
// (bit_k = the k-th bit of idx from the top) the folded key of round j is G_j[i] = sum_{idx = i mod m} W_j[idx] G[idx], m = n / 2^j, so
//     L_j = <a_lo, G_j,hi> = sum_{idx : idx mod m >= m/2} W_j[idx] a_j[idx mod m - m/2] G[idx]      (R_j alike on the low halves)
// -- one Pippenger pass over the ORIGINAL key per commitment (the bucket sort drops the zero half) instead of m / 2 latency-bound
// 254-bit double-scalar multiplications per round; the key is not consumed and a fixed-base table of it can be reused.
template <class F>
__global__ void __launch_bounds__(256) ipa_weighted_kernel(const F *__restrict__ w, const F *__restrict__ a, size_t n, size_t m, F *__restrict__ sl, F *__restrict__ sr) {
    const size_t half = m / 2;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
        const size_t i = idx & (m - 1);
        const F wi = load_fe<F>(w + idx);
        if (i >= half) { store_fe(sl + idx, wi * load_fe<F>(a + (i - half))); store_fe(sr + idx, F::zero()); }
        else { store_fe(sr + idx, wi * load_fe<F>(a + (i + half))); store_fe(sl + idx, F::zero()); }
    }
}
// W_{j+1}[idx] = W_j[idx] * (idx mod m >= m/2 ? r : 1/r)
template <class F>
__global__ void __launch_bounds__(256) ipa_weights_update_kernel(F *w, size_t n, size_t m, F r, F r_inv) {
    const size_t half = m / 2;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x)
        store_fe(w + idx, load_fe<F>(w + idx) * (((idx & (m - 1)) >= half) ? r : r_inv));
}
template <class F>
__global__ void __launch_bounds__(256) fill_one_kernel(F *w, size_t n) {
    const F one = F::one();
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) store_fe(w + idx, one);
}

template <class C>
static int ipa_prove(lurk_msm_ctx *ck, const uint8_t *gc_bytes, void *d_a, void *d_b, int log_n, lurk_challenge_fn challenge, void *user,
                     uint8_t *L_out, uint8_t *R_out, uint8_t *a_final, uint8_t *b_final, int fmt, cudaStream_t s) {
    using Fb = typename C::Base;
    using Fs = typename C::Scalar;
    Affine<Fb> gc;
    if (!fe_in(gc_bytes, fmt, gc.x) || !fe_in(gc_bytes + 32, fmt, gc.y)) { set_error("ck_c is not reduced"); return LURK_ERR_RANGE; }
    const HostFixedBase<Fb> gc_mul(gc);
    ScScratch<Fs> sc;
    LURK_TRY(sc.init(s));
    Fs *a = static_cast<Fs *>(d_a), *b = static_cast<Fs *>(d_b);
    const size_t n = (size_t)1 << log_n;
    DevBuf wbuf;
    LURK_TRY(wbuf.alloc(3 * n * sizeof(Fs)));
    Fs *W = wbuf.as<Fs>(), *sl = W + n, *sr = W + 2 * n;
    fill_one_kernel<Fs><<<sc_grid(n, 256), 256, 0, s>>>(W, n);
    LURK_CUDA_TRY(cudaGetLastError());
    // L and R of a round are independent: the second one runs on a clone of the context (same resident key, own scratch) and a side stream
    MsmCloneGuard ck_r;
    LURK_TRY(lurk_msm_ctx_clone(ck, &ck_r.c));
    StreamGuard s_r;
    LURK_TRY(s_r.create());
    EventGuard weighted;
    LURK_TRY(weighted.create());
    size_t m = n;
    for (int round = 0; round < log_n; round++) {
        const size_t half = m / 2;
        Fs cl, cr;
        LURK_TRY(dot_dev<Fs>(a, b + half, half, &cl, sc, s));
        LURK_TRY(dot_dev<Fs>(a + half, b, half, &cr, sc, s));
        ipa_weighted_kernel<Fs><<<sc_grid(n, 256), 256, 0, s>>>(W, a, n, m, sl, sr);
        LURK_CUDA_TRY(cudaGetLastError());
        LURK_CUDA_TRY(cudaEventRecord(weighted.e, s));
        LURK_CUDA_TRY(cudaStreamWaitEvent(s_r.s, weighted.e, 0));
        uint8_t parts[2][96];
        LURK_TRY(lurk_msm_ctx_launch_dev(ck, sl, n, LURK_FMT_MONTGOMERY, s));
        LURK_TRY(lurk_msm_ctx_launch_dev(ck_r.c, sr, n, LURK_FMT_MONTGOMERY, s_r.s));
        LURK_TRY(lurk_msm_ctx_finish(ck, parts[0]));
        LURK_TRY(lurk_msm_ctx_finish(ck_r.c, parts[1]));        // both passes are complete before anything below touches sl / sr / a
        uint8_t lr[192];
        for (int side = 0; side < 2; side++) {
            // L = <a_lo, G_hi> + c_L ck_c,  R = <a_hi, G_lo> + c_R ck_c  (G = the folded key of this round, never materialised)
            const uint8_t *part = parts[side];
            XYZZ<Fb> acc = XYZZ<Fb>::identity();
            Fb z;
            memcpy(z.v, part + 64, 32);
            if (!z.is_zero()) { Affine<Fb> p; memcpy(p.x.v, part, 32); memcpy(p.y.v, part + 32, 32); acc.add_affine(p); }
            const Fs c = (side == 0 ? cl : cr).to_canonical();
            acc.add(gc_mul.mul(c.v));
            point_to_bytes_fmt(acc, fmt, lr + 96 * side);
        }
        if (L_out) memcpy(L_out + 96 * (size_t)round, lr, 96);
        if (R_out) memcpy(R_out + 96 * (size_t)round, lr + 96, 96);
        uint8_t rbytes[32];
        int rc = challenge(user, round, lr, 192, rbytes);
        if (rc != 0) { set_error("challenge callback failed in round %d (%d)", round, rc); return LURK_ERR_ARG; }
        Fs r;
        if (!fe_in(rbytes, fmt, r) || r.is_zero()) { set_error("challenge of round %d is zero or not reduced", round); return LURK_ERR_RANGE; }
        const Fs r_inv = r.inv();
        // a' = a_lo r + a_hi r^-1;  b' = b_lo r^-1 + b_hi r;  key weights: low half r^-1, high half r
        ipa_fold_scalars_kernel<Fs><<<sc_grid(half, 256), 256, 0, s>>>(a, half, r, r_inv);
        ipa_fold_scalars_kernel<Fs><<<sc_grid(half, 256), 256, 0, s>>>(b, half, r_inv, r);
        ipa_weights_update_kernel<Fs><<<sc_grid(n, 256), 256, 0, s>>>(W, n, m, r, r_inv);
        LURK_CUDA_TRY(cudaGetLastError());
        m = half;
    }
    Fs fin[2];
    LURK_CUDA_TRY(cudaMemcpyAsync(&fin[0], a, sizeof(Fs), cudaMemcpyDeviceToHost, s));
    LURK_CUDA_TRY(cudaMemcpyAsync(&fin[1], b, sizeof(Fs), cudaMemcpyDeviceToHost, s));
    LURK_CUDA_TRY(cudaStreamSynchronize(s));
    if (a_final) fe_out(fin[0], fmt, a_final);
    if (b_final) fe_out(fin[1], fmt, b_final);
    return LURK_OK;
}

}  // namespace lurk

using namespace lurk;

extern "C" {

int lurk_ipa_fold_scalars_dev(int field_id, void *d_a, size_t n, const uint8_t x[32], const uint8_t y[32], int fmt, void *stream) {
    if (!d_a || !x || !y || n < 2 || (n & (n - 1))) { set_error("bad argument (n must be a power of two >= 2)"); return LURK_ERR_ARG; }
    if (fmt != LURK_FMT_CANONICAL && fmt != LURK_FMT_MONTGOMERY) { set_error("bad format %d", fmt); return LURK_ERR_ARG; }
    LURK_TRY(require_gpu());
    return dispatch_field(field_id, [&](auto f) {
        using F = decltype(f);
        F fx, fy;
        if (!fe_in(x, fmt, fx) || !fe_in(y, fmt, fy)) { set_error("scalar is not reduced"); return LURK_ERR_RANGE; }
        ipa_fold_scalars_kernel<F><<<sc_grid(n / 2, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<F *>(d_a), n / 2, fx, fy);
        LURK_CUDA_TRY(cudaGetLastError());
        return LURK_OK;
    });
}

int lurk_ipa_fold_bases_dev(int curve_id, void *d_bases_mont, size_t n, const uint8_t x[32], const uint8_t y[32], int fmt, void *stream) {
    if (!d_bases_mont || !x || !y || n < 2 || (n & (n - 1))) { set_error("bad argument (n must be a power of two >= 2)"); return LURK_ERR_ARG; }
    if (fmt != LURK_FMT_CANONICAL && fmt != LURK_FMT_MONTGOMERY) { set_error("bad format %d", fmt); return LURK_ERR_ARG; }
    LURK_TRY(require_gpu());
    return dispatch_curve(curve_id, [&](auto c) {
        using C = decltype(c);
        using Fs = typename C::Scalar;
        using Fb = typename C::Base;
        Fs fx, fy;
        if (!fe_in(x, fmt, fx) || !fe_in(y, fmt, fy)) { set_error("scalar is not reduced"); return LURK_ERR_RANGE; }
        Scalar256 sx, sy;
        const Fs cx = fx.to_canonical(), cy = fy.to_canonical();
        for (int i = 0; i < 8; i++) { sx.w[i] = cx.v[i]; sy.w[i] = cy.v[i]; }
        ipa_fold_bases_kernel<Fb><<<sc_grid(n / 2, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<Affine<Fb> *>(d_bases_mont), n / 2, sx, sy);
        LURK_CUDA_TRY(cudaGetLastError());
        return LURK_OK;
    });
}

int lurk_ipa_prove_dev(int curve_id, lurk_msm_ctx *ck, const uint8_t ck_c[64], void *d_a, void *d_b, int log_n, lurk_challenge_fn challenge,
                       void *user, uint8_t *L_out, uint8_t *R_out, uint8_t a_final[32], uint8_t b_final[32], int fmt, void *stream) {
    if (!ck || !ck_c || !d_a || !d_b || !challenge) { set_error("null argument"); return LURK_ERR_ARG; }
    if (log_n < 0 || log_n > 30) { set_error("bad log_n %d", log_n); return LURK_ERR_ARG; }
    if (fmt != LURK_FMT_CANONICAL && fmt != LURK_FMT_MONTGOMERY) { set_error("bad format %d", fmt); return LURK_ERR_ARG; }
    LURK_TRY(require_gpu());
    int ck_curve = -1;
    size_t ck_n = 0;
    LURK_TRY(lurk_msm_ctx_info(ck, &ck_curve, &ck_n));
    if (ck_curve != curve_id || ck_n < ((size_t)1 << log_n)) {
        set_error("commitment key: curve %d with %zu bases, need curve %d with >= 2^%d", ck_curve, ck_n, curve_id, log_n);
        return LURK_ERR_ARG;
    }
    return dispatch_curve(curve_id, [&](auto c) {
        return ipa_prove<decltype(c)>(ck, ck_c, d_a, d_b, log_n, challenge, user, L_out, R_out, a_final, b_final, fmt,
                                      static_cast<cudaStream_t>(stream));
    });
}

}  // extern "C"
