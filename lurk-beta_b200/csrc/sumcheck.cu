// N4 -- the data-parallel half of `compress` behind the C ABI (include/lurk_b200.h, "N4"): sum-check prover rounds over
// device-resident multilinear polynomials and the folding rounds of the inner-product argument, i.e. the loops of Arecibo's
// SumcheckProof::prove_quad / prove_cubic_with_additive_term and InnerProductArgument::prove as RelaxedR1CSSNARK::prove runs
// them (reached from reference src/proof/nova.rs:341-356, supernova.rs:293-317).  The Fiat-Shamir transcript stays with the
// caller: every round hands its message to a callback and gets the challenge back (32 bytes in, <= 192 bytes out per round).
//
// Kernels (all grid-stride, 256-thread CTAs, grid <= 4 CTAs per SM, 128-bit loads / stores of 32-byte elements):
//   sc_round_kernel<KIND, BIND>  one pass per round: [bind the previous challenge into all K polynomials in place] + this round's
//                                s(0), s(2)[, s(3)] -- the bind of round j and the evaluation of round j + 1 read the same data,
//                                fusing them moves 3 n / 2 elements per polynomial and round instead of 2 n (and halves the launches).
//                                Bytes per index pair: K x (4 x 32 read + 2 x 32 written); products: K x 2 + 2 (quad: 4) / 6 (cubic).
//   eq_kernel                    EqPolynomial::evals: 16 outputs per thread (prefix product over the high bits, doubling over the low 4).
//   dot_kernel                   inner product (MultilinearPolynomial::evaluate = <Z, eq(r)>, IPA's c_L / c_R).
//   ipa_fold_scalars / _bases    a' = x a_lo + y a_hi;  G' = x G_lo + y G_hi (interleaved double-and-add, uniform branches).
// Reductions: per-thread modular sums -> warp shuffles -> shared memory -> one partial per CTA -> the last CTA to finish adds the
// partials (single launch, no second kernel, no atomics on field elements).
#include "common.cuh"
#include "sumcheck.cuh"
#include "reduce.cuh"

#include <algorithm>
#include <vector>

namespace lurk {

static inline int sc_grid(size_t n, int block) {
    size_t want = (n + block - 1) / block;
    size_t cap = (size_t)sm_count() * 4;
    return (int)(want < cap ? (want ? want : 1) : cap);
}

// ------------------------------------------------------------------------------------------------ sum-check round
template <class F>
struct ScArgs {
    F *poly[4];
    size_t len;          // length of every polynomial on entry
    F r;                 // BIND: the previous round's challenge (Montgomery)
    F *partial;          // grid x EVALS
    unsigned *counter;
    F *result;           // EVALS
};

template <class F, int KIND, bool BIND>
__global__ void __launch_bounds__(256, 2) sc_round_kernel(const __grid_constant__ ScArgs<F> a) {
    constexpr int K = ScShape<KIND>::POLYS, E = ScShape<KIND>::EVALS;
    F acc[E];
#pragma unroll
    for (int e = 0; e < E; e++) acc[e] = F::zero();
    const size_t half = BIND ? a.len / 4 : a.len / 2;     // index pairs of THIS round
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < half; i += (size_t)gridDim.x * blockDim.x) {
        F lo[K], hi[K];
#pragma unroll
        for (int k = 0; k < K; k++) {
            if (BIND) {
                const F l0 = load_fe<F>(a.poly[k] + i), l1 = load_fe<F>(a.poly[k] + i + a.len / 2);
                const F h0 = load_fe<F>(a.poly[k] + i + a.len / 4), h1 = load_fe<F>(a.poly[k] + i + a.len / 4 + a.len / 2);
                lo[k] = sc_bind(l0, l1, a.r);
                hi[k] = sc_bind(h0, h1, a.r);
                store_fe(a.poly[k] + i, lo[k]);                 // only this thread ever touches these four slots
                store_fe(a.poly[k] + i + a.len / 4, hi[k]);
            } else {
                lo[k] = load_fe<F>(a.poly[k] + i);
                hi[k] = load_fe<F>(a.poly[k] + i + half);
            }
        }
        sc_accumulate<F, KIND>(lo, hi, acc);
    }
    grid_sum<F, E>(acc, a.partial, a.counter, a.result);
}

// the last bind (length 2 -> 1): the final evaluations of the K polynomials
template <class F>
__global__ void sc_final_bind_kernel(const __grid_constant__ ScArgs<F> a, int k_polys) {
    if (threadIdx.x < k_polys && blockIdx.x == 0) {
        F v = sc_bind(load_fe<F>(a.poly[threadIdx.x]), load_fe<F>(a.poly[threadIdx.x] + 1), a.r);
        store_fe(a.poly[threadIdx.x], v);
        store_fe(&a.result[threadIdx.x], v);
    }
}

template <class F>
__global__ void __launch_bounds__(256) dot_kernel(const F *__restrict__ x, const F *__restrict__ y, size_t n, F *partial, unsigned *counter, F *result) {
    F acc[1] = {F::zero()};
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        acc[0] += load_fe<F>(x + i) * load_fe<F>(y + i);
    grid_sum<F, 1>(acc, partial, counter, result);
}

// ------------------------------------------------------------------------------------------------ eq table
template <class F>
struct EqArgs { F tau[32], one_minus[32]; int l; };

template <class F, int LOW>
__global__ void __launch_bounds__(128) eq_kernel(const __grid_constant__ EqArgs<F> a, F *__restrict__ out, int to_canonical) {
    const int high = a.l - LOW;
    const size_t groups = (size_t)1 << high;
    for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (size_t)gridDim.x * blockDim.x) {
        F vals[1 << LOW];
        F v = F::one();
        for (int j = 0; j < high; j++) v = v * (((g >> (high - 1 - j)) & 1) ? a.tau[j] : a.one_minus[j]);
        vals[0] = v;
#pragma unroll
        for (int j = 0; j < LOW; j++) {
#pragma unroll
            for (int t = (1 << j) - 1; t >= 0; t--) {
                const F hi = vals[t] * a.tau[high + j];
                vals[2 * t + 1] = hi;
                vals[2 * t] = vals[t] - hi;
            }
        }
#pragma unroll
        for (int t = 0; t < (1 << LOW); t++) store_fe(out + (g << LOW) + t, to_canonical ? vals[t].to_canonical() : vals[t]);
    }
}

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

// ------------------------------------------------------------------------------------------------ host-side helpers
// Scratch of the reductions: per-CTA partials, the ticket counter, the result slots and their pinned mirror.  One per host thread
// and device, kept for the life of the thread: allocating (and above all freeing) device / pinned memory inside every call would
// synchronise the whole device each time.  Every field element is 32 bytes, so the pool is type-agnostic.
struct ScPool {
    void *dev = nullptr, *pinned = nullptr;
    int device = -1;
    ~ScPool() { if (dev) cudaFree(dev); if (pinned) cudaFreeHost(pinned); }
};
static ScPool &sc_pool() {
    static thread_local ScPool pool;
    return pool;
}
template <class F>
struct ScScratch {
    F *partial = nullptr, *result = nullptr;
    unsigned *counter = nullptr;
    void *pinned = nullptr;
    int init(cudaStream_t s) {
        ScPool &pool = sc_pool();
        const size_t cap = (size_t)sm_count() * 4;
        int dev = -1;
        LURK_CUDA_TRY(cudaGetDevice(&dev));
        if (pool.device != dev) {
            if (pool.dev) { cudaFree(pool.dev); pool.dev = nullptr; }
            if (pool.pinned) { cudaFreeHost(pool.pinned); pool.pinned = nullptr; }
            LURK_CUDA_TRY(cudaMalloc(&pool.dev, 32 * (cap * 3 + 8) + 64));
            LURK_CUDA_TRY(cudaHostAlloc(&pool.pinned, 32 * 8, cudaHostAllocDefault));
            pool.device = dev;
        }
        partial = static_cast<F *>(pool.dev);
        counter = reinterpret_cast<unsigned *>(partial + cap * 3 + 8);
        pinned = pool.pinned;
        // the result slots ARE the pinned host buffer (unified addressing: the last CTA stores <= 128 bytes across PCIe), so a round
        // costs one launch + one stream synchronisation and no copy
        result = static_cast<F *>(pool.pinned);
        LURK_CUDA_TRY(cudaMemsetAsync(counter, 0, 64, s));     // a kernel that died mid-way must not poison the next call
        return LURK_OK;
    }
    // waits for the kernel that wrote result[0..k)
    int fetch(int k, F *out, cudaStream_t s) {
        LURK_CUDA_TRY(cudaStreamSynchronize(s));
        memcpy(out, pinned, sizeof(F) * k);
        return LURK_OK;
    }
};

template <class F>
static inline void fe_out(const F &x_mont, int fmt, uint8_t *out) {
    F v = fmt == LURK_FMT_CANONICAL ? x_mont.to_canonical() : x_mont;
    memcpy(out, v.v, 32);
}
template <class F>
static inline bool fe_in(const uint8_t *in, int fmt, F &x_mont) {
    F v;
    memcpy(v.v, in, 32);
    if (!v.is_reduced()) return false;
    x_mont = fmt == LURK_FMT_CANONICAL ? F::from_canonical(v) : v;
    return true;
}

template <class F, int KIND>
static int sumcheck_prove(void *const *d_polys, int num_rounds, const uint8_t *claim_in, lurk_challenge_fn challenge, void *user,
                          uint8_t *round_evals, uint8_t *challenges, uint8_t *final_evals, int fmt, cudaStream_t s) {
    constexpr int K = ScShape<KIND>::POLYS, E = ScShape<KIND>::EVALS, DEG1 = E + 1;
    F claim;
    if (!fe_in(claim_in, fmt, claim)) { set_error("claim is not reduced"); return LURK_ERR_RANGE; }
    ScScratch<F> sc;
    LURK_TRY(sc.init(s));
    ScArgs<F> a;
    memset(&a, 0, sizeof a);
    for (int k = 0; k < K; k++) a.poly[k] = static_cast<F *>(d_polys[k]);
    a.partial = sc.partial; a.counter = sc.counter; a.result = sc.result;
    a.r = F::zero();
    size_t len = (size_t)1 << num_rounds;
    for (int round = 0; round < num_rounds; round++) {
        // entry length of this launch: the first launch only evaluates; later ones first bind the previous challenge
        a.len = round == 0 ? len : len << 1;
        const size_t pairs = len / 2;
        const int grid = sc_grid(pairs, 256);
        if (round == 0) sc_round_kernel<F, KIND, false><<<grid, 256, 0, s>>>(a);
        else sc_round_kernel<F, KIND, true><<<grid, 256, 0, s>>>(a);
        LURK_CUDA_TRY(cudaGetLastError());
        F e[E];
        LURK_TRY(sc.fetch(E, e, s));
        // s(0), s(1) = claim - s(0), s(2)[, s(3)]
        F evals[DEG1];
        evals[0] = e[0];
        evals[1] = claim - e[0];
        for (int t = 1; t < E; t++) evals[t + 1] = e[t];
        uint8_t msg[DEG1 * 32], rbytes[32];
        for (int t = 0; t < DEG1; t++) fe_out(evals[t], fmt, msg + 32 * t);
        if (round_evals) memcpy(round_evals + (size_t)round * DEG1 * 32, msg, DEG1 * 32);
        int rc = challenge(user, round, msg, DEG1 * 32, rbytes);
        if (rc != 0) { set_error("challenge callback failed in round %d (%d)", round, rc); return LURK_ERR_ARG; }
        F r;
        if (!fe_in(rbytes, fmt, r)) { set_error("challenge of round %d is not reduced", round); return LURK_ERR_RANGE; }
        if (challenges) memcpy(challenges + (size_t)round * 32, rbytes, 32);
        claim = sc_interpolate(evals, DEG1, r);
        a.r = r;
        len >>= 1;
    }
    F fin[K];
    if (num_rounds == 0) {
        for (int k = 0; k < K; k++) LURK_CUDA_TRY(cudaMemcpyAsync(&fin[k], a.poly[k], sizeof(F), cudaMemcpyDeviceToHost, s));
        LURK_CUDA_TRY(cudaStreamSynchronize(s));
    } else {
        a.len = 2;
        sc_final_bind_kernel<F><<<1, 32, 0, s>>>(a, K);
        LURK_CUDA_TRY(cudaGetLastError());
        LURK_TRY(sc.fetch(K, fin, s));
    }
    if (final_evals)
        for (int k = 0; k < K; k++) fe_out(fin[k], fmt, final_evals + 32 * k);
    return LURK_OK;
}

template <class F>
static int eq_evals(const uint8_t *tau, int l, void *d_out, int fmt, cudaStream_t s) {
    EqArgs<F> a;
    memset(&a, 0, sizeof a);
    a.l = l;
    for (int j = 0; j < l; j++) {
        if (!fe_in(tau + 32 * j, fmt, a.tau[j])) { set_error("tau[%d] is not reduced", j); return LURK_ERR_RANGE; }
        a.one_minus[j] = F::one() - a.tau[j];
    }
    const int to_canonical = fmt == LURK_FMT_CANONICAL;
    F *out = static_cast<F *>(d_out);
    switch (std::min(l, 4)) {
        case 0: eq_kernel<F, 0><<<1, 128, 0, s>>>(a, out, to_canonical); break;
        case 1: eq_kernel<F, 1><<<1, 128, 0, s>>>(a, out, to_canonical); break;
        case 2: eq_kernel<F, 2><<<1, 128, 0, s>>>(a, out, to_canonical); break;
        case 3: eq_kernel<F, 3><<<1, 128, 0, s>>>(a, out, to_canonical); break;
        default: eq_kernel<F, 4><<<sc_grid((size_t)1 << (l - 4), 128), 128, 0, s>>>(a, out, to_canonical); break;
    }
    LURK_CUDA_TRY(cudaGetLastError());
    return LURK_OK;
}

template <class F>
static int dot_dev(const void *d_x, const void *d_y, size_t n, F *out, ScScratch<F> &sc, cudaStream_t s) {
    dot_kernel<F><<<sc_grid(n, 256), 256, 0, s>>>(static_cast<const F *>(d_x), static_cast<const F *>(d_y), n, sc.partial, sc.counter, sc.result);
    LURK_CUDA_TRY(cudaGetLastError());
    return sc.fetch(1, out, s);
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
    size_t m = n;
    for (int round = 0; round < log_n; round++) {
        const size_t half = m / 2;
        Fs cl, cr;
        LURK_TRY(dot_dev<Fs>(a, b + half, half, &cl, sc, s));
        LURK_TRY(dot_dev<Fs>(a + half, b, half, &cr, sc, s));
        ipa_weighted_kernel<Fs><<<sc_grid(n, 256), 256, 0, s>>>(W, a, n, m, sl, sr);
        LURK_CUDA_TRY(cudaGetLastError());
        uint8_t lr[192];
        for (int side = 0; side < 2; side++) {
            // L = <a_lo, G_hi> + c_L ck_c,  R = <a_hi, G_lo> + c_R ck_c  (G = the folded key of this round, never materialised)
            uint8_t part[96];
            LURK_TRY(lurk_msm_ctx_run_dev(ck, side == 0 ? sl : sr, n, LURK_FMT_MONTGOMERY, part, s));
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

int lurk_sumcheck_prove_dev(int field_id, int kind, void *const *d_polys, int num_rounds, const uint8_t claim[32], lurk_challenge_fn challenge,
                            void *user, uint8_t *round_evals, uint8_t *challenges, uint8_t *final_evals, int fmt, void *stream) {
    if (!d_polys || !claim || !challenge) { set_error("null argument"); return LURK_ERR_ARG; }
    if (kind != LURK_SUMCHECK_QUAD && kind != LURK_SUMCHECK_CUBIC) { set_error("unknown sum-check kind %d", kind); return LURK_ERR_ARG; }
    if (num_rounds < 0 || num_rounds > 40) { set_error("bad number of rounds %d", num_rounds); return LURK_ERR_ARG; }
    if (fmt != LURK_FMT_CANONICAL && fmt != LURK_FMT_MONTGOMERY) { set_error("bad format %d", fmt); return LURK_ERR_ARG; }
    for (int k = 0; k < (kind == LURK_SUMCHECK_QUAD ? 2 : 4); k++)
        if (!d_polys[k]) { set_error("polynomial %d is null", k); return LURK_ERR_ARG; }
    LURK_TRY(require_gpu());
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    return dispatch_field(field_id, [&](auto f) {
        using F = decltype(f);
        return kind == LURK_SUMCHECK_QUAD
                   ? sumcheck_prove<F, SC_QUAD>(d_polys, num_rounds, claim, challenge, user, round_evals, challenges, final_evals, fmt, s)
                   : sumcheck_prove<F, SC_CUBIC>(d_polys, num_rounds, claim, challenge, user, round_evals, challenges, final_evals, fmt, s);
    });
}

int lurk_eq_evals_dev(int field_id, const uint8_t *tau, int num_vars, void *d_out, int fmt, void *stream) {
    if ((num_vars && !tau) || !d_out) { set_error("null argument"); return LURK_ERR_ARG; }
    if (num_vars < 0 || num_vars > 32) { set_error("bad number of variables %d", num_vars); return LURK_ERR_ARG; }
    if (fmt != LURK_FMT_CANONICAL && fmt != LURK_FMT_MONTGOMERY) { set_error("bad format %d", fmt); return LURK_ERR_ARG; }
    LURK_TRY(require_gpu());
    return dispatch_field(field_id, [&](auto f) { return eq_evals<decltype(f)>(tau, num_vars, d_out, fmt, static_cast<cudaStream_t>(stream)); });
}

int lurk_inner_product_dev(int field_id, const void *d_a, const void *d_b, size_t n, uint8_t out[32], int fmt, void *stream) {
    if (!out || (n && (!d_a || !d_b))) { set_error("null argument"); return LURK_ERR_ARG; }
    if (fmt != LURK_FMT_CANONICAL && fmt != LURK_FMT_MONTGOMERY) { set_error("bad format %d", fmt); return LURK_ERR_ARG; }
    LURK_TRY(require_gpu());
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    return dispatch_field(field_id, [&](auto f) {
        using F = decltype(f);
        ScScratch<F> sc;
        LURK_TRY(sc.init(s));
        F r = F::zero();
        if (n) LURK_TRY(dot_dev<F>(d_a, d_b, n, &r, sc, s));
        fe_out(r, fmt, out);
        return LURK_OK;
    });
}

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
