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
//   dot_kernel (sc_scratch.cuh)  inner product (MultilinearPolynomial::evaluate = <Z, eq(r)>, IPA's c_L / c_R).
// The inner-product argument lives in ipa.cu.
// Reductions: per-thread modular sums -> warp shuffles -> shared memory -> one partial per CTA -> the last CTA to finish adds the
// partials (single launch, no second kernel, no atomics on field elements).
#include "common.cuh"
#include "sumcheck.cuh"
#include "reduce.cuh"
#include "sc_scratch.cuh"

#include <algorithm>
#include <vector>

namespace lurk {

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

// ------------------------------------------------------------------------------------------------ host-side helpers
// SumcheckProof::prove_quad_batch / prove_cubic_with_additive_term_batch (the BatchedRelaxedR1CSSNARK of SuperNova's `compress`,
// reference src/proof/supernova.rs:293-317) -- and, with one instance and coefficient 1, the plain prove_quad /
// prove_cubic_with_additive_term.  Instance i has its own polynomials of 2^nr[i] elements and joins in round max - nr[i]; until then
// its round polynomial is the constant 2^(remaining - nr[i] - 1) claim_i.  The round message is sum_i coeff_i s_i(X).
constexpr int SC_MAX_INSTANCES = 60;
template <class F, int KIND>
static int sumcheck_prove_batch(int n_inst, void *const *d_polys, const int *nr, const uint8_t *claims_in, const uint8_t *coeffs_in,
                                lurk_challenge_fn challenge, void *user, uint8_t *round_evals, uint8_t *challenges, uint8_t *final_evals, int fmt,
                                cudaStream_t s) {
    constexpr int K = ScShape<KIND>::POLYS, E = ScShape<KIND>::EVALS, DEG1 = E + 1;
    std::vector<F> claim(n_inst), coeff(n_inst);
    int max_rounds = 0;
    for (int i = 0; i < n_inst; i++) {
        if (!fe_in(claims_in + 32 * i, fmt, claim[i])) { set_error("claim %d is not reduced", i); return LURK_ERR_RANGE; }
        if (coeffs_in) { if (!fe_in(coeffs_in + 32 * i, fmt, coeff[i])) { set_error("coefficient %d is not reduced", i); return LURK_ERR_RANGE; } }
        else coeff[i] = F::one();
        max_rounds = std::max(max_rounds, nr[i]);
    }
    ScScratch<F> sc;
    LURK_TRY(sc.init(s));
    std::vector<ScArgs<F>> args(n_inst);
    std::vector<size_t> cur(n_inst);
    for (int i = 0; i < n_inst; i++) {
        memset(&args[i], 0, sizeof(ScArgs<F>));
        for (int k = 0; k < K; k++) args[i].poly[k] = static_cast<F *>(d_polys[i * K + k]);
        args[i].partial = sc.partial; args[i].counter = sc.counter; args[i].result = sc.result + 4 * i;
        args[i].r = F::zero();
        cur[i] = (size_t)1 << nr[i];
    }
    const F two = F::from_u64(2);
    auto pow2 = [&](int k) { F r = F::one(); for (int j = 0; j < k; j++) r = r * two; return r; };
    F e = F::zero();
    for (int i = 0; i < n_inst; i++) e += coeff[i] * claim[i] * pow2(max_rounds - nr[i]);
    F r_prev = F::zero();
    const ScLagrange<F> lagrange(DEG1);
    for (int round = 0; round < max_rounds; round++) {
        const int remaining = max_rounds - round;
        for (int i = 0; i < n_inst; i++) {
            if (remaining > nr[i]) continue;
            ScArgs<F> &a = args[i];
            const int grid = sc_grid(cur[i] / 2, 256);
            if (remaining == nr[i]) {                  // the instance's first round: evaluate only
                a.len = cur[i];
                sc_round_kernel<F, KIND, false><<<grid, 256, 0, s>>>(a);
            } else {                                   // bind the previous challenge, then evaluate
                a.len = cur[i] << 1;
                a.r = r_prev;
                sc_round_kernel<F, KIND, true><<<grid, 256, 0, s>>>(a);
            }
        }
        LURK_CUDA_TRY(cudaGetLastError());
        LURK_CUDA_TRY(cudaStreamSynchronize(s));       // the result slots are pinned host memory
        F comb[E];
        for (int t = 0; t < E; t++) comb[t] = F::zero();
        for (int i = 0; i < n_inst; i++) {
            if (remaining > nr[i]) {
                const F c = coeff[i] * claim[i] * pow2(remaining - nr[i] - 1);
                for (int t = 0; t < E; t++) comb[t] += c;
            } else {
                const F *res = static_cast<const F *>(sc.pinned) + 4 * i;
                for (int t = 0; t < E; t++) comb[t] += coeff[i] * res[t];
            }
        }
        // s(0), s(1) = claim - s(0), s(2)[, s(3)]
        F evals[DEG1];
        evals[0] = comb[0];
        evals[1] = e - comb[0];
        for (int t = 1; t < E; t++) evals[t + 1] = comb[t];
        uint8_t msg[DEG1 * 32], rbytes[32];
        for (int t = 0; t < DEG1; t++) fe_out(evals[t], fmt, msg + 32 * t);
        if (round_evals) memcpy(round_evals + (size_t)round * DEG1 * 32, msg, DEG1 * 32);
        int rc = challenge(user, round, msg, DEG1 * 32, rbytes);
        if (rc != 0) { set_error("challenge callback failed in round %d (%d)", round, rc); return LURK_ERR_ARG; }
        F r;
        if (!fe_in(rbytes, fmt, r)) { set_error("challenge of round %d is not reduced", round); return LURK_ERR_RANGE; }
        if (challenges) memcpy(challenges + (size_t)round * 32, rbytes, 32);
        e = lagrange.eval(evals, r);
        r_prev = r;
        for (int i = 0; i < n_inst; i++)
            if (remaining <= nr[i]) cur[i] >>= 1;
    }
    // final evaluations: the last bind of every instance that took part; instances without variables are their single element
    for (int i = 0; i < n_inst; i++) {
        if (nr[i] == 0) {
            for (int k = 0; k < K; k++) LURK_CUDA_TRY(cudaMemcpyAsync(sc.result + 4 * i + k, args[i].poly[k], sizeof(F), cudaMemcpyDeviceToHost, s));
        } else {
            args[i].len = 2;
            args[i].r = r_prev;
            sc_final_bind_kernel<F><<<1, 32, 0, s>>>(args[i], K);
        }
    }
    LURK_CUDA_TRY(cudaGetLastError());
    LURK_CUDA_TRY(cudaStreamSynchronize(s));
    if (final_evals)
        for (int i = 0; i < n_inst; i++)
            for (int k = 0; k < K; k++) fe_out(static_cast<const F *>(sc.pinned)[4 * i + k], fmt, final_evals + 32 * (i * K + k));
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

}  // namespace lurk

using namespace lurk;

extern "C" {

int lurk_sumcheck_prove_batch_dev(int field_id, int kind, int n_instances, void *const *d_polys, const int *num_rounds, const uint8_t *claims,
                                  const uint8_t *coeffs, lurk_challenge_fn challenge, void *user, uint8_t *round_evals, uint8_t *challenges,
                                  uint8_t *final_evals, int fmt, void *stream) {
    if (!d_polys || !claims || !challenge || !num_rounds) { set_error("null argument"); return LURK_ERR_ARG; }
    if (kind != LURK_SUMCHECK_QUAD && kind != LURK_SUMCHECK_CUBIC) { set_error("unknown sum-check kind %d", kind); return LURK_ERR_ARG; }
    if (n_instances < 1 || n_instances > SC_MAX_INSTANCES) { set_error("1..%d instances, got %d", SC_MAX_INSTANCES, n_instances); return LURK_ERR_ARG; }
    if (fmt != LURK_FMT_CANONICAL && fmt != LURK_FMT_MONTGOMERY) { set_error("bad format %d", fmt); return LURK_ERR_ARG; }
    const int k = kind == LURK_SUMCHECK_QUAD ? 2 : 4;
    for (int i = 0; i < n_instances; i++) {
        if (num_rounds[i] < 0 || num_rounds[i] > 40) { set_error("bad number of rounds %d (instance %d)", num_rounds[i], i); return LURK_ERR_ARG; }
        for (int j = 0; j < k; j++)
            if (!d_polys[i * k + j]) { set_error("polynomial %d of instance %d is null", j, i); return LURK_ERR_ARG; }
    }
    LURK_TRY(require_gpu());
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    return dispatch_field(field_id, [&](auto f) {
        using F = decltype(f);
        return kind == LURK_SUMCHECK_QUAD
                   ? sumcheck_prove_batch<F, SC_QUAD>(n_instances, d_polys, num_rounds, claims, coeffs, challenge, user, round_evals, challenges, final_evals, fmt, s)
                   : sumcheck_prove_batch<F, SC_CUBIC>(n_instances, d_polys, num_rounds, claims, coeffs, challenge, user, round_evals, challenges, final_evals, fmt, s);
    });
}

int lurk_sumcheck_prove_dev(int field_id, int kind, void *const *d_polys, int num_rounds, const uint8_t claim[32], lurk_challenge_fn challenge,
                            void *user, uint8_t *round_evals, uint8_t *challenges, uint8_t *final_evals, int fmt, void *stream) {
    return lurk_sumcheck_prove_batch_dev(field_id, kind, 1, d_polys, &num_rounds, claim, nullptr, challenge, user, round_evals, challenges, final_evals,
                                         fmt, stream);
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

}  // extern "C"
