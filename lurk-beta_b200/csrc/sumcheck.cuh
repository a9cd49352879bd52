// Element-level arithmetic of the sum-check prover and of the inner-product argument's folding step:
// what one index i contributes in one round of Arecibo's SumcheckProof::prove_quad / prove_cubic_with_additive_term
// (compute_eval_points_* + bind_poly_var_top) and in one round of InnerProductArgument::prove (reached from `compress`,
// reference src/proof/nova.rs:341-356).  LURK_HD: the CPU test-suite runs these exact templates against oracle/sumcheck.py;
// sumcheck.cu wraps them in the grid-wide kernels.
#pragma once
#include "curve.cuh"

namespace lurk {

constexpr int SC_QUAD = 0;    // sum_i A[i] B[i]                      -- 2 polynomials, degree 2: s(0), s(2)
constexpr int SC_CUBIC = 1;   // sum_i A[i] (B[i] C[i] - D[i])        -- 4 polynomials, degree 3: s(0), s(2), s(3)
template <int KIND> struct ScShape { static constexpr int POLYS = KIND == SC_QUAD ? 2 : 4; static constexpr int EVALS = KIND == SC_QUAD ? 2 : 3; };

// bind_poly_var_top for one index: lo + r (hi - lo)
template <class F>
LURK_HD F sc_bind(const F &lo, const F &hi, const F &r) { return lo + r * (hi - lo); }

// adds this index's contribution to the round polynomial's values at 0, 2 (and 3); lo = P[i], hi = P[i + n/2]
template <class F, int KIND>
LURK_HD void sc_accumulate(const F *lo, const F *hi, F *acc) {
    constexpr int K = ScShape<KIND>::POLYS;
    F p2[K], p3[K];
#pragma unroll
    for (int k = 0; k < K; k++) {
        const F d = hi[k] - lo[k];
        p2[k] = hi[k] + d;          // 2 hi - lo
        p3[k] = p2[k] + d;          // 3 hi - 2 lo
    }
    if (KIND == SC_QUAD) {
        acc[0] += lo[0] * lo[1];
        acc[1] += p2[0] * p2[1];
    } else {
        acc[0] += lo[0] * (lo[1] * lo[2] - lo[3]);
        acc[1] += p2[0] * (p2[1] * p2[2] - p2[3]);
        acc[2] += p3[0] * (p3[1] * p3[2] - p3[3]);
    }
}

// value at x of the polynomial of degree n - 1 through (0, e[0]), (1, e[1]), ... (n <= 4): UniPoly::from_evals + evaluate.
// The Lagrange denominators prod_{j != i} (i - j) are constants of n: ScLagrange inverts them ONCE per prover call -- a Fermat
// inversion is ~380 host products, and four of them per round were half of a round's host time.
template <class F>
struct ScLagrange {
    int n;
    F inv_den[4];
    explicit ScLagrange(int n_) : n(n_) {
        for (int i = 0; i < n; i++) {
            F den = F::one();
            for (int j = 0; j < n; j++) {
                if (j == i) continue;
                den = den * (i > j ? F::from_u64((uint64_t)(i - j)) : F::from_u64((uint64_t)(j - i)).neg());
            }
            inv_den[i] = den.inv();
        }
    }
    F eval(const F *e, const F &x) const {
        F total = F::zero();
        for (int i = 0; i < n; i++) {
            F num = F::one();
            for (int j = 0; j < n; j++)
                if (j != i) num = num * (x - F::from_u64((uint64_t)j));
            total += e[i] * num * inv_den[i];
        }
        return total;
    }
};
template <class F>
inline F sc_interpolate(const F *e, int n, const F &x) { return ScLagrange<F>(n).eval(e, x); }

// IPA scalar fold: x a[i] + y a[i + n/2] with one reduction
template <class F>
LURK_HD F ipa_fold_scalar(const F &lo, const F &hi, const F &x, const F &y) {
    WideAcc<typename F::Params> acc;
    acc.clear();
    acc.mul_acc(lo, x);
    acc.mul_acc(hi, y);
    return acc.reduce();
}

// IPA key fold: x P + y Q for two affine points and two scalars given as canonical 256-bit integers (the same for every thread:
// the branch pattern of the interleaved double-and-add is uniform across a warp).  Result affine, identity = (0, 0).
template <class F>
LURK_HD Affine<F> ipa_fold_point(const Affine<F> &P, const Affine<F> &Q, const uint32_t x[8], const uint32_t y[8]) {
    XYZZ<F> pq = XYZZ<F>::from_affine(P);
    pq.add_affine(Q);
    XYZZ<F> acc = XYZZ<F>::identity();
    int top = 255;
    while (top > 0 && !(((x[top >> 5] | y[top >> 5]) >> (top & 31)) & 1)) top--;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int i = top; i >= 0; i--) {
        acc = acc.dbl();
        const uint32_t bx = (x[i >> 5] >> (i & 31)) & 1, by = (y[i >> 5] >> (i & 31)) & 1;
        if (bx & by) acc.add(pq);
        else if (bx) acc.add_affine(P);
        else if (by) acc.add_affine(Q);
    }
    return acc.to_affine();
}

}  // namespace lurk
