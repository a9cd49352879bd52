// Element / segment level arithmetic of the KZG side of the BN256 path (SURVEY.md 8(f) N3 + N4):
//   * N3: powers-of-tau key  g, beta g, beta^2 g, ...  (Arecibo hyperkzg CommitmentKey::setup -> UniversalKZGParam::gen_srs_for_testing,
//     reached from public_params for Bn256EngineKZG -- reference src/proof/nova.rs:65-71, 196-216) by fixed-base windows of g;
//   * N4: the prover loops of provider::hyperkzg::EvaluationEngine::prove (reached from `compress`, src/proof/nova.rs:341-356):
//     Pi+1[j] = Pi[2j] + x (Pi[2j+1] - Pi[2j]);  v = P_j(u_i);  B = sum_j q^j P_j;  h[i-1] = B[i] + u h[i]  (witness polynomial).
// The witness polynomial is a first-order linear recurrence; it and the evaluations are computed by a segmented scheme: with
// H_X(i) = sum_{k >= i} X[k] v^(k-i), the aggregates Y[s] = sum_{k in segment s} X[k] v^(k - sL) satisfy H_X(sL) = H_Y(s) with
// multiplier v^L -- so an up-sweep (aggregates of aggregates) and a down-sweep (each segment re-runs the recurrence from the value
// just above it) give every H_X(i) in O(n / L) parallel segments per level.  LURK_HD: the CPU suite runs the same functions.
#pragma once
#include "curve.cuh"

namespace lurk {

constexpr int KZG_SEG = 32;         // elements per segment (= per thread)
constexpr int KZG_WINDOW_BITS = 8;  // fixed-base windows of the powers-of-tau generator
constexpr int KZG_WINDOWS = 32;

// sum_{k in [lo, hi)} X[k] v^(k - lo)
template <class F>
LURK_HD F kzg_seg_horner(const F *X, size_t lo, size_t hi, const F &v) {
    F r = F::zero();
    for (size_t k = hi; k-- > lo;) r = load_fe<F>(X + k) + v * r;
    return r;
}
// r = carry (= H_X(hi)); for k = hi-1 .. lo: r = X[k] + v r = H_X(k); out[k - shift] = r  (k < shift is not written)
template <class F>
LURK_HD void kzg_seg_down(const F *X, size_t lo, size_t hi, const F &v, F carry, F *out, size_t shift) {
    F r = carry;
    for (size_t k = hi; k-- > lo;) {
        r = load_fe<F>(X + k) + v * r;
        if (k >= shift) store_fe(out + (k - shift), r);
    }
}
// x^e for a small exponent
template <class F>
LURK_HD F kzg_pow_small(const F &x, uint64_t e) {
    F acc = F::one(), base = x;
    while (e) {
        if (e & 1) acc = acc * base;
        e >>= 1;
        if (e) base = base.sqr();
    }
    return acc;
}
// Pi+1[j] = Pi[2j] + x (Pi[2j+1] - Pi[2j])
template <class F>
LURK_HD F kzg_fold_low(const F &even, const F &odd, const F &x) { return even + x * (odd - even); }

// offset / length of polynomial j in the concatenated buffer P_0 | P_1 | ... (len_j = n >> j)
LURK_HD size_t kzg_poly_offset(size_t n, int j) { return j == 0 ? 0 : 2 * n - (n >> (j - 1)); }

// [s] g from the window table: table[w * 255 + (d - 1)] = d 2^(8w) g (affine, Montgomery); s canonical
template <class F>
LURK_HD XYZZ<F> kzg_fixed_base_mul(const Affine<F> *table, const uint32_t s[8]) {
    XYZZ<F> acc = XYZZ<F>::identity();
    for (int w = 0; w < KZG_WINDOWS; w++) {
        const uint32_t d = (s[w >> 2] >> (8 * (w & 3))) & 0xffu;
        if (d) {
            Affine<F> t;
            t.x = load_fe<F>(&table[w * 255 + (d - 1)].x);
            t.y = load_fe<F>(&table[w * 255 + (d - 1)].y);
            acc.add_affine(t);
        }
    }
    return acc;
}

}  // namespace lurk
