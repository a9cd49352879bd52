// Host build of lurk-beta_b200/csrc/sumcheck.cuh: the per-index arithmetic of the sum-check rounds and of the IPA folds, run on the
// CPU (optionally with the GPU limb arithmetic, -DLURK_HOST_EMULATE_CC) against oracle/sumcheck.py.  The kernels in sumcheck.cu
// only add the grid-stride loop and the grid-wide sum around these functions.  Test-only helper.
#include "sumcheck.cuh"
#include <vector>
using namespace lurk;

template <class F>
static F in_fe(const uint8_t *p) { F v; memcpy(v.v, p, 32); return F::from_canonical(v); }
template <class F>
static void out_fe(uint8_t *p, const F &x) { F c = x.to_canonical(); memcpy(p, c.v, 32); }

// one whole round on the host: [bind with r], then the evaluations; polys: k arrays of len canonical elements (updated in place)
template <class F, int KIND>
static int round_host(uint8_t *polys, size_t len, int bind, const uint8_t *r_bytes, uint8_t *evals) {
    constexpr int K = ScShape<KIND>::POLYS, E = ScShape<KIND>::EVALS;
    std::vector<std::vector<F>> P(K, std::vector<F>(len));
    for (int k = 0; k < K; k++)
        for (size_t i = 0; i < len; i++) P[k][i] = in_fe<F>(polys + 32 * (k * len + i));
    size_t cur = len;
    if (bind) {
        F r = in_fe<F>(r_bytes);
        for (int k = 0; k < K; k++)
            for (size_t i = 0; i < len / 2; i++) P[k][i] = sc_bind(P[k][i], P[k][i + len / 2], r);
        cur = len / 2;
    }
    F acc[E];
    for (int e = 0; e < E; e++) acc[e] = F::zero();
    for (size_t i = 0; i < cur / 2; i++) {
        F lo[K], hi[K];
        for (int k = 0; k < K; k++) { lo[k] = P[k][i]; hi[k] = P[k][i + cur / 2]; }
        sc_accumulate<F, KIND>(lo, hi, acc);
    }
    for (int e = 0; e < E; e++) out_fe(evals + 32 * e, acc[e]);
    for (int k = 0; k < K; k++)
        for (size_t i = 0; i < cur; i++) out_fe(polys + 32 * (k * len + i), P[k][i]);
    return 0;
}
extern "C" int sc_test_round(int field, int kind, uint8_t *polys, size_t len, int bind, const uint8_t *r, uint8_t *evals) {
#define CASE(ID, P) case ID: return kind == 0 ? round_host<Fe<P>, SC_QUAD>(polys, len, bind, r, evals) : round_host<Fe<P>, SC_CUBIC>(polys, len, bind, r, evals);
    switch (field) { CASE(0, Bn254Fr) CASE(1, Bn254Fq) CASE(2, PallasFq) CASE(3, PallasFp) }
#undef CASE
    return -3;
}

template <class F>
static int interp(const uint8_t *evals, int n, const uint8_t *x, uint8_t *out) {
    F e[4];
    for (int i = 0; i < n; i++) e[i] = in_fe<F>(evals + 32 * i);
    out_fe(out, sc_interpolate(e, n, in_fe<F>(x)));
    return 0;
}
extern "C" int sc_test_interpolate(int field, const uint8_t *evals, int n, const uint8_t *x, uint8_t *out) {
    switch (field) {
        case 0: return interp<Fe<Bn254Fr>>(evals, n, x, out);
        case 1: return interp<Fe<Bn254Fq>>(evals, n, x, out);
        case 2: return interp<Fe<PallasFq>>(evals, n, x, out);
        case 3: return interp<Fe<PallasFp>>(evals, n, x, out);
    }
    return -3;
}

template <class F>
static int fold_s(const uint8_t *lo, const uint8_t *hi, const uint8_t *x, const uint8_t *y, uint8_t *out) {
    out_fe(out, ipa_fold_scalar(in_fe<F>(lo), in_fe<F>(hi), in_fe<F>(x), in_fe<F>(y)));
    return 0;
}
extern "C" int sc_test_fold_scalar(int field, const uint8_t *lo, const uint8_t *hi, const uint8_t *x, const uint8_t *y, uint8_t *out) {
    switch (field) {
        case 0: return fold_s<Fe<Bn254Fr>>(lo, hi, x, y, out);
        case 1: return fold_s<Fe<Bn254Fq>>(lo, hi, x, y, out);
        case 2: return fold_s<Fe<PallasFq>>(lo, hi, x, y, out);
        case 3: return fold_s<Fe<PallasFp>>(lo, hi, x, y, out);
    }
    return -3;
}

template <class C>
static int fold_p(const uint8_t *p, const uint8_t *q, const uint8_t *x, const uint8_t *y, uint8_t *out) {
    using F = typename C::Base;
    Affine<F> P, Q;
    P.x = in_fe<F>(p); P.y = in_fe<F>(p + 32); Q.x = in_fe<F>(q); Q.y = in_fe<F>(q + 32);
    uint32_t xs[8], ys[8];
    memcpy(xs, x, 32); memcpy(ys, y, 32);
    Affine<F> r = ipa_fold_point(P, Q, xs, ys);
    out_fe(out, r.x); out_fe(out + 32, r.y);
    return 0;
}
extern "C" int sc_test_fold_point(int curve, const uint8_t *p, const uint8_t *q, const uint8_t *x, const uint8_t *y, uint8_t *out) {
    switch (curve) {
        case 0: return fold_p<CurveBn254G1>(p, q, x, y, out);
        case 1: return fold_p<CurveGrumpkin>(p, q, x, y, out);
        case 2: return fold_p<CurvePallas>(p, q, x, y, out);
        case 3: return fold_p<CurveVesta>(p, q, x, y, out);
    }
    return -3;
}
