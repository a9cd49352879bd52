// Host build of lurk-beta_b200/csrc/kzg.cuh: the segment functions of the HyperKZG prover kernels and of the powers-of-tau kernel,
// driven by plain loops in place of the grid (same level structure as kzg_witness_polys in kzg.cu), against oracle/kzg.py.
// Test-only helper.
#include "kzg.cuh"
#include <vector>
using namespace lurk;

template <class F> static F in_fe(const uint8_t *p) { F v; memcpy(v.v, p, 32); return F::from_canonical(v); }
template <class F> static void out_fe(uint8_t *p, const F &x) { F c = x.to_canonical(); memcpy(p, c.v, 32); }

// witness polynomial h of B at u by the up-sweep / down-sweep (segments of KZG_SEG), and B(u) as a by-product
template <class F>
static int witness(const uint8_t *B_bytes, size_t n, const uint8_t *u_bytes, uint8_t *h_bytes, uint8_t *eval_bytes) {
    std::vector<F> B(n), h(n, F::zero());
    for (size_t i = 0; i < n; i++) B[i] = in_fe<F>(B_bytes + 32 * i);
    const F u = in_fe<F>(u_bytes);
    std::vector<std::vector<F>> Y{B};
    std::vector<F> mult{u};
    while (Y.back().size() > (size_t)KZG_SEG) {
        const std::vector<F> &X = Y.back();
        const size_t nseg = (X.size() + KZG_SEG - 1) / KZG_SEG;
        std::vector<F> nxt(nseg);
        for (size_t s = 0; s < nseg; s++) nxt[s] = kzg_seg_horner(X.data(), s * KZG_SEG, std::min(X.size(), (s + 1) * KZG_SEG), mult.back());
        mult.push_back(kzg_pow_small(mult.back(), KZG_SEG));
        Y.push_back(nxt);
    }
    const int K = (int)Y.size() - 1;
    for (int k = K; k >= 0; k--) {
        std::vector<F> &X = Y[k];
        const size_t nseg = (X.size() + KZG_SEG - 1) / KZG_SEG;
        std::vector<F> out(X.size(), F::zero());
        for (size_t s = 0; s < nseg; s++) {
            const F carry = (k < K && s + 1 < nseg) ? Y[k + 1][s + 1] : F::zero();
            if (k == 0) kzg_seg_down(X.data(), s * KZG_SEG, std::min(X.size(), (s + 1) * KZG_SEG), mult[k], carry, h.data(), 1);
            else kzg_seg_down(X.data(), s * KZG_SEG, std::min(X.size(), (s + 1) * KZG_SEG), mult[k], carry, out.data(), 0);
        }
        if (k) X = out;
    }
    for (size_t i = 0; i < n; i++) out_fe(h_bytes + 32 * i, h[i]);
    // B(u) = B[0] + u h[0]
    out_fe(eval_bytes, B[0] + u * h[0]);
    return 0;
}
extern "C" int kzg_test_witness(int field, const uint8_t *B, size_t n, const uint8_t *u, uint8_t *h, uint8_t *eval) {
    switch (field) {
        case 0: return witness<Fe<Bn254Fr>>(B, n, u, h, eval);
        case 1: return witness<Fe<Bn254Fq>>(B, n, u, h, eval);
        case 2: return witness<Fe<PallasFq>>(B, n, u, h, eval);
        case 3: return witness<Fe<PallasFp>>(B, n, u, h, eval);
    }
    return -3;
}

template <class F>
static int fold(const uint8_t *in, size_t half, const uint8_t *x, uint8_t *out) {
    const F xx = in_fe<F>(x);
    for (size_t j = 0; j < half; j++) out_fe(out + 32 * j, kzg_fold_low(in_fe<F>(in + 64 * j), in_fe<F>(in + 64 * j + 32), xx));
    return 0;
}
extern "C" int kzg_test_fold(int field, const uint8_t *in, size_t half, const uint8_t *x, uint8_t *out) {
    switch (field) {
        case 0: return fold<Fe<Bn254Fr>>(in, half, x, out);
        case 2: return fold<Fe<PallasFq>>(in, half, x, out);
    }
    return -3;
}
extern "C" size_t kzg_test_offset(size_t n, int j) { return kzg_poly_offset(n, j); }

// [s] g through an 8-bit window table built the way kzg.cu builds it (table passed in by the test: 32 x 255 affine canonical points)
template <class C>
static int fixed_mul(const uint8_t *table_bytes, const uint8_t *s_bytes, uint8_t *out) {
    using F = typename C::Base;
    std::vector<Affine<F>> table(KZG_WINDOWS * 255);
    for (size_t i = 0; i < table.size(); i++) { table[i].x = in_fe<F>(table_bytes + 64 * i); table[i].y = in_fe<F>(table_bytes + 64 * i + 32); }
    uint32_t s[8];
    memcpy(s, s_bytes, 32);
    Affine<F> a = kzg_fixed_base_mul(table.data(), s).to_affine();
    out_fe(out, a.x); out_fe(out + 32, a.y);
    return 0;
}
extern "C" int kzg_test_fixed_mul(int curve, const uint8_t *table, const uint8_t *s, uint8_t *out) {
    switch (curve) {
        case 0: return fixed_mul<CurveBn254G1>(table, s, out);
        case 2: return fixed_mul<CurvePallas>(table, s, out);
    }
    return -3;
}
