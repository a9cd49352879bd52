// Host-side generation of Neptune-compatible Poseidon constants (product code; independent of oracle/).
//
// Mirrors what the reference obtains from `PoseidonConstants::<F, U_A>::new()` (src/hash.rs:41-84):
// Strength::Standard, HashType::MerkleTree; width t = arity + 1; domain tag 2^arity - 1; R_F = 8 and R_P from
// Neptune's round-number search (n = 255, M = 128 hard-coded); round constants from the Poseidon Grain LFSR;
// Cauchy MDS M[i][j] = 1/(i + t + j); plus the "optimised" forms Neptune's static hasher and circuit use:
// compressed round constants, the pre-sparse matrix and the R_P sparse factors.
// neptune is a git dependency of the reference (Cargo.toml:32,127), not in tree; pinned by golden digests.
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

#include "field.cuh"

namespace lurk {

inline void poseidon_round_numbers(int t, int *rf_out, int *rp_out) {
    const double n = 255.0, M = 128.0;
    long best_cost = -1;
    int best_rf = 0, best_rp = 0;
    for (int rp = 1; rp < 200; rp++) {
        for (int rf = 4; rf <= 100; rf += 2) {
            double c = (M <= (n - 3.0) * (t + 1.0)) ? 6.0 : 10.0;
            double rf_interp = 0.43 * M + std::log2((double)t) - rp;
            double rf_grob1 = 0.21 * n - rp;
            double rf_grob2 = (0.14 * n - 1.0 - rp) / (t - 1.0);
            double rf_max = std::fmax(std::fmax(std::ceil(c), std::ceil(rf_interp)),
                                      std::fmax(std::ceil(rf_grob1), std::ceil(rf_grob2)));
            if ((double)rf >= rf_max) {
                int rf2 = rf + 2;
                int rp2 = (int)std::ceil(1.075 * rp);
                long cost = (long)t * rf2 + rp2;
                if (best_cost < 0 || cost < best_cost || (cost == best_cost && rf2 < best_rf)) {
                    best_cost = cost; best_rf = rf2; best_rp = rp2;
                }
            }
        }
    }
    *rf_out = best_rf;
    *rp_out = best_rp;
}

// 80-bit Grain LFSR in self-shrinking mode
class GrainLfsr {
  public:
    GrainLfsr(int nbits, int t, int rf, int rp) {
        int k = 0;
        auto push = [&](uint32_t v, int n) { for (int i = n - 1; i >= 0; i--) s_[k++] = (v >> i) & 1; };
        push(1, 2); push(1, 4); push(nbits, 12); push(t, 12); push(rf, 10); push(rp, 10); push(0x3fffffffu, 30);
        head_ = 0;
        for (int i = 0; i < 160; i++) next();
    }
    int bit() {
        for (;;) { int a = next(), b = next(); if (a) return b; }
    }
  private:
    int next() {
        auto at = [&](int i) { return s_[(head_ + i) % 80]; };
        int b = at(62) ^ at(51) ^ at(38) ^ at(23) ^ at(13) ^ at(0);
        s_[head_] = (uint8_t)b;           // overwrite the oldest bit, it becomes the newest
        head_ = (head_ + 1) % 80;
        return b;
    }
    uint8_t s_[80];
    int head_;
};

template <class F>
struct PoseidonParams {
    int arity = 0, t = 0, rf = 0, rp = 0;
    F domain_tag;
    std::vector<F> round_constants;   // t * (rf + rp), textbook order
    std::vector<F> mds;               // t * t row-major
    std::vector<F> compressed;        // t * rf + rp
    std::vector<F> pre_sparse;        // t * t
    std::vector<F> sparse_w;          // rp * t       first column of each sparse factor
    std::vector<F> sparse_v;          // rp * (t - 1) rest of the first row of each sparse factor

    // flat device image: [compressed | mds | pre_sparse | sparse_w | sparse_v], Montgomery limbs
    size_t off_mds() const { return compressed.size(); }
    size_t off_pre() const { return off_mds() + (size_t)t * t; }
    size_t off_sw() const { return off_pre() + (size_t)t * t; }
    size_t off_sv() const { return off_sw() + (size_t)rp * t; }
    size_t flat_len() const { return off_sv() + (size_t)rp * (t - 1); }
    std::vector<F> flat() const {
        std::vector<F> o;
        o.reserve(flat_len());
        o.insert(o.end(), compressed.begin(), compressed.end());
        o.insert(o.end(), mds.begin(), mds.end());
        o.insert(o.end(), pre_sparse.begin(), pre_sparse.end());
        o.insert(o.end(), sparse_w.begin(), sparse_w.end());
        o.insert(o.end(), sparse_v.begin(), sparse_v.end());
        return o;
    }
    int num_aux() const { return 3 * (t * rf + rp); }
};

namespace detail {
template <class F> using Mat = std::vector<std::vector<F>>;

template <class F>
Mat<F> mat_mul(const Mat<F> &a, const Mat<F> &b) {
    size_t n = a.size(), m = b[0].size(), k = b.size();
    Mat<F> r(n, std::vector<F>(m, F::zero()));
    for (size_t i = 0; i < n; i++)
        for (size_t j = 0; j < m; j++) {
            F acc = F::zero();
            for (size_t x = 0; x < k; x++) acc = acc + a[i][x] * b[x][j];
            r[i][j] = acc;
        }
    return r;
}
template <class F>
Mat<F> mat_inv(const Mat<F> &a) {
    size_t n = a.size();
    Mat<F> m(n, std::vector<F>(2 * n, F::zero()));
    for (size_t i = 0; i < n; i++) {
        for (size_t j = 0; j < n; j++) m[i][j] = a[i][j];
        m[i][n + i] = F::one();
    }
    for (size_t c = 0; c < n; c++) {
        size_t piv = c;
        while (m[piv][c].is_zero()) piv++;
        std::swap(m[c], m[piv]);
        F inv = m[c][c].inv();
        for (auto &v : m[c]) v = v * inv;
        for (size_t r = 0; r < n; r++) {
            if (r == c || m[r][c].is_zero()) continue;
            F f = m[r][c];
            for (size_t j = 0; j < 2 * n; j++) m[r][j] = m[r][j] - f * m[c][j];
        }
    }
    Mat<F> out(n, std::vector<F>(n));
    for (size_t i = 0; i < n; i++)
        for (size_t j = 0; j < n; j++) out[i][j] = m[i][n + j];
    return out;
}
// row vector times matrix
template <class F>
std::vector<F> vec_mat(const std::vector<F> &v, const Mat<F> &m) {
    size_t t = v.size();
    std::vector<F> o(t, F::zero());
    for (size_t j = 0; j < t; j++) {
        F acc = F::zero();
        for (size_t i = 0; i < t; i++) acc = acc + v[i] * m[i][j];
        o[j] = acc;
    }
    return o;
}
}  // namespace detail

template <class F>
PoseidonParams<F> make_poseidon_params(int arity) {
    using P = typename F::Params;
    using namespace detail;
    PoseidonParams<F> pp;
    const int t = arity + 1;
    pp.arity = arity;
    pp.t = t;
    poseidon_round_numbers(t, &pp.rf, &pp.rp);
    const int rf = pp.rf, rp = pp.rp, half = rf / 2;
    pp.domain_tag = F::from_u64((1ull << arity) - 1);

    // round constants: NBITS bits MSB-first per candidate, rejection-sampled below p
    GrainLfsr g(P::NBITS, t, rf, rp);
    pp.round_constants.resize((size_t)t * (rf + rp));
    for (auto &rc : pp.round_constants) {
        for (;;) {
            F raw = F::zero();
            for (int b = P::NBITS - 1; b >= 0; b--)
                if (g.bit()) raw.v[b >> 5] |= 1u << (b & 31);
            if (raw.is_reduced()) { rc = F::from_canonical(raw); break; }
        }
    }
    Mat<F> mds(t, std::vector<F>(t));
    for (int i = 0; i < t; i++)
        for (int j = 0; j < t; j++) mds[i][j] = F::from_u64((uint64_t)(i + t + j)).inv();
    for (int i = 0; i < t; i++)
        for (int j = 0; j < t; j++) pp.mds.push_back(mds[i][j]);

    // --- compressed round constants
    Mat<F> minv = mat_inv(mds);
    auto rnd = [&](int r) { return std::vector<F>(pp.round_constants.begin() + (size_t)r * t, pp.round_constants.begin() + (size_t)(r + 1) * t); };
    std::vector<F> comp = rnd(0);
    for (int i = 0; i < half - 1; i++) { auto v = vec_mat(rnd(i + 1), minv); comp.insert(comp.end(), v.begin(), v.end()); }
    std::vector<F> acc = rnd(half + rp);
    std::vector<F> partial_keys;
    for (int i = 0; i < rp; i++) {
        auto inv = vec_mat(acc, minv);
        partial_keys.push_back(inv[0]);
        inv[0] = F::zero();
        auto prev = rnd(half + rp - 1 - i);
        for (int k = 0; k < t; k++) acc[k] = prev[k] + inv[k];
    }
    { auto v = vec_mat(acc, minv); comp.insert(comp.end(), v.begin(), v.end()); }
    for (int i = rp - 1; i >= 0; i--) comp.push_back(partial_keys[i]);
    for (int i = 1; i < half; i++) { auto v = vec_mat(rnd(half + rp + i), minv); comp.insert(comp.end(), v.begin(), v.end()); }
    pp.compressed = comp;

    // --- sparse factorisation of the partial-round matrices
    Mat<F> cur = mds;
    std::vector<std::vector<F>> ws, vs;
    for (int r = 0; r < rp; r++) {
        Mat<F> hat(t - 1, std::vector<F>(t - 1));
        for (int i = 1; i < t; i++)
            for (int j = 1; j < t; j++) hat[i - 1][j - 1] = cur[i][j];
        Mat<F> hat_inv = mat_inv(hat);
        std::vector<F> w(t), v(t - 1);
        w[0] = cur[0][0];
        for (int i = 0; i < t - 1; i++) {
            F a = F::zero();
            for (int k = 0; k < t - 1; k++) a = a + hat_inv[i][k] * cur[k + 1][0];
            w[i + 1] = a;
        }
        for (int j = 1; j < t; j++) v[j - 1] = cur[0][j];
        ws.push_back(w);
        vs.push_back(v);
        Mat<F> mprime(t, std::vector<F>(t, F::zero()));
        mprime[0][0] = F::one();
        for (int i = 1; i < t; i++)
            for (int j = 1; j < t; j++) mprime[i][j] = hat[i - 1][j - 1];
        cur = mat_mul(mds, mprime);
    }
    for (int i = 0; i < t; i++)
        for (int j = 0; j < t; j++) pp.pre_sparse.push_back(cur[i][j]);
    for (int r = rp - 1; r >= 0; r--) {
        pp.sparse_w.insert(pp.sparse_w.end(), ws[r].begin(), ws[r].end());
        pp.sparse_v.insert(pp.sparse_v.end(), vs[r].begin(), vs[r].end());
    }
    return pp;
}

}  // namespace lurk
