// Hash-to-curve for commitment-key generation.
//
// Replaces the per-point body of Arecibo `DlogGroup::from_label` (reached from `public_params`, reference
// src/proof/nova.rs:196-216 / supernova.rs:117-137 via R1CSShape::commitment_key -> CommitmentKey::setup(b"ck", n)):
//     G_i = Curve::hash_to_curve("from_uniform_bytes")(uniform_i)
// for the four curves of the two cycles.  halo2curves 0.6 (BN254 G1, Grumpkin): hash_to_field (expand_message_xmd over
// BLAKE2b-512) -> Shallue-van de Woestijne map (RFC 9380 6.6.1) of both elements -> sum.  pasta_curves 0.5 (Pallas, Vesta):
// same hash_to_field -> simplified SWU onto the 3-isogenous curve -> sum there -> isogeny (WB2019 4.3).
// Everything is LURK_HD: the CPU test-suite runs these exact templates on the host against oracle/h2c.py.
//
// One point costs ~9 fixed-exponent exponentiations (square roots, one batched and one final inversion): the kernel is bound by
// the integer multiplier like every other kernel of this library; BLAKE2b is noise (4 compressions per point).
#pragma once
#include "curve.cuh"

// The building blocks below are deliberately NOT inlined on the device: one point runs ~5 square roots, 3 inversions and 4
// BLAKE2b compressions, and inlining every one of them at every call site makes the kernel ~10x larger (and nvcc minutes
// slower) for nothing -- each call amortises over hundreds of field products.
#if defined(__CUDACC__)
#define LURK_HD_NI __host__ __device__ __noinline__
#else
#define LURK_HD_NI inline
#endif

namespace lurk {

// ------------------------------------------------------------------------------------------------ BLAKE2b (RFC 7693)
struct Blake2b {
    uint64_t h[8];
    LURK_HD static constexpr uint64_t IV(int i) {
        constexpr uint64_t t[8] = {0x6a09e667f3bcc908ull, 0xbb67ae8584caa73bull, 0x3c6ef372fe94f82bull, 0xa54ff53a5f1d36f1ull,
                                   0x510e527fade682d1ull, 0x9b05688c2b3e6c1full, 0x1f83d9abfb41bd6bull, 0x5be0cd19137e2179ull};
        return t[i];
    }
    LURK_HD static constexpr uint8_t SIGMA(int r, int i) {
        constexpr uint8_t s[10][16] = {
            {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
            {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
            {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
            {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
            {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0}};
        return s[r][i];
    }
    LURK_HD static uint64_t rotr(uint64_t x, int n) { return (x >> n) | (x << (64 - n)); }
    // unkeyed, 64-byte digest, empty salt / personalisation (blake2b_simd Params::new().hash_length(64))
    LURK_HD void init() {
        for (int i = 0; i < 8; i++) h[i] = IV(i);
        h[0] ^= 0x01010040ull;
    }
    // one 128-byte block (16 little-endian words); t = bytes hashed so far including this block
    LURK_HD_NI void compress(const uint64_t m[16], uint64_t t, bool last) {
        uint64_t v[16];
        for (int i = 0; i < 8; i++) { v[i] = h[i]; v[i + 8] = IV(i); }
        v[12] ^= t;
        if (last) v[14] = ~v[14];
#define LURK_B2_G(a, b, c, d, x, y)                     \
    v[a] = v[a] + v[b] + (x); v[d] = rotr(v[d] ^ v[a], 32); \
    v[c] = v[c] + v[d];       v[b] = rotr(v[b] ^ v[c], 24); \
    v[a] = v[a] + v[b] + (y); v[d] = rotr(v[d] ^ v[a], 16); \
    v[c] = v[c] + v[d];       v[b] = rotr(v[b] ^ v[c], 63);
#pragma unroll
        for (int r = 0; r < 12; r++) {
            const int s = r % 10;
            LURK_B2_G(0, 4, 8, 12, m[SIGMA(s, 0)], m[SIGMA(s, 1)])
            LURK_B2_G(1, 5, 9, 13, m[SIGMA(s, 2)], m[SIGMA(s, 3)])
            LURK_B2_G(2, 6, 10, 14, m[SIGMA(s, 4)], m[SIGMA(s, 5)])
            LURK_B2_G(3, 7, 11, 15, m[SIGMA(s, 6)], m[SIGMA(s, 7)])
            LURK_B2_G(0, 5, 10, 15, m[SIGMA(s, 8)], m[SIGMA(s, 9)])
            LURK_B2_G(1, 6, 11, 12, m[SIGMA(s, 10)], m[SIGMA(s, 11)])
            LURK_B2_G(2, 7, 8, 13, m[SIGMA(s, 12)], m[SIGMA(s, 13)])
            LURK_B2_G(3, 4, 9, 14, m[SIGMA(s, 14)], m[SIGMA(s, 15)])
        }
#undef LURK_B2_G
        for (int i = 0; i < 8; i++) h[i] ^= v[i] ^ v[i + 8];
    }
    LURK_HD uint8_t digest_byte(int j) const { return (uint8_t)(h[j >> 3] >> (8 * (j & 7))); }
};

// Little helper: a 128-byte block assembled byte by byte
struct Blake2bBlock {
    uint64_t m[16];
    LURK_HD void clear() { for (int i = 0; i < 16; i++) m[i] = 0; }
    LURK_HD void put(int pos, uint8_t b) { m[pos >> 3] |= (uint64_t)b << (8 * (pos & 7)); }
};

// ------------------------------------------------------------------------------------------------ parameters
constexpr int H2C_MAX_DST = 120;   // DST' = DST || len(DST), see hash_to_field
constexpr int H2C_MAX_MSG = 64;

template <class F>
struct H2cParams {
    int method;                // 0 = SVDW (a = 0 curves of halo2curves), 1 = SSWU + 3-isogeny (pasta_curves)
    F b;                       // target curve y^2 = x^3 + b
    F z;                       // SVDW: Z = 1; SSWU: Z = -13
    F c1, c2, c3, c4;          // SVDW constants (RFC 9380 6.6.1)
    F iso_a, iso_b;            // SSWU: the isogenous curve y^2 = x^3 + a x + b
    F nb_over_a, b_over_za;    // SSWU: -b/a and b/(Z a)
    F iso[13];                 // SSWU: isogeny constants (pasta_curves ISOGENY_CONSTANTS layout)
    uint32_t sqrt_exp[8];      // (t - 1) / 2 with p - 1 = 2^s t, t odd
    uint8_t dst[H2C_MAX_DST];  // DST' bytes
    uint32_t dst_len;
    Blake2b zero_block;        // BLAKE2b state after the 128 zero bytes that expand_message_xmd prepends to b_0's input
};

// ------------------------------------------------------------------------------------------------ field helpers
// x^e for a raw 256-bit exponent that is the same for every thread: MSB-first square and multiply (no divergence)
template <class F>
LURK_HD_NI F pow_fixed(const F &x, const uint32_t e[8]) {
    int top = 255;
    while (top > 0 && !((e[top >> 5] >> (top & 31)) & 1)) top--;
    F acc = ((e[top >> 5] >> (top & 31)) & 1) ? x : F::one();
#pragma unroll 1
    for (int i = top - 1; i >= 0; i--) {
        acc = acc.sqr();
        if ((e[i >> 5] >> (i & 31)) & 1) acc = acc * x;
    }
    return acc;
}
template <class F>
LURK_HD F inv_fixed(const F &x) {        // Fermat, fixed structure (the binary-GCD inversion diverges inside a warp); 0 -> 0
    uint32_t e[8];
    e[0] = cc::sub_cc(F::Params::MOD(0), 2);
    for (int i = 1; i < 8; i++) e[i] = cc::subc_cc(F::Params::MOD(i), 0);
    return pow_fixed(x, e);
}
// Tonelli-Shanks with a fixed instruction sequence (RFC 9380 I.4): returns z and whether z^2 == x.
// s (s - 1) / 2 extra squarings for two-adicity s: 0 for BN254 Fq, 378 for BN254 Fr, 496 for the Pasta fields.
template <class F>
LURK_HD_NI F sqrt_fixed(const F &x, const uint32_t sqrt_exp[8], bool *is_square) {
    constexpr int S = F::Params::TWO_ADICITY;
    F z = pow_fixed(x, sqrt_exp);      // x^((t-1)/2)
    F t = z.sqr() * x;                 // x^t
    z = z * x;                         // x^((t+1)/2)
    F c;
    for (int i = 0; i < 8; i++) c.v[i] = F::Params::ROOT(i);   // g^t, g the field's multiplicative generator (a non-square)
    F b = t;
    const F one = F::one();
#pragma unroll 1
    for (int i = S; i >= 2; i--) {
#pragma unroll 1
        for (int j = 1; j <= i - 2; j++) b = b.sqr();
        const bool e = (b == one);
        const F zt = z * c;
        if (!e) z = zt;
        c = c.sqr();
        const F tt = t * c;
        if (!e) t = tt;
        b = t;
    }
    *is_square = (z.sqr() == x);
    return z;
}
template <class F>
LURK_HD uint32_t sgn0(const F &x_mont) { return x_mont.to_canonical().v[0] & 1u; }

// ------------------------------------------------------------------------------------------------ hash_to_field
// expand_message_xmd(msg, DST, 128) with BLAKE2b-512 (RFC 9380 5.3.1, ell = 2), the two 64-byte halves read big-endian mod p.
// Requires msg_len + 3 + dst_len <= 128 and 65 + dst_len <= 128 (checked by the host entry points).
template <class F>
LURK_HD_NI void hash_to_field(const H2cParams<F> &P, const uint8_t *msg, uint32_t msg_len, F out[2]) {
    Blake2bBlock blk;
    // b_0 = H(Z_pad || msg || I2OSP(128, 2) || I2OSP(0, 1) || DST')
    Blake2b b0 = P.zero_block;
    blk.clear();
    int pos = 0;
    for (uint32_t i = 0; i < msg_len; i++) blk.put(pos++, msg[i]);
    blk.put(pos++, 0); blk.put(pos++, 128); blk.put(pos++, 0);
    for (uint32_t i = 0; i < P.dst_len; i++) blk.put(pos++, P.dst[i]);
    b0.compress(blk.m, 128 + (uint64_t)pos, true);
    // b_1 = H(b_0 || 1 || DST'),  b_2 = H((b_0 xor b_1) || 2 || DST')
    Blake2b bi[2];
    for (int k = 0; k < 2; k++) {
        blk.clear();
        for (int w = 0; w < 8; w++) blk.m[w] = k == 0 ? b0.h[w] : (b0.h[w] ^ bi[0].h[w]);
        pos = 64;
        blk.put(pos++, (uint8_t)(k + 1));
        for (uint32_t i = 0; i < P.dst_len; i++) blk.put(pos++, P.dst[i]);
        bi[k].init();
        bi[k].compress(blk.m, (uint64_t)pos, true);
    }
    for (int k = 0; k < 2; k++) {
        // digest bytes d[0..63] big-endian: value = hi * 2^256 + lo
        F lo, hi;
        for (int w = 0; w < 8; w++) {
            uint32_t l = 0, h = 0;
            for (int byte = 0; byte < 4; byte++) {
                l |= (uint32_t)bi[k].digest_byte(63 - 4 * w - byte) << (8 * byte);
                h |= (uint32_t)bi[k].digest_byte(31 - 4 * w - byte) << (8 * byte);
            }
            lo.v[w] = l; hi.v[w] = h;
        }
        // the unreduced 256-bit halves go in as the MULTIPLIER operand (the multiplicand must be < p)
        const F rr = F::rr();
        out[k] = rr * lo + rr * (rr * hi);     // lo R + hi R^2 = Montgomery form of lo + hi 2^256
    }
}

// ------------------------------------------------------------------------------------------------ affine + affine -> XYZZ, any a
template <class F>
LURK_HD XYZZ<F> add_affine_pair(const Affine<F> &p, const Affine<F> &q, const F &a) {
    XYZZ<F> r;
    F pp_ = q.x - p.x, rr_ = q.y - p.y;
    if (pp_.is_zero()) {
        if (!rr_.is_zero() || p.y.is_zero()) return XYZZ<F>::identity();
        F u = p.y.dbl(), v = u.sqr(), w = u * v, s = p.x * v;
        F xx = p.x.sqr();
        F m = xx.dbl() + xx + a;
        r.x = m.sqr() - s.dbl();
        r.y = mul_sub_mul(m, s - r.x, w, p.y);
        r.zz = v; r.zzz = w;
        return r;
    }
    F pp = pp_.sqr(), ppp = pp_ * pp, q_ = p.x * pp;
    r.x = rr_.sqr() - ppp - q_.dbl();
    r.y = mul_sub_mul(rr_, q_ - r.x, p.y, ppp);
    r.zz = pp; r.zzz = ppp;
    return r;
}

// ------------------------------------------------------------------------------------------------ the two maps
// First half of either map up to the value that has to be inverted (both elements share ONE inversion).
template <class F>
struct H2cHalf { F u, t1, t2, den; };

template <class F>
LURK_HD H2cHalf<F> map_prepare(const H2cParams<F> &P, const F &u) {
    H2cHalf<F> h;
    h.u = u;
    if (P.method == 0) {                    // SVDW steps 1-5
        F tv1 = u.sqr() * P.c1;
        h.t2 = F::one() + tv1;
        h.t1 = F::one() - tv1;
        h.den = h.t1 * h.t2;
    } else {                                // SSWU: ta = Z^2 u^4 + Z u^2
        F zu2 = P.z * u.sqr();
        h.t1 = zu2;
        h.t2 = F::zero();
        h.den = zu2.sqr() + zu2;
    }
    return h;
}
template <class F>
LURK_HD F curve_rhs(const F &x, const F &a, const F &b) { return (x.sqr() + a) * x + b; }

// second half: inv = inv0(den)
template <class F>
LURK_HD_NI Affine<F> map_finish(const H2cParams<F> &P, const H2cHalf<F> &h, const F &inv) {
    Affine<F> r;
    bool sq;
    F y;
    if (P.method == 0) {                    // SVDW steps 7-35
        const F zero = F::zero();
        F tv4 = h.u * h.t1 * inv * P.c3;
        F x = P.c2 - tv4;
        F gx = curve_rhs(x, zero, P.b);
        y = sqrt_fixed(gx, P.sqrt_exp, &sq);
        if (!sq) {
            x = P.c2 + tv4;
            gx = curve_rhs(x, zero, P.b);
            y = sqrt_fixed(gx, P.sqrt_exp, &sq);
            if (!sq) {
                F x3 = h.t2.sqr() * inv;
                x = x3.sqr() * P.c4 + P.z;
                gx = curve_rhs(x, zero, P.b);
                y = sqrt_fixed(gx, P.sqrt_exp, &sq);   // always a square here
            }
        }
        r.x = x;
    } else {                                // SSWU steps 2-8
        F x1 = inv.is_zero() ? P.b_over_za : P.nb_over_a * (F::one() + inv);
        F gx = curve_rhs(x1, P.iso_a, P.iso_b);
        y = sqrt_fixed(gx, P.sqrt_exp, &sq);
        r.x = x1;
        if (!sq) {
            r.x = h.t1 * x1;                // x2 = Z u^2 x1; g(x2) is a square whenever g(x1) is not
            gx = curve_rhs(r.x, P.iso_a, P.iso_b);
            y = sqrt_fixed(gx, P.sqrt_exp, &sq);
        }
    }
    r.y = (sgn0(h.u) != sgn0(y)) ? y.neg() : y;
    return r;
}

// ------------------------------------------------------------------------------------------------ one point
// hash_to_curve(domain_prefix)(msg) as an affine point in Montgomery form; the identity is (0, 0)
template <class F>
LURK_HD Affine<F> hash_to_curve_point(const H2cParams<F> &P, const uint8_t *msg, uint32_t msg_len) {
    F u[2];
    hash_to_field(P, msg, msg_len, u);
    H2cHalf<F> h0 = map_prepare(P, u[0]), h1 = map_prepare(P, u[1]);
    // inv0 of both denominators with one exponentiation (zeros are replaced by one and restored afterwards)
    const bool z0 = h0.den.is_zero(), z1 = h1.den.is_zero();
    const F d0 = z0 ? F::one() : h0.den, d1 = z1 ? F::one() : h1.den;
    const F inv01 = inv_fixed(d0 * d1);
    const F i0 = z0 ? F::zero() : inv01 * d1, i1 = z1 ? F::zero() : inv01 * d0;
    const Affine<F> q0 = map_finish(P, h0, i0), q1 = map_finish(P, h1, i1);
    Affine<F> out;
    if (P.method == 0) {
        XYZZ<F> r = add_affine_pair(q0, q1, F::zero());
        if (r.is_identity()) { out.x = F::zero(); out.y = F::zero(); return out; }
        F zi = inv_fixed(r.zzz);
        F zz_inv = (zi * r.zz).sqr();
        out.x = r.x * zz_inv;
        out.y = r.y * zi;
        return out;
    }
    // Pasta: the sum lives on the isogenous curve; isogeny evaluated on x = X / ZZ, y = Y / ZZZ with one inversion
    XYZZ<F> r = add_affine_pair(q0, q1, P.iso_a);
    const F zz = r.zz, zz2 = zz.sqr(), zz3 = zz2 * zz;
    const F nx = ((P.iso[0] * r.x + P.iso[1] * zz) * r.x + P.iso[2] * zz2) * r.x + P.iso[3] * zz3;
    const F dx = (r.x + P.iso[4] * zz) * r.x + P.iso[5] * zz2;
    const F ny = ((P.iso[6] * r.x + P.iso[7] * zz) * r.x + P.iso[8] * zz2) * r.x + P.iso[9] * zz3;
    const F dy = ((r.x + P.iso[10] * zz) * r.x + P.iso[11] * zz2) * r.x + P.iso[12] * zz3;
    const F den1 = dx * zz, den2 = dy * r.zzz;        // x' = nx / den1, y' = Y ny / den2
    const F den = den1 * den2;
    if (den.is_zero()) { out.x = F::zero(); out.y = F::zero(); return out; }   // identity or the kernel of the isogeny
    const F inv = inv_fixed(den);
    out.x = nx * den2 * inv;
    out.y = r.y * ny * den1 * inv;
    return out;
}

}  // namespace lurk
