// Host build of lurk-beta_b200/csrc/h2c.cuh + h2c_params.h: the exact hash-to-curve templates the GPU kernel runs, compiled for
// the CPU (with -DLURK_HOST_EMULATE_CC also the GPU limb arithmetic), so that the CPU test-suite can compare them with
// oracle/h2c.py point by point.  Test-only helper, not part of the product.
#include "h2c_params.h"
using namespace lurk;

template <class C>
static int point(const char *prefix, const uint8_t *msg, size_t msg_len, uint8_t *out) {
    using F = typename C::Base;
    H2cParams<F> P;
    if (!h2c_make_params<C>(prefix, msg_len, P)) return -1;
    Affine<F> a = hash_to_curve_point(P, msg, (uint32_t)msg_len);
    F x = a.x.to_canonical(), y = a.y.to_canonical();
    memcpy(out, x.v, 32);
    memcpy(out + 32, y.v, 32);
    return 0;
}
extern "C" int h2c_test_point(int curve, const char *prefix, const uint8_t *msg, size_t msg_len, uint8_t *out) {
    switch (curve) {
        case 0: return point<CurveBn254G1>(prefix, msg, msg_len, out);
        case 1: return point<CurveGrumpkin>(prefix, msg, msg_len, out);
        case 2: return point<CurvePallas>(prefix, msg, msg_len, out);
        case 3: return point<CurveVesta>(prefix, msg, msg_len, out);
    }
    return -3;
}

template <class C>
static int field_pair(const char *prefix, const uint8_t *msg, size_t msg_len, uint8_t *out) {
    using F = typename C::Base;
    H2cParams<F> P;
    if (!h2c_make_params<C>(prefix, msg_len, P)) return -1;
    F u[2];
    hash_to_field(P, msg, (uint32_t)msg_len, u);
    for (int k = 0; k < 2; k++) { F c = u[k].to_canonical(); memcpy(out + 32 * k, c.v, 32); }
    return 0;
}
extern "C" int h2c_test_hash_to_field(int curve, const char *prefix, const uint8_t *msg, size_t msg_len, uint8_t *out) {
    switch (curve) {
        case 0: return field_pair<CurveBn254G1>(prefix, msg, msg_len, out);
        case 1: return field_pair<CurveGrumpkin>(prefix, msg, msg_len, out);
        case 2: return field_pair<CurvePallas>(prefix, msg, msg_len, out);
        case 3: return field_pair<CurveVesta>(prefix, msg, msg_len, out);
    }
    return -3;
}

// one map of a given field element u (canonical in, affine canonical out; for the Pasta curves a point of the ISO curve)
template <class C>
static int map_one(const uint8_t *u_canon, uint8_t *out) {
    using F = typename C::Base;
    H2cParams<F> P;
    if (!h2c_make_params<C>("from_uniform_bytes", 32, P)) return -1;
    F u;
    memcpy(u.v, u_canon, 32);
    if (!u.is_reduced()) return -2;
    u = F::from_canonical(u);
    H2cHalf<F> h = map_prepare(P, u);
    F inv = inv_fixed(h.den);
    Affine<F> a = map_finish(P, h, inv);
    F x = a.x.to_canonical(), y = a.y.to_canonical();
    memcpy(out, x.v, 32);
    memcpy(out + 32, y.v, 32);
    return 0;
}
extern "C" int h2c_test_map(int curve, const uint8_t *u_canon, uint8_t *out) {
    switch (curve) {
        case 0: return map_one<CurveBn254G1>(u_canon, out);
        case 1: return map_one<CurveGrumpkin>(u_canon, out);
        case 2: return map_one<CurvePallas>(u_canon, out);
        case 3: return map_one<CurveVesta>(u_canon, out);
    }
    return -3;
}

template <class F>
static int sqrt_one(const uint8_t *x_canon, uint8_t *out) {
    H2cParams<F> P;
    // only sqrt_exp is needed: rebuild it the way h2c_make_params does
    uint32_t t[8];
    for (int i = 0; i < 8; i++) t[i] = F::Params::MOD(i);
    t[0] -= 1;
    for (int k = 0; k < F::Params::TWO_ADICITY + 1; k++) { for (int i = 0; i < 7; i++) t[i] = (t[i] >> 1) | (t[i + 1] << 31); t[7] >>= 1; }
    F x;
    memcpy(x.v, x_canon, 32);
    x = F::from_canonical(x);
    bool sq = false;
    F z = sqrt_fixed(x, t, &sq).to_canonical();
    memcpy(out, z.v, 32);
    (void)P;
    return sq ? 1 : 0;
}
extern "C" int h2c_test_sqrt(int field, const uint8_t *x_canon, uint8_t *out) {
    switch (field) {
        case 0: return sqrt_one<Fe<Bn254Fr>>(x_canon, out);
        case 1: return sqrt_one<Fe<Bn254Fq>>(x_canon, out);
        case 2: return sqrt_one<Fe<PallasFq>>(x_canon, out);
        case 3: return sqrt_one<Fe<PallasFp>>(x_canon, out);
    }
    return -3;
}

// BLAKE2b-512 of an arbitrary message through the same compression function (multi-block driver only here)
extern "C" void h2c_test_blake2b(const uint8_t *data, size_t len, uint8_t *out) {
    Blake2b b;
    b.init();
    size_t off = 0;
    uint64_t m[16];
    while (len - off > 128) {
        memcpy(m, data + off, 128);
        off += 128;
        b.compress(m, off, false);
    }
    uint8_t last[128] = {0};
    memcpy(last, data + off, len - off);
    memcpy(m, last, 128);
    b.compress(m, len, true);
    for (int j = 0; j < 64; j++) out[j] = b.digest_byte(j);
}

extern "C" void h2c_test_shake256(const uint8_t *in, size_t in_len, uint8_t *out, size_t out_len, size_t step) {
    Shake256 x;
    x.absorb(in, in_len);
    if (!step) step = out_len ? out_len : 1;
    for (size_t off = 0; off < out_len; off += step) x.squeeze(out + off, step < out_len - off ? step : out_len - off);
}

extern "C" int h2c_test_iso_constants(int curve, uint8_t *out /* 14 x 32: iso_a, 13 constants */) {
    auto dump = [&](auto c) {
        using C = decltype(c);
        using F = typename C::Base;
        H2cParams<F> P;
        if (!h2c_make_params<C>("x", 1, P)) return -1;
        F a = P.iso_a.to_canonical();
        memcpy(out, a.v, 32);
        for (int i = 0; i < 13; i++) { F v = P.iso[i].to_canonical(); memcpy(out + 32 * (i + 1), v.v, 32); }
        return 0;
    };
    return curve == 2 ? dump(CurvePallas()) : curve == 3 ? dump(CurveVesta()) : -3;
}
template <class C>
static int svdw_dump(uint8_t *out) {
    using F = typename C::Base;
    H2cParams<F> P;
    if (!h2c_make_params<C>("x", 1, P)) return -1;
    const F *c[4] = {&P.c1, &P.c2, &P.c3, &P.c4};
    for (int i = 0; i < 4; i++) { F v = c[i]->to_canonical(); memcpy(out + 32 * i, v.v, 32); }
    return 0;
}
extern "C" int h2c_test_svdw_constants(int curve, uint8_t *out /* 4 x 32 */) {
    return curve == 0 ? svdw_dump<CurveBn254G1>(out) : curve == 1 ? svdw_dump<CurveGrumpkin>(out) : -3;
}
