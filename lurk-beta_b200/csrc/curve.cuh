// Short-Weierstrass a = 0 group law (y^2 = x^3 + b) for the four curves on the lurk-beta proving path:
// BN254 G1 / Grumpkin (Arecibo Bn256EngineKZG / GrumpkinEngine, reference src/proof/nova.rs:57-71) and
// Pallas / Vesta.  Bucket accumulators use XYZZ coordinates (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2):
// mixed addition with an affine base costs 8M + 2S and needs no inversion; infinity is ZZ == 0.
// Affine identity is (0, 0) as in pasta_curves / halo2curves.
#pragma once
#include "field.cuh"

namespace lurk {

template <class F>
struct Affine {
    F x, y;
    LURK_HD bool is_identity() const { return x.is_zero() && y.is_zero(); }
};

// a*b - c*d with one Montgomery reduction (lazy 512-bit accumulation): 200 instead of 272 IMAD.WIDE
template <class F>
LURK_HD F mul_sub_mul(const F &a, const F &b, const F &c, const F &d) {
    WideAcc<typename F::Params> acc;
    acc.clear();
    acc.mul_acc(a, b);
    acc.mul_acc(c.neg(), d);
    return acc.reduce();
}

template <class F>
struct XYZZ {
    F x, y, zz, zzz;

    LURK_HD static XYZZ identity() { XYZZ r; r.x = F::zero(); r.y = F::zero(); r.zz = F::zero(); r.zzz = F::zero(); return r; }
    LURK_HD bool is_identity() const { return zz.is_zero(); }
    LURK_HD static XYZZ from_affine(const Affine<F> &p) {
        if (p.is_identity()) return identity();
        XYZZ r; r.x = p.x; r.y = p.y; r.zz = F::one(); r.zzz = F::one(); return r;
    }
    LURK_HD XYZZ neg() const { XYZZ r = *this; r.y = y.neg(); return r; }

    // 2 * affine point (mdbl-2008-s-1)
    LURK_HD static XYZZ dbl_affine(const Affine<F> &p) {
        if (p.is_identity() || p.y.is_zero()) return identity();
        F u = p.y.dbl();
        F v = u.sqr();
        F w = u * v;
        F s = p.x * v;
        F xx = p.x.sqr();
        F m = xx.dbl() + xx;
        XYZZ r;
        r.x = m.sqr() - s.dbl();
        r.y = mul_sub_mul(m, s - r.x, w, p.y);
        r.zz = v;
        r.zzz = w;
        return r;
    }
    // dbl-2008-s-1
    LURK_HD XYZZ dbl() const {
        if (is_identity() || y.is_zero()) return identity();
        F u = y.dbl();
        F v = u.sqr();
        F w = u * v;
        F s = x * v;
        F xx = x.sqr();
        F m = xx.dbl() + xx;
        XYZZ r;
        r.x = m.sqr() - s.dbl();
        r.y = mul_sub_mul(m, s - r.x, w, y);
        r.zz = v * zz;
        r.zzz = w * zzz;
        return r;
    }
    // this += affine (madd-2008-s); `negate` adds -q
    LURK_HD void add_affine(const Affine<F> &q, bool negate = false) {
        if (q.is_identity()) return;
        F qy = negate ? q.y.neg() : q.y;
        if (is_identity()) { x = q.x; y = qy; zz = F::one(); zzz = F::one(); return; }
        F u2 = q.x * zz;
        F s2 = qy * zzz;
        F p = u2 - x;
        F r = s2 - y;
        if (p.is_zero()) {
            if (r.is_zero()) { Affine<F> t; t.x = q.x; t.y = qy; *this = dbl_affine(t); }
            else *this = identity();
            return;
        }
        F pp = p.sqr();
        F ppp = p * pp;
        F q_ = x * pp;
        F x3 = r.sqr() - ppp - q_.dbl();
        y = mul_sub_mul(r, q_ - x3, y, ppp);
        x = x3;
        zz = zz * pp;
        zzz = zzz * ppp;
    }
    // this += o (add-2008-s)
    LURK_HD void add(const XYZZ &o) {
        if (o.is_identity()) return;
        if (is_identity()) { *this = o; return; }
        F u1 = x * o.zz;
        F u2 = o.x * zz;
        F s1 = y * o.zzz;
        F s2 = o.y * zzz;
        F p = u2 - u1;
        F r = s2 - s1;
        if (p.is_zero()) {
            if (r.is_zero()) *this = dbl(); else *this = identity();
            return;
        }
        F pp = p.sqr();
        F ppp = p * pp;
        F q_ = u1 * pp;
        F x3 = r.sqr() - ppp - q_.dbl();
        y = mul_sub_mul(r, q_ - x3, s1, ppp);
        x = x3;
        zz = zz * o.zz * pp;
        zzz = zzz * o.zzz * ppp;
    }
    // one inversion; identity -> (0, 0)
    LURK_HD Affine<F> to_affine() const {
        Affine<F> a;
        if (is_identity()) { a.x = F::zero(); a.y = F::zero(); return a; }
        F zi = zzz.inv();              // 1/ZZZ
        F zz_inv = (zi * zz).sqr();    // (ZZ/ZZZ)^2 = 1/ZZ   (ZZ^3 = ZZZ^2)
        a.x = x * zz_inv;
        a.y = y * zi;
        return a;
    }
    // [k] * this, k a small unsigned integer (host + bucket-reduce use)
    LURK_HD XYZZ mul_u32(uint32_t k) const {
        XYZZ acc = identity();
        for (int b = 31; b >= 0; b--) {
            acc = acc.dbl();
            if ((k >> b) & 1) acc.add(*this);
        }
        return acc;
    }
};

// Standard generators (canonical integers): BN254 G1 (1, 2); Grumpkin (1, sqrt(-16)); Pallas / Vesta (-1, 2).
// GEN_X_NEG1: x = p - 1.  GEN_Y: canonical y as 8 x u32, little-endian.
struct CurveBn254G1 {
    using Base = Fe<Bn254Fq>; using Scalar = Fe<Bn254Fr>;
    static constexpr int ID = 0; static constexpr bool GEN_X_NEG1 = false;
    LURK_HD static constexpr uint32_t GEN_Y(int i) { constexpr uint32_t t[8] = {2, 0, 0, 0, 0, 0, 0, 0}; return t[i]; }
};
struct CurveGrumpkin {
    using Base = Fe<Bn254Fr>; using Scalar = Fe<Bn254Fq>;
    static constexpr int ID = 1; static constexpr bool GEN_X_NEG1 = false;
    LURK_HD static constexpr uint32_t GEN_Y(int i) {
        constexpr uint32_t t[8] = {0x823f272cu, 0x833fc48du, 0xf1181294u, 0x2d270d45u, 0x06a45d63u, 0xcf135e75u, 0x00000002u, 0x00000000u};
        return t[i];
    }
};
struct CurvePallas {
    using Base = Fe<PallasFp>; using Scalar = Fe<PallasFq>;
    static constexpr int ID = 2; static constexpr bool GEN_X_NEG1 = true;
    LURK_HD static constexpr uint32_t GEN_Y(int i) { constexpr uint32_t t[8] = {2, 0, 0, 0, 0, 0, 0, 0}; return t[i]; }
};
struct CurveVesta {
    using Base = Fe<PallasFq>; using Scalar = Fe<PallasFp>;
    static constexpr int ID = 3; static constexpr bool GEN_X_NEG1 = true;
    LURK_HD static constexpr uint32_t GEN_Y(int i) { constexpr uint32_t t[8] = {2, 0, 0, 0, 0, 0, 0, 0}; return t[i]; }
};

template <class C>
inline Affine<typename C::Base> curve_generator() {
    using F = typename C::Base;
    Affine<F> g;
    g.x = C::GEN_X_NEG1 ? F::one().neg() : F::one();
    F y;
    for (int i = 0; i < 8; i++) y.v[i] = C::GEN_Y(i);
    g.y = F::from_canonical(y);
    return g;
}

}  // namespace lurk
