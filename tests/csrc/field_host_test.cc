// Host build of lurk-beta_b200/csrc/field.cuh + curve.cuh (carry flag emulated): lets the CPU test-suite
// exercise the exact limb algorithms the GPU kernels use.  Test-only helper, not part of the product.
#include "field.cuh"
#include <cstring>
using namespace lurk;

template <class F>
static int run(int op, const uint8_t *a, const uint8_t *b, uint8_t *out) {
    F x, y, r;
    memcpy(x.v, a, 32); memcpy(y.v, b, 32);
    if (!x.is_reduced() || !y.is_reduced()) return -1;
    x = F::from_canonical(x); y = F::from_canonical(y);
    switch (op) {
        case 0: r = x * y; break;
        case 1: r = x + y; break;
        case 2: r = x - y; break;
        case 3: r = x.inv(); break;
        case 4: r = x.neg(); break;
        case 5: r = x.pow5(); break;
        case 6: r = x.sqr(); break;
        case 7: r = x.inv_vartime(); break;
        default: return -2;
    }
    r = r.to_canonical();
    memcpy(out, r.v, 32);
    return 0;
}
extern "C" int fe_test_op(int field, int op, const uint8_t *a, const uint8_t *b, uint8_t *out) {
    switch (field) {
        case 0: return run<Fe<Bn254Fr>>(op, a, b, out);
        case 1: return run<Fe<Bn254Fq>>(op, a, b, out);
        case 2: return run<Fe<PallasFq>>(op, a, b, out);
        case 3: return run<Fe<PallasFp>>(op, a, b, out);
    }
    return -3;
}

template <class F>
static int run_dot(int k, const uint8_t *a, const uint8_t *b, uint8_t *out) {
    WideAcc<typename F::Params> acc;
    acc.clear();
    for (int i = 0; i < k; i++) {
        F x, y;
        memcpy(x.v, a + 32 * i, 32); memcpy(y.v, b + 32 * i, 32);
        acc.mul_acc(F::from_canonical(x), F::from_canonical(y));
    }
    F r = (k > 11 ? acc.template reduce<4>() : acc.reduce()).to_canonical();
    memcpy(out, r.v, 32);
    return 0;
}
// Montgomery-form dot product of k <= 15 pairs through the lazy accumulator
extern "C" int fe_test_dot(int field, int k, const uint8_t *a, const uint8_t *b, uint8_t *out) {
    switch (field) {
        case 0: return run_dot<Fe<Bn254Fr>>(k, a, b, out);
        case 1: return run_dot<Fe<Bn254Fq>>(k, a, b, out);
        case 2: return run_dot<Fe<PallasFq>>(k, a, b, out);
        case 3: return run_dot<Fe<PallasFp>>(k, a, b, out);
    }
    return -3;
}

// the 2^s-th root of unity the NTT kernels start from (Params::ROOT), canonical
template <class F>
static int run_root(uint8_t *out) {
    F w;
    for (int i = 0; i < 8; i++) w.v[i] = F::Params::ROOT(i);
    w = w.to_canonical();
    memcpy(out, w.v, 32);
    return F::Params::TWO_ADICITY;
}
extern "C" int fe_test_root(int field, uint8_t *out) {
    switch (field) {
        case 0: return run_root<Fe<Bn254Fr>>(out);
        case 1: return run_root<Fe<Bn254Fq>>(out);
        case 2: return run_root<Fe<PallasFq>>(out);
        case 3: return run_root<Fe<PallasFp>>(out);
    }
    return -3;
}
