/*
 * ORACLE -- test infrastructure, NOT product code.
 *
 * Plain-C CPU restatement of the lurk-beta proving hot path, used only as the
 * checker in tests/, in __graft_entry__.smoke() and as bench.py's cpu_baseline / --impl reference arm.
 * Nothing under lurk-beta_b200/ may link, import or call this file.
 *
 * The reference is Rust and its arithmetic lives in un-vendored git dependencies (neptune@dev,
 * arecibo@dev, bellpepper-core 0.4, pasta_curves 0.5, halo2curves 0.6 -- Cargo.toml:32,42,68,120-131),
 * no Rust toolchain exists here, so this is a restatement of the published algorithms anchored on the
 * reference's call sites and golden vectors:
 *   - Poseidon digest          src/hash.rs:180-203 (PoseidonCache::hash3/4/6/8 -> neptune Poseidon::hash)
 *   - Poseidon slot witness    src/lem/circuit.rs:212-315 (allocate_slot -> neptune circuit2 witness)
 *   - bit-decomposition slot   src/lem/circuit.rs:241-243 (AllocatedNum::to_bits_le_strict)
 *   - DAG hydration            src/lem/store_core.rs:199-269 + src/lem/store.rs:29-78 (preimage layouts)
 *   - Pedersen commit (MSM)    Arecibo vartime_multiscalar_mul; call sites src/proof/nova.rs:287,292
 *   - fold helpers             Arecibo NIFS::prove / commit_T / fold
 *   - NTT                      no call site in the reference: parity unpinned
 * Pinning: Poseidon digests are pinned by goldens G1..G27 (tests/test_oracle_golden.py); witness aux
 * counts and bit-decomp sizes are pinned by src/lem/multiframe.rs:495-516,991-1016; witness aux ORDER,
 * MSM outputs and NTT are "parity unpinned" (cross-checked only against oracle/spec.py, an independent
 * from-spec Python restatement).
 *
 * Poseidon constants are not generated here: oracle/spec.py generates them (Grain LFSR, Cauchy MDS,
 * Neptune's optimised forms) and installs them with oracle_set_poseidon_params().
 *
 * Arithmetic: 4x64-bit Montgomery, unsigned __int128, OpenMP across independent units.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned __int128 u128;
typedef struct { uint64_t l[4]; } fe;

typedef struct {
    uint64_t p[4];
    uint64_t inv;       /* -p^-1 mod 2^64 */
    fe r, r2;           /* R mod p, R^2 mod p */
    int nbits;
} fctx;

static const char *MOD_HEX[4] = {
    "30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001", /* 0 bn254 Fr */
    "30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47", /* 1 bn254 Fq */
    "40000000000000000000000000000000224698fc0994a8dd8c46eb2100000001", /* 2 pallas scalar Fq */
    "40000000000000000000000000000000224698fc094cf91b992d30ed00000001", /* 3 pallas base Fp */
};
static const int MOD_BITS[4] = {254, 254, 255, 255};
static fctx F[4];
static int f_ready = 0;

/* ------------------------------------------------------------------ field */
static inline int ge(const uint64_t *a, const uint64_t *b) {
    for (int i = 3; i >= 0; i--) { if (a[i] > b[i]) return 1; if (a[i] < b[i]) return 0; }
    return 1;
}
static inline uint64_t sub_n(uint64_t *r, const uint64_t *a, const uint64_t *b) {
    uint64_t br = 0;
    for (int i = 0; i < 4; i++) { u128 d = (u128)a[i] - b[i] - br; r[i] = (uint64_t)d; br = (uint64_t)(d >> 64) & 1; }
    return br;
}
static inline uint64_t add_n(uint64_t *r, const uint64_t *a, const uint64_t *b) {
    uint64_t c = 0;
    for (int i = 0; i < 4; i++) { u128 s = (u128)a[i] + b[i] + c; r[i] = (uint64_t)s; c = (uint64_t)(s >> 64); }
    return c;
}
static inline void f_add(const fctx *f, fe *r, const fe *a, const fe *b) {
    uint64_t c = add_n(r->l, a->l, b->l);
    if (c || ge(r->l, f->p)) sub_n(r->l, r->l, f->p);
}
static inline void f_sub(const fctx *f, fe *r, const fe *a, const fe *b) {
    if (sub_n(r->l, a->l, b->l)) add_n(r->l, r->l, f->p);
}
static inline int f_is_zero(const fe *a) { return (a->l[0] | a->l[1] | a->l[2] | a->l[3]) == 0; }
static inline int f_eq(const fe *a, const fe *b) { return memcmp(a, b, 32) == 0; }
static inline void f_neg(const fctx *f, fe *r, const fe *a) {
    if (f_is_zero(a)) { *r = *a; return; }
    sub_n(r->l, f->p, a->l);
}
/* CIOS Montgomery product */
static inline void f_mul(const fctx *f, fe *r, const fe *a, const fe *b) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        uint64_t c = 0;
        for (int j = 0; j < 4; j++) {
            u128 s = (u128)a->l[j] * b->l[i] + t[j] + c;
            t[j] = (uint64_t)s; c = (uint64_t)(s >> 64);
        }
        u128 s = (u128)t[4] + c; t[4] = (uint64_t)s; t[5] = (uint64_t)(s >> 64);
        uint64_t m = t[0] * f->inv;
        s = (u128)m * f->p[0] + t[0]; c = (uint64_t)(s >> 64);
        for (int j = 1; j < 4; j++) {
            s = (u128)m * f->p[j] + t[j] + c;
            t[j - 1] = (uint64_t)s; c = (uint64_t)(s >> 64);
        }
        s = (u128)t[4] + c; t[3] = (uint64_t)s; t[4] = t[5] + (uint64_t)(s >> 64);
    }
    if (t[4] || ge(t, f->p)) sub_n(r->l, t, f->p); else memcpy(r->l, t, 32);
}
static inline void f_sqr(const fctx *f, fe *r, const fe *a) { f_mul(f, r, a, a); }
static void f_from_raw(const fctx *f, fe *r, const uint64_t raw[4]) { fe t; memcpy(t.l, raw, 32); f_mul(f, r, &t, &f->r2); }
static void f_to_raw(const fctx *f, uint64_t raw[4], const fe *a) { fe one = {{1, 0, 0, 0}}, t; f_mul(f, &t, a, &one); memcpy(raw, t.l, 32); }
static void f_pow(const fctx *f, fe *r, const fe *a, const uint64_t e[4]) {
    fe acc = f->r, base = *a;
    for (int i = 0; i < 256; i++) {
        if ((e[i / 64] >> (i % 64)) & 1) f_mul(f, &acc, &acc, &base);
        f_sqr(f, &base, &base);
    }
    *r = acc;
}
static void f_inv(const fctx *f, fe *r, const fe *a) {
    uint64_t e[4], two[4] = {2, 0, 0, 0};
    sub_n(e, f->p, two);
    f_pow(f, r, a, e);
}
static int raw_reduced(const fctx *f, const uint64_t raw[4]) { return !ge(raw, f->p); }

static void init_fields(void) {
    if (f_ready) return;
    for (int k = 0; k < 4; k++) {
        fctx *f = &F[k];
        for (int i = 0; i < 4; i++) {
            char buf[17]; memcpy(buf, MOD_HEX[k] + 16 * (3 - i), 16); buf[16] = 0;
            f->p[i] = strtoull(buf, NULL, 16);
        }
        f->nbits = MOD_BITS[k];
        uint64_t inv = 1;
        for (int i = 0; i < 6; i++) inv *= 2 - f->p[0] * inv;   /* Newton: p^-1 mod 2^64 */
        f->inv = (uint64_t)(0 - inv);
        /* R mod p by 256 doublings of 1; R^2 by 256 more */
        uint64_t x[4] = {1, 0, 0, 0};
        for (int i = 0; i < 512; i++) {
            uint64_t c = add_n(x, x, x);
            if (c || ge(x, f->p)) sub_n(x, x, f->p);
            if (i == 255) memcpy(f->r.l, x, 32);
        }
        memcpy(f->r2.l, x, 32);
    }
    f_ready = 1;
}

int oracle_field_selftest(int field_id) {
    init_fields();
    const fctx *f = &F[field_id];
    fe a, b, c, d;
    uint64_t ra[4] = {0x123456789abcdefULL, 77, 0xdeadbeef, 0x1fffffff}, rb[4] = {5, 6, 7, 8}, out[4];
    f_from_raw(f, &a, ra); f_from_raw(f, &b, rb);
    f_mul(f, &c, &a, &b); f_inv(f, &d, &b); f_mul(f, &c, &c, &d);
    f_to_raw(f, out, &c);
    return memcmp(out, ra, 32) == 0 ? 0 : -1;
}

/* ------------------------------------------------------------------ Poseidon */
typedef struct {
    int t, rf, rp, ready;
    fe domain_tag;
    fe *rc;          /* t*(rf+rp) textbook round constants */
    fe *mds;         /* t*t row-major */
    fe *comp;        /* t*rf+rp compressed constants */
    fe *pre;         /* t*t pre-sparse */
    fe *sp_w;        /* rp * t      first column of each sparse matrix */
    fe *sp_v;        /* rp * (t-1)  first row (rest) of each sparse matrix */
} pparams;
static pparams PP[4][9];

/* all arrays: canonical 32-byte little-endian elements, installed by oracle/spec.py */
int oracle_set_poseidon_params(int field_id, int arity, int rf, int rp, const uint8_t *domain_tag,
                               const uint8_t *rc, const uint8_t *mds, const uint8_t *comp,
                               const uint8_t *pre, const uint8_t *sp_w, const uint8_t *sp_v) {
    init_fields();
    if (field_id < 0 || field_id > 3 || arity < 1 || arity > 8) return -1;
    const fctx *f = &F[field_id];
    pparams *P = &PP[field_id][arity];
    int t = arity + 1;
    P->t = t; P->rf = rf; P->rp = rp;
#define LOADV(dst, src, cnt) do { free(dst); dst = malloc(sizeof(fe) * (cnt)); \
        for (int i_ = 0; i_ < (cnt); i_++) { uint64_t raw_[4]; memcpy(raw_, (src) + 32 * i_, 32); f_from_raw(f, &dst[i_], raw_); } } while (0)
    uint64_t raw[4]; memcpy(raw, domain_tag, 32); f_from_raw(f, &P->domain_tag, raw);
    LOADV(P->rc, rc, t * (rf + rp));
    LOADV(P->mds, mds, t * t);
    LOADV(P->comp, comp, t * rf + rp);
    LOADV(P->pre, pre, t * t);
    LOADV(P->sp_w, sp_w, rp * t);
    LOADV(P->sp_v, sp_v, rp * (t - 1));
    P->ready = 1;
    return 0;
}

static inline void sbox5(const fctx *f, fe *x) { fe x2, x4; f_sqr(f, &x2, x); f_sqr(f, &x4, &x2); f_mul(f, x, &x4, x); }

static void vec_mat(const fctx *f, int t, fe *s, const fe *m) {
    fe out[9];
    for (int j = 0; j < t; j++) {
        fe acc = {{0, 0, 0, 0}}, tmp;
        for (int i = 0; i < t; i++) { f_mul(f, &tmp, &s[i], &m[i * t + j]); f_add(f, &acc, &acc, &tmp); }
        out[j] = acc;
    }
    memcpy(s, out, sizeof(fe) * t);
}

/* textbook Poseidon (ARK, S-box, MDS) -- state in Montgomery form */
static void permute_correct(const fctx *f, const pparams *P, fe *s) {
    int t = P->t, half = P->rf / 2;
    for (int r = 0; r < P->rf + P->rp; r++) {
        for (int i = 0; i < t; i++) f_add(f, &s[i], &s[i], &P->rc[r * t + i]);
        if (r < half || r >= half + P->rp) for (int i = 0; i < t; i++) sbox5(f, &s[i]);
        else sbox5(f, &s[0]);
        vec_mat(f, t, s, P->mds);
    }
}

/* Neptune optimised schedule; if aux != NULL writes 3 values per S-box (Montgomery form) */
static void permute_optimised(const fctx *f, const pparams *P, fe *s, fe *aux) {
    int t = P->t, half = P->rf / 2, k = 0;
    const fe *c = P->comp;
    fe zero = {{0, 0, 0, 0}};
#define SBOX(x, key) do { fe x2_, x4_; f_sqr(f, &x2_, &(x)); f_sqr(f, &x4_, &x2_); f_mul(f, &(x), &x4_, &(x)); \
        f_add(f, &(x), &(x), (key)); if (aux) { aux[0] = x2_; aux[1] = x4_; aux[2] = (x); aux += 3; } } while (0)
    for (int i = 0; i < t; i++) f_add(f, &s[i], &s[i], &c[k + i]);
    k += t;
    for (int r = 0; r < half; r++) {
        for (int i = 0; i < t; i++) SBOX(s[i], &c[k + i]);
        k += t;
        vec_mat(f, t, s, r == half - 1 ? P->pre : P->mds);
    }
    for (int r = 0; r < P->rp; r++) {
        SBOX(s[0], &c[k]); k++;
        const fe *w = &P->sp_w[r * t], *v = &P->sp_v[r * (t - 1)];
        fe s0 = {{0, 0, 0, 0}}, tmp;
        for (int i = 0; i < t; i++) { f_mul(f, &tmp, &s[i], &w[i]); f_add(f, &s0, &s0, &tmp); }
        for (int j = 1; j < t; j++) { f_mul(f, &tmp, &s[0], &v[j - 1]); f_add(f, &s[j], &s[j], &tmp); }
        s[0] = s0;
    }
    for (int r = 0; r < half - 1; r++) {
        for (int i = 0; i < t; i++) SBOX(s[i], &c[k + i]);
        k += t;
        vec_mat(f, t, s, P->mds);
    }
    for (int i = 0; i < t; i++) SBOX(s[i], &zero);
    vec_mat(f, t, s, P->mds);
#undef SBOX
}

static int load_state(const fctx *f, const pparams *P, fe *s, const uint8_t *pre) {
    s[0] = P->domain_tag;
    for (int i = 1; i < P->t; i++) {
        uint64_t raw[4]; memcpy(raw, pre + 32 * (i - 1), 32);
        if (!raw_reduced(f, raw)) return -1;
        f_from_raw(f, &s[i], raw);
    }
    return 0;
}

/* mode 0: textbook rounds, mode 1: optimised rounds.  Inputs/outputs canonical 32-byte LE. */
int oracle_poseidon_hash_batch(int field_id, int arity, const uint8_t *pre, size_t n, uint8_t *out,
                               int mode, int nthreads) {
    init_fields();
    const fctx *f = &F[field_id];
    const pparams *P = &PP[field_id][arity];
    if (!P->ready) return -2;
    int bad = 0;
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1)
    for (size_t h = 0; h < n; h++) {
        fe s[9];
        if (load_state(f, P, s, pre + h * 32 * arity)) { bad = 1; continue; }
        if (mode == 0) permute_correct(f, P, s); else permute_optimised(f, P, s, NULL);
        uint64_t raw[4]; f_to_raw(f, raw, &s[1]);
        memcpy(out + 32 * h, raw, 32);
    }
    return bad ? -3 : 0;
}

/* slot block = preimage (A) | aux (3*(t*rf+rp)) | digest (1), canonical LE (src/lem/circuit.rs:264-299) */
int oracle_poseidon_witness_batch(int field_id, int arity, const uint8_t *pre, size_t n, uint8_t *out,
                                  int nthreads) {
    init_fields();
    const fctx *f = &F[field_id];
    const pparams *P = &PP[field_id][arity];
    if (!P->ready) return -2;
    int naux = 3 * (P->t * P->rf + P->rp);
    size_t block = (size_t)(arity + naux + 1) * 32;
    int bad = 0;
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1)
    for (size_t h = 0; h < n; h++) {
        fe s[9];
        fe *aux = malloc(sizeof(fe) * naux);
        if (load_state(f, P, s, pre + h * 32 * arity)) { bad = 1; free(aux); continue; }
        permute_optimised(f, P, s, aux);
        uint8_t *o = out + h * block;
        memcpy(o, pre + h * 32 * arity, 32 * arity);
        o += 32 * arity;
        for (int i = 0; i < naux; i++) { uint64_t raw[4]; f_to_raw(f, raw, &aux[i]); memcpy(o + 32 * i, raw, 32); }
        uint64_t raw[4]; f_to_raw(f, raw, &s[1]);
        memcpy(o + 32 * naux, raw, 32);
        free(aux);
    }
    return bad ? -3 : 0;
}

/* ------------------------------------------------------------------ bit decomposition slot */
/* aux order of bellpepper-core to_bits_le_strict preceded by the allocated preimage element */
int oracle_bitdecomp_size(int field_id) {
    init_fields();
    const fctx *f = &F[field_id];
    uint64_t b[4], one[4] = {1, 0, 0, 0};
    sub_n(b, f->p, one);
    int cnt = 1, found = 0, run = 0, have_last = 0;
    for (int i = 255; i >= 0; i--) {
        int bb = (b[i / 64] >> (i % 64)) & 1;
        found |= bb;
        if (!found) continue;
        if (bb) { cnt++; run++; }
        else {
            if (run) { cnt += run + have_last - 1; have_last = 1; run = 0; }
            cnt++;
        }
    }
    return cnt;
}

int oracle_bitdecomp_witness_batch(int field_id, const uint8_t *vals, size_t n, uint8_t *out, int nthreads) {
    init_fields();
    const fctx *f = &F[field_id];
    int size = oracle_bitdecomp_size(field_id);
    uint64_t b[4], one[4] = {1, 0, 0, 0};
    sub_n(b, f->p, one);
    int bad = 0;
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1)
    for (size_t h = 0; h < n; h++) {
        uint64_t x[4]; memcpy(x, vals + 32 * h, 32);
        if (!raw_reduced(f, x)) { bad = 1; continue; }
        uint8_t *o = out + h * (size_t)size * 32;
        memset(o, 0, (size_t)size * 32);
        memcpy(o, x, 32);
        int k = 1, found = 0, nrun = 0, have_last = 0, last = 0;
        int runbits[256];
        for (int i = 255; i >= 0; i--) {
            int bb = (b[i / 64] >> (i % 64)) & 1, ab = (x[i / 64] >> (i % 64)) & 1;
            found |= bb;
            if (!found) continue;
            if (bb) { o[32 * k++] = (uint8_t)ab; runbits[nrun++] = ab; }
            else {
                if (nrun) {
                    if (have_last) runbits[nrun++] = last;
                    int cur = runbits[0];
                    for (int j = 1; j < nrun; j++) { cur &= runbits[j]; o[32 * k++] = (uint8_t)cur; }
                    last = cur; have_last = 1; nrun = 0;
                }
                o[32 * k++] = (uint8_t)ab;
            }
        }
        if (k != size) bad = 1;
    }
    return bad ? -3 : 0;
}

/* ------------------------------------------------------------------ DAG hydration */
/* Node record mirrors include/lurk_b200.h lurk_dag_node: kind 2/3/4 = tupleN (children as (tag,digest)
 * pairs -> H4/H6/H8), kind 5 = compact (H4 [d0, tag1, d1, d2], src/lem/store.rs:75-77), kind 6 =
 * commitment (H3 [secret=d0, tag1, d1], src/lem/store.rs:70-73).  child index < n_atoms refers to an atom
 * digest, otherwise to node (index - n_atoms).  Nodes must be topologically ordered (children first). */
typedef struct { uint8_t kind; uint8_t pad; uint16_t tag[4]; uint32_t child[4]; } dag_node;

int oracle_dag_hash(int field_id, const dag_node *nodes, size_t n, const uint8_t *atoms, size_t n_atoms,
                    uint8_t *out) {
    init_fields();
    const fctx *f = &F[field_id];
    for (size_t i = 0; i < n; i++) {
        const dag_node *nd = &nodes[i];
        uint8_t pre[8 * 32];
        memset(pre, 0, sizeof pre);
        int arity = 0;
        const uint8_t *d[4];
        int nch = nd->kind == 2 ? 2 : nd->kind == 3 ? 3 : nd->kind == 4 ? 4 : nd->kind == 5 ? 3 : nd->kind == 6 ? 2 : -1;
        if (nch < 0) return -1;
        for (int c = 0; c < nch; c++) {
            uint32_t ix = nd->child[c];
            if (ix < n_atoms) d[c] = atoms + 32 * (size_t)ix;
            else { if (ix - n_atoms >= i) return -4; d[c] = out + 32 * (size_t)(ix - n_atoms); }
        }
        if (nd->kind <= 4) {
            arity = 2 * nch;
            for (int c = 0; c < nch; c++) { memcpy(pre + 64 * c, &nd->tag[c], 2); memcpy(pre + 64 * c + 32, d[c], 32); }
        } else if (nd->kind == 5) {
            arity = 4;
            memcpy(pre, d[0], 32); memcpy(pre + 32, &nd->tag[1], 2); memcpy(pre + 64, d[1], 32); memcpy(pre + 96, d[2], 32);
        } else {
            arity = 3;
            memcpy(pre, d[0], 32); memcpy(pre + 32, &nd->tag[1], 2); memcpy(pre + 64, d[1], 32);
        }
        const pparams *P = &PP[field_id][arity];
        if (!P->ready) return -2;
        fe s[9];
        if (load_state(f, P, s, pre)) return -3;
        permute_optimised(f, P, s, NULL);
        uint64_t raw[4]; f_to_raw(f, raw, &s[1]);
        memcpy(out + 32 * i, raw, 32);
    }
    return 0;
}

/* ------------------------------------------------------------------ curves: Jacobian, a = 0 */
typedef struct { fe x, y, z; } jac;   /* z == 0 -> identity */
static const int CURVE_BASE[4] = {1, 0, 3, 2};
static const int CURVE_SCALAR[4] = {0, 1, 2, 3};

static void j_dbl(const fctx *f, jac *r, const jac *p) {
    if (f_is_zero(&p->z) || f_is_zero(&p->y)) { memset(r, 0, sizeof *r); return; }
    fe a, b, c, d, e, g, t, x3, y3, z3;
    f_sqr(f, &a, &p->x); f_sqr(f, &b, &p->y); f_sqr(f, &c, &b);
    f_add(f, &t, &p->x, &b); f_sqr(f, &t, &t); f_sub(f, &t, &t, &a); f_sub(f, &t, &t, &c); f_add(f, &d, &t, &t);
    f_add(f, &e, &a, &a); f_add(f, &e, &e, &a);
    f_sqr(f, &g, &e);
    f_sub(f, &x3, &g, &d); f_sub(f, &x3, &x3, &d);
    f_mul(f, &z3, &p->y, &p->z); f_add(f, &z3, &z3, &z3);
    f_sub(f, &t, &d, &x3); f_mul(f, &y3, &e, &t);
    f_add(f, &c, &c, &c); f_add(f, &c, &c, &c); f_add(f, &c, &c, &c);
    f_sub(f, &y3, &y3, &c);
    r->x = x3; r->y = y3; r->z = z3;
}
static void j_add(const fctx *f, jac *r, const jac *p, const jac *q) {
    if (f_is_zero(&p->z)) { *r = *q; return; }
    if (f_is_zero(&q->z)) { *r = *p; return; }
    fe z1z1, z2z2, u1, u2, s1, s2, h, rr, t, hh, hhh, v, x3, y3, z3;
    f_sqr(f, &z1z1, &p->z); f_sqr(f, &z2z2, &q->z);
    f_mul(f, &u1, &p->x, &z2z2); f_mul(f, &u2, &q->x, &z1z1);
    f_mul(f, &t, &q->z, &z2z2); f_mul(f, &s1, &p->y, &t);
    f_mul(f, &t, &p->z, &z1z1); f_mul(f, &s2, &q->y, &t);
    f_sub(f, &h, &u2, &u1); f_sub(f, &rr, &s2, &s1);
    if (f_is_zero(&h)) { if (f_is_zero(&rr)) { j_dbl(f, r, p); } else memset(r, 0, sizeof *r); return; }
    f_sqr(f, &hh, &h); f_mul(f, &hhh, &hh, &h); f_mul(f, &v, &u1, &hh);
    f_sqr(f, &x3, &rr); f_sub(f, &x3, &x3, &hhh); f_sub(f, &x3, &x3, &v); f_sub(f, &x3, &x3, &v);
    f_sub(f, &t, &v, &x3); f_mul(f, &y3, &rr, &t); f_mul(f, &t, &s1, &hhh); f_sub(f, &y3, &y3, &t);
    f_mul(f, &z3, &p->z, &q->z); f_mul(f, &z3, &z3, &h);
    r->x = x3; r->y = y3; r->z = z3;
}
static void j_from_affine(const fctx *f, jac *r, const fe *x, const fe *y) {
    if (f_is_zero(x) && f_is_zero(y)) { memset(r, 0, sizeof *r); return; }
    r->x = *x; r->y = *y; r->z = f->r;
}
static void j_neg(const fctx *f, jac *r, const jac *p) { *r = *p; f_neg(f, &r->y, &p->y); }
/* out: x | y | z with z in {0,1}; identity = all zero.  canonical LE */
static void j_to_affine_bytes(const fctx *f, uint8_t out[96], const jac *p) {
    memset(out, 0, 96);
    if (f_is_zero(&p->z)) return;
    fe zi, zi2, zi3, x, y; uint64_t raw[4];
    f_inv(f, &zi, &p->z); f_sqr(f, &zi2, &zi); f_mul(f, &zi3, &zi2, &zi);
    f_mul(f, &x, &p->x, &zi2); f_mul(f, &y, &p->y, &zi3);
    f_to_raw(f, raw, &x); memcpy(out, raw, 32);
    f_to_raw(f, raw, &y); memcpy(out + 32, raw, 32);
    out[64] = 1;
}

static int load_bases(const fctx *f, const uint8_t *bases, size_t n, fe *bx, fe *by) {
    int bad = 0;
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        uint64_t rx[4], ry[4];
        memcpy(rx, bases + 64 * i, 32); memcpy(ry, bases + 64 * i + 32, 32);
        if (!raw_reduced(f, rx) || !raw_reduced(f, ry)) bad = 1;
        f_from_raw(f, &bx[i], rx); f_from_raw(f, &by[i], ry);
    }
    return bad;
}

/* naive: sum_i [k_i]P_i by double-and-add */
int oracle_msm_naive(int curve_id, const uint8_t *bases, const uint8_t *scalars, size_t n, uint8_t out[96]) {
    init_fields();
    const fctx *f = &F[CURVE_BASE[curve_id]];
    fe *bx = malloc(sizeof(fe) * (n + 1)), *by = malloc(sizeof(fe) * (n + 1));
    if (load_bases(f, bases, n, bx, by)) { free(bx); free(by); return -3; }
    jac acc; memset(&acc, 0, sizeof acc);
    for (size_t i = 0; i < n; i++) {
        uint64_t k[4]; memcpy(k, scalars + 32 * i, 32);
        jac p, r; j_from_affine(f, &p, &bx[i], &by[i]); memset(&r, 0, sizeof r);
        for (int b = 255; b >= 0; b--) {
            j_dbl(f, &r, &r);
            if ((k[b / 64] >> (b % 64)) & 1) j_add(f, &r, &r, &p);
        }
        j_add(f, &acc, &acc, &r);
    }
    j_to_affine_bytes(f, out, &acc);
    free(bx); free(by);
    return 0;
}

/* Pippenger, unsigned windows; OpenMP over (window, chunk-of-points) tasks so that all host threads are busy
 * (the shape of a multi-threaded CPU vartime MSM).  Window width minimises nwin * (n + chunks * 2^(c+1)). */
int oracle_msm_pippenger(int curve_id, const uint8_t *bases, const uint8_t *scalars, size_t n, uint8_t out[96],
                         int nthreads) {
    init_fields();
    const fctx *f = &F[CURVE_BASE[curve_id]];
    if (nthreads < 1) nthreads = 1;
    int c = 4, chunks = 1;
    double best = -1;
    for (int cc = 3; cc <= 16; cc++) {
        int nw = (255 + cc - 1) / cc;
        int ch = (nthreads + nw - 1) / nw;
        if (ch < 1) ch = 1;
        if ((size_t)ch > n / 64 + 1) ch = (int)(n / 64 + 1);
        double cost = (double)nw * ((double)n + (double)ch * (double)((size_t)2 << cc));
        if (best < 0 || cost < best) { best = cost; c = cc; chunks = ch; }
    }
    int nwin = (255 + c - 1) / c;
    fe *bx = malloc(sizeof(fe) * (n + 1)), *by = malloc(sizeof(fe) * (n + 1));
    if (load_bases(f, bases, n, bx, by)) { free(bx); free(by); return -3; }
    int ntasks = nwin * chunks;
    jac *part = calloc(ntasks, sizeof(jac));
    size_t per = (n + chunks - 1) / chunks;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
    for (int task = 0; task < ntasks; task++) {
        int w = task / chunks, ch = task % chunks;
        size_t lo = (size_t)ch * per, hi = lo + per < n ? lo + per : n;
        size_t nb = ((size_t)1 << c) - 1;
        jac *bk = calloc(nb, sizeof(jac));
        for (size_t i = lo; i < hi; i++) {
            uint64_t k[4]; memcpy(k, scalars + 32 * i, 32);
            int bit = w * c;
            uint64_t d = k[bit / 64] >> (bit % 64);
            if ((bit % 64) + c > 64 && bit / 64 < 3) d |= k[bit / 64 + 1] << (64 - bit % 64);
            d &= ((uint64_t)1 << c) - 1;
            if (!d) continue;
            jac p; j_from_affine(f, &p, &bx[i], &by[i]);
            j_add(f, &bk[d - 1], &bk[d - 1], &p);
        }
        jac run, sum; memset(&run, 0, sizeof run); memset(&sum, 0, sizeof sum);
        for (size_t b = nb; b-- > 0;) { j_add(f, &run, &run, &bk[b]); j_add(f, &sum, &sum, &run); }
        part[task] = sum;
        free(bk);
    }
    jac acc; memset(&acc, 0, sizeof acc);
    for (int w = nwin - 1; w >= 0; w--) {
        for (int i = 0; i < c; i++) j_dbl(f, &acc, &acc);
        for (int ch = 0; ch < chunks; ch++) j_add(f, &acc, &acc, &part[w * chunks + ch]);
    }
    j_to_affine_bytes(f, out, &acc);
    free(bx); free(by); free(part);
    return 0;
}

/* bases_out[i] = [start + i + 1] G, affine canonical (running sum from the generator) */
int oracle_gen_bases(int curve_id, const uint8_t gen[64], uint64_t start, size_t n, uint8_t *bases_out) {
    init_fields();
    const fctx *f = &F[CURVE_BASE[curve_id]];
    uint64_t rx[4], ry[4]; memcpy(rx, gen, 32); memcpy(ry, gen + 32, 32);
    fe gx, gy; f_from_raw(f, &gx, rx); f_from_raw(f, &gy, ry);
    jac g, acc; j_from_affine(f, &g, &gx, &gy);
    /* acc = [start+1] G */
    memset(&acc, 0, sizeof acc);
    uint64_t k = start + 1;
    for (int b = 63; b >= 0; b--) { j_dbl(f, &acc, &acc); if ((k >> b) & 1) j_add(f, &acc, &acc, &g); }
    /* batches of running sums, normalised with one inversion per batch (Montgomery trick) */
    const size_t B = 1024;
    jac *pts = malloc(sizeof(jac) * B);
    fe *pref = malloc(sizeof(fe) * B);
    for (size_t base = 0; base < n; base += B) {
        size_t m = n - base < B ? n - base : B;
        for (size_t i = 0; i < m; i++) { pts[i] = acc; j_add(f, &acc, &acc, &g); }
        fe run = f->r;
        for (size_t i = 0; i < m; i++) { pref[i] = run; f_mul(f, &run, &run, &pts[i].z); }
        fe inv; f_inv(f, &inv, &run);
        for (size_t i = m; i-- > 0;) {
            fe zi, zi2, zi3, x, y; uint64_t raw[4];
            f_mul(f, &zi, &inv, &pref[i]); f_mul(f, &inv, &inv, &pts[i].z);
            f_sqr(f, &zi2, &zi); f_mul(f, &zi3, &zi2, &zi);
            f_mul(f, &x, &pts[i].x, &zi2); f_mul(f, &y, &pts[i].y, &zi3);
            f_to_raw(f, raw, &x); memcpy(bases_out + 64 * (base + i), raw, 32);
            f_to_raw(f, raw, &y); memcpy(bases_out + 64 * (base + i) + 32, raw, 32);
        }
    }
    free(pts); free(pref);
    return 0;
}

/* sum of affine points (x|y|z flag records of 96 bytes) -> affine record; used to check the N-GPU combine */
int oracle_point_sum(int curve_id, const uint8_t *pts96, size_t n, uint8_t out[96]) {
    init_fields();
    const fctx *f = &F[CURVE_BASE[curve_id]];
    jac acc; memset(&acc, 0, sizeof acc);
    for (size_t i = 0; i < n; i++) {
        if (!pts96[96 * i + 64]) continue;
        uint64_t rx[4], ry[4]; memcpy(rx, pts96 + 96 * i, 32); memcpy(ry, pts96 + 96 * i + 32, 32);
        fe x, y; f_from_raw(f, &x, rx); f_from_raw(f, &y, ry);
        jac p; j_from_affine(f, &p, &x, &y);
        j_add(f, &acc, &acc, &p);
    }
    j_to_affine_bytes(f, out, &acc);
    return 0;
}
void oracle_unused_neg(void) { (void)j_neg; (void)CURVE_SCALAR; (void)f_eq; }

/* ------------------------------------------------------------------ fold helpers (field vectors, canonical LE) */
/* out[i] = a[i] + r * b[i]   (Arecibo RelaxedR1CSWitness::fold: W <- W1 + r W2, E <- E1 + r T) */
int oracle_axpy(int field_id, const uint8_t *a, const uint8_t *b, const uint8_t r[32], size_t n, uint8_t *out,
                int nthreads) {
    init_fields();
    const fctx *f = &F[field_id];
    uint64_t raw[4]; memcpy(raw, r, 32);
    fe rm; f_from_raw(f, &rm, raw);
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1)
    for (size_t i = 0; i < n; i++) {
        uint64_t ra[4], rb[4], ro[4]; memcpy(ra, a + 32 * i, 32); memcpy(rb, b + 32 * i, 32);
        fe x, y; f_from_raw(f, &x, ra); f_from_raw(f, &y, rb);
        f_mul(f, &y, &y, &rm); f_add(f, &x, &x, &y);
        f_to_raw(f, ro, &x); memcpy(out + 32 * i, ro, 32);
    }
    return 0;
}
/* y = M z, CSR (row_ptr: rows+1 u32... u64, col: u32, val: canonical LE) */
int oracle_spmv(int field_id, const uint64_t *row_ptr, const uint32_t *col, const uint8_t *val, size_t rows,
                const uint8_t *z, uint8_t *y, int nthreads) {
    init_fields();
    const fctx *f = &F[field_id];
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1)
    for (size_t i = 0; i < rows; i++) {
        fe acc = {{0, 0, 0, 0}};
        for (uint64_t k = row_ptr[i]; k < row_ptr[i + 1]; k++) {
            uint64_t rv[4], rz[4]; memcpy(rv, val + 32 * k, 32); memcpy(rz, z + 32 * (size_t)col[k], 32);
            fe v, zz; f_from_raw(f, &v, rv); f_from_raw(f, &zz, rz);
            f_mul(f, &v, &v, &zz); f_add(f, &acc, &acc, &v);
        }
        uint64_t ro[4]; f_to_raw(f, ro, &acc); memcpy(y + 32 * i, ro, 32);
    }
    return 0;
}
/* T = az1*bz2 + az2*bz1 - u1*cz2 - u2*cz1   (Nova cross term) */
int oracle_cross_term(int field_id, const uint8_t *az1, const uint8_t *bz1, const uint8_t *cz1,
                      const uint8_t *az2, const uint8_t *bz2, const uint8_t *cz2,
                      const uint8_t u1[32], const uint8_t u2[32], size_t n, uint8_t *t_out, int nthreads) {
    init_fields();
    const fctx *f = &F[field_id];
    uint64_t raw[4]; fe U1, U2;
    memcpy(raw, u1, 32); f_from_raw(f, &U1, raw);
    memcpy(raw, u2, 32); f_from_raw(f, &U2, raw);
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1)
    for (size_t i = 0; i < n; i++) {
        fe a1, b1, c1, a2, b2, c2, t, s; uint64_t r_[4];
#define LD(dst, src) do { memcpy(r_, (src) + 32 * i, 32); f_from_raw(f, &dst, r_); } while (0)
        LD(a1, az1); LD(b1, bz1); LD(c1, cz1); LD(a2, az2); LD(b2, bz2); LD(c2, cz2);
#undef LD
        f_mul(f, &t, &a1, &b2); f_mul(f, &s, &a2, &b1); f_add(f, &t, &t, &s);
        f_mul(f, &s, &U1, &c2); f_sub(f, &t, &t, &s);
        f_mul(f, &s, &U2, &c1); f_sub(f, &t, &t, &s);
        f_to_raw(f, r_, &t); memcpy(t_out + 32 * i, r_, 32);
    }
    return 0;
}

/* ------------------------------------------------------------------ NTT (iterative radix-2, natural order) */
int oracle_ntt(int field_id, uint8_t *data, int log_n, const uint8_t root[32], int nthreads) {
    init_fields();
    const fctx *f = &F[field_id];
    size_t n = (size_t)1 << log_n;
    fe *a = malloc(sizeof(fe) * n);
    for (size_t i = 0; i < n; i++) { uint64_t raw[4]; memcpy(raw, data + 32 * i, 32); f_from_raw(f, &a[i], raw); }
    for (size_t i = 0; i < n; i++) {           /* bit reversal */
        size_t j = 0;
        for (int b = 0; b < log_n; b++) if (i & ((size_t)1 << b)) j |= (size_t)1 << (log_n - 1 - b);
        if (i < j) { fe t = a[i]; a[i] = a[j]; a[j] = t; }
    }
    uint64_t raw[4]; memcpy(raw, root, 32);
    fe w_n; f_from_raw(f, &w_n, raw);
    for (int s = 1; s <= log_n; s++) {
        size_t m = (size_t)1 << s, half = m >> 1;
        fe wm = w_n;
        for (int k = 0; k < log_n - s; k++) f_sqr(f, &wm, &wm);
        fe *tw = malloc(sizeof(fe) * half);
        tw[0] = f->r;
        for (size_t j = 1; j < half; j++) f_mul(f, &tw[j], &tw[j - 1], &wm);
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1)
        for (size_t k = 0; k < n / 2; k++) {
            size_t blk = k / half, j = k % half, i0 = blk * m + j, i1 = i0 + half;
            fe t, u = a[i0];
            f_mul(f, &t, &tw[j], &a[i1]);
            f_add(f, &a[i0], &u, &t); f_sub(f, &a[i1], &u, &t);
        }
        free(tw);
    }
    for (size_t i = 0; i < n; i++) { uint64_t r_[4]; f_to_raw(f, r_, &a[i]); memcpy(data + 32 * i, r_, 32); }
    free(a);
    return 0;
}

/* ------------------------------------------------------------------ N3: from_label / hash_to_curve (CPU port) */
/* Restates oracle/h2c.py in C (BLAKE2b-XMD hash_to_field, SVDW or SSWU + 3-isogeny, affine sum): the CPU baseline of the N3 row and a
 * third implementation beside the Python restatement and the CUDA templates.  Every curve constant arrives from the caller
 * (oracle/capi.py passes what oracle/h2c.py computes / derives): consts = canonical 32-byte elements,
 *   method 0 (SVDW):  b, Z, c1, c2, c3, c4            method 1 (SSWU): b, Z, iso_a, iso_b, iso[0..12] */
static const uint64_t B2B_IV[8] = {0x6a09e667f3bcc908ull, 0xbb67ae8584caa73bull, 0x3c6ef372fe94f82bull, 0xa54ff53a5f1d36f1ull,
                                   0x510e527fade682d1ull, 0x9b05688c2b3e6c1full, 0x1f83d9abfb41bd6bull, 0x5be0cd19137e2179ull};
static const uint8_t B2B_SIGMA[12][16] = {
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
    {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
    {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
    {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
    {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0},
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3}};
static inline uint64_t rotr64(uint64_t x, int n) { return (x >> n) | (x << (64 - n)); }
static void b2b_compress(uint64_t h[8], const uint8_t block[128], uint64_t t, int last) {
    uint64_t m[16], v[16];
    memcpy(m, block, 128);
    for (int i = 0; i < 8; i++) { v[i] = h[i]; v[i + 8] = B2B_IV[i]; }
    v[12] ^= t;
    if (last) v[14] = ~v[14];
#define B2B_G(a, b, c, d, x, y) \
    v[a] += v[b] + (x); v[d] = rotr64(v[d] ^ v[a], 32); v[c] += v[d]; v[b] = rotr64(v[b] ^ v[c], 24); \
    v[a] += v[b] + (y); v[d] = rotr64(v[d] ^ v[a], 16); v[c] += v[d]; v[b] = rotr64(v[b] ^ v[c], 63);
    for (int r = 0; r < 12; r++) {
        const uint8_t *s = B2B_SIGMA[r];
        B2B_G(0, 4, 8, 12, m[s[0]], m[s[1]]) B2B_G(1, 5, 9, 13, m[s[2]], m[s[3]]) B2B_G(2, 6, 10, 14, m[s[4]], m[s[5]]) B2B_G(3, 7, 11, 15, m[s[6]], m[s[7]])
        B2B_G(0, 5, 10, 15, m[s[8]], m[s[9]]) B2B_G(1, 6, 11, 12, m[s[10]], m[s[11]]) B2B_G(2, 7, 8, 13, m[s[12]], m[s[13]]) B2B_G(3, 4, 9, 14, m[s[14]], m[s[15]])
    }
#undef B2B_G
    for (int i = 0; i < 8; i++) h[i] ^= v[i] ^ v[i + 8];
}
/* BLAKE2b-512 of a message of at most 256 bytes */
static void b2b_512(const uint8_t *msg, size_t len, uint8_t out[64]) {
    uint64_t h[8];
    uint8_t blk[128];
    for (int i = 0; i < 8; i++) h[i] = B2B_IV[i];
    h[0] ^= 0x01010040ull;
    size_t off = 0;
    while (len - off > 128) { b2b_compress(h, msg + off, off + 128, 0); off += 128; }
    memset(blk, 0, 128);
    memcpy(blk, msg + off, len - off);
    b2b_compress(h, blk, len, 1);
    memcpy(out, h, 64);
}
/* 64 big-endian bytes -> field element (Montgomery): (hi 2^256 + lo) mod p */
static void fe_from_be64(const fctx *f, fe *r, const uint8_t d[64]) {
    uint64_t lo[4], hi[4];
    for (int w = 0; w < 4; w++) {
        uint64_t l = 0, h = 0;
        for (int b = 0; b < 8; b++) { l |= (uint64_t)d[63 - 8 * w - b] << (8 * b); h |= (uint64_t)d[31 - 8 * w - b] << (8 * b); }
        lo[w] = l; hi[w] = h;
    }
    while (ge(lo, f->p)) sub_n(lo, lo, f->p);
    while (ge(hi, f->p)) sub_n(hi, hi, f->p);
    fe lm, hm;
    f_from_raw(f, &lm, lo);
    f_from_raw(f, &hm, hi);
    f_mul(f, &hm, &hm, &f->r2);        /* times 2^256 */
    f_add(f, r, &lm, &hm);
}
typedef struct { int s; uint64_t t[4], t1h[4]; fe c; } sqrt_ctx;     /* p - 1 = 2^s t; t1h = (t + 1) / 2; c = z^t, z a non-square */
static void sqrt_init(const fctx *f, sqrt_ctx *q) {
    uint64_t one[4] = {1, 0, 0, 0}, pm1[4], half[4];
    sub_n(pm1, f->p, one);
    memcpy(q->t, pm1, 32);
    q->s = 0;
    while (!(q->t[0] & 1)) { for (int i = 0; i < 3; i++) q->t[i] = (q->t[i] >> 1) | (q->t[i + 1] << 63); q->t[3] >>= 1; q->s++; }
    add_n(q->t1h, q->t, one);
    for (int i = 0; i < 3; i++) q->t1h[i] = (q->t1h[i] >> 1) | (q->t1h[i + 1] << 63);
    q->t1h[3] >>= 1;
    memcpy(half, pm1, 32);
    for (int i = 0; i < 3; i++) half[i] = (half[i] >> 1) | (half[i + 1] << 63);
    half[3] >>= 1;
    for (uint64_t z = 2;; z++) {
        uint64_t raw[4] = {z, 0, 0, 0};
        fe zm, e;
        f_from_raw(f, &zm, raw);
        f_pow(f, &e, &zm, half);
        if (!f_eq(&e, &f->r)) { f_pow(f, &q->c, &zm, q->t); break; }
    }
}
/* returns 1 and a square root when x is a square (Tonelli-Shanks), 0 otherwise */
static int f_sqrt(const fctx *f, const sqrt_ctx *q, fe *r, const fe *x) {
    if (f_is_zero(x)) { *r = *x; return 1; }
    fe c = q->c, rr, tt;
    f_pow(f, &rr, x, q->t1h);
    f_pow(f, &tt, x, q->t);
    int m = q->s;
    while (!f_eq(&tt, &f->r)) {
        int i = 0;
        fe x2 = tt;
        while (!f_eq(&x2, &f->r)) { f_sqr(f, &x2, &x2); i++; if (i == m) return 0; }
        fe b = c;
        for (int k = 0; k < m - i - 1; k++) f_sqr(f, &b, &b);
        f_mul(f, &rr, &rr, &b);
        f_sqr(f, &c, &b);
        f_mul(f, &tt, &tt, &c);
        m = i;
    }
    *r = rr;
    return 1;
}
static int f_parity(const fctx *f, const fe *a) { uint64_t raw[4]; f_to_raw(f, raw, a); return (int)(raw[0] & 1); }
static void curve_rhs(const fctx *f, fe *r, const fe *x, const fe *a, const fe *b) {
    fe t;
    f_sqr(f, &t, x); f_add(f, &t, &t, a); f_mul(f, &t, &t, x); f_add(f, r, &t, b);
}
/* affine chord-and-tangent on y^2 = x^3 + a x + b; returns 0 for the identity */
static int aff_add(const fctx *f, fe *x3, fe *y3, const fe *x1, const fe *y1, const fe *x2, const fe *y2, const fe *a) {
    fe lam, num, den, t;
    if (f_eq(x1, x2)) {
        f_add(f, &t, y1, y2);
        if (f_is_zero(&t)) return 0;
        f_sqr(f, &num, x1); f_add(f, &t, &num, &num); f_add(f, &num, &t, &num); f_add(f, &num, &num, a);
        f_add(f, &den, y1, y1);
    } else { f_sub(f, &num, y2, y1); f_sub(f, &den, x2, x1); }
    f_inv(f, &den, &den);
    f_mul(f, &lam, &num, &den);
    f_sqr(f, &t, &lam); f_sub(f, &t, &t, x1); f_sub(f, x3, &t, x2);
    f_sub(f, &t, x1, x3); f_mul(f, &t, &lam, &t); f_sub(f, y3, &t, y1);
    return 1;
}
int oracle_hash_to_curve_batch(int curve_id, int method, const uint8_t *consts, const uint8_t *dst_prime, size_t dst_len, const uint8_t *msgs,
                               size_t msg_len, size_t n, uint8_t *out, int nthreads) {
    init_fields();
    if (curve_id < 0 || curve_id > 3 || dst_len > 120 || msg_len > 100) return -1;
    const fctx *f = &F[CURVE_BASE[curve_id]];
    const int nc = method == 0 ? 6 : 17;
    fe K[17];
    for (int i = 0; i < nc; i++) { uint64_t raw[4]; memcpy(raw, consts + 32 * i, 32); if (!raw_reduced(f, raw)) return -2; f_from_raw(f, &K[i], raw); }
    sqrt_ctx sq;
    sqrt_init(f, &sq);
    fe zero;
    memset(&zero, 0, sizeof zero);
    fe nb_over_a, b_over_za;
    if (method == 1) {
        fe ia, iz;
        f_inv(f, &ia, &K[2]); f_mul(f, &nb_over_a, &K[3], &ia); f_neg(f, &nb_over_a, &nb_over_a);
        f_inv(f, &iz, &K[1]); f_mul(f, &b_over_za, &K[3], &ia); f_mul(f, &b_over_za, &b_over_za, &iz);
    }
    int bad = 0;
#pragma omp parallel for schedule(dynamic, 16) num_threads(nthreads > 0 ? nthreads : 1) reduction(| : bad)
    for (long long idx = 0; idx < (long long)n; idx++) {
        const uint8_t *msg = msgs + (size_t)idx * msg_len;
        uint8_t buf[400], b0[64], b1[64], b2[64];
        size_t len = 0;
        memset(buf, 0, 128); len = 128;
        memcpy(buf + len, msg, msg_len); len += msg_len;
        buf[len++] = 0; buf[len++] = 128; buf[len++] = 0;
        memcpy(buf + len, dst_prime, dst_len); len += dst_len;
        b2b_512(buf, len, b0);
        memcpy(buf, b0, 64); buf[64] = 1; memcpy(buf + 65, dst_prime, dst_len);
        b2b_512(buf, 65 + dst_len, b1);
        for (int i = 0; i < 64; i++) buf[i] = b0[i] ^ b1[i];
        buf[64] = 2;
        b2b_512(buf, 65 + dst_len, b2);
        fe u[2], px[2], py[2];
        fe_from_be64(f, &u[0], b1);
        fe_from_be64(f, &u[1], b2);
        for (int k = 0; k < 2; k++) {
            fe x, y, gx, t;
            if (method == 0) {                       /* SVDW, RFC 9380 6.6.1 */
                const fe *b = &K[0], *Z = &K[1], *c1 = &K[2], *c2 = &K[3], *c3 = &K[4], *c4 = &K[5];
                fe tv1, tv2, tv3, tv4, x1, x2, x3;
                f_sqr(f, &tv1, &u[k]); f_mul(f, &tv1, &tv1, c1);
                f_add(f, &tv2, &f->r, &tv1); f_sub(f, &tv1, &f->r, &tv1);
                f_mul(f, &tv3, &tv1, &tv2); f_inv(f, &tv3, &tv3);
                f_mul(f, &tv4, &u[k], &tv1); f_mul(f, &tv4, &tv4, &tv3); f_mul(f, &tv4, &tv4, c3);
                f_sub(f, &x1, c2, &tv4); f_add(f, &x2, c2, &tv4);
                f_sqr(f, &x3, &tv2); f_mul(f, &x3, &x3, &tv3); f_sqr(f, &x3, &x3); f_mul(f, &x3, &x3, c4); f_add(f, &x3, &x3, Z);
                curve_rhs(f, &gx, &x1, &zero, b);
                if (f_sqrt(f, &sq, &y, &gx)) x = x1;
                else {
                    curve_rhs(f, &gx, &x2, &zero, b);
                    if (f_sqrt(f, &sq, &y, &gx)) x = x2;
                    else { curve_rhs(f, &gx, &x3, &zero, b); if (!f_sqrt(f, &sq, &y, &gx)) bad |= 1; x = x3; }
                }
            } else {                                 /* simplified SWU on the isogenous curve, RFC 9380 6.6.2 */
                const fe *Z = &K[1], *ia = &K[2], *ib = &K[3];
                fe zu2, ta, tv1, x1;
                f_sqr(f, &zu2, &u[k]); f_mul(f, &zu2, &zu2, Z);
                f_sqr(f, &ta, &zu2); f_add(f, &ta, &ta, &zu2);
                f_inv(f, &tv1, &ta);
                if (f_is_zero(&tv1)) x1 = b_over_za;
                else { f_add(f, &t, &f->r, &tv1); f_mul(f, &x1, &nb_over_a, &t); }
                curve_rhs(f, &gx, &x1, ia, ib);
                if (f_sqrt(f, &sq, &y, &gx)) x = x1;
                else { f_mul(f, &x, &zu2, &x1); curve_rhs(f, &gx, &x, ia, ib); if (!f_sqrt(f, &sq, &y, &gx)) bad |= 1; }
            }
            if (f_parity(f, &u[k]) != f_parity(f, &y)) f_neg(f, &y, &y);
            px[k] = x; py[k] = y;
        }
        fe rx, ry;
        int finite = aff_add(f, &rx, &ry, &px[0], &py[0], &px[1], &py[1], method == 0 ? &zero : &K[2]);
        if (finite && method == 1) {               /* the 3-isogeny onto the target curve */
            const fe *c = &K[4];
            fe nx, dx, ny, dy, t;
            f_mul(f, &nx, &c[0], &rx); f_add(f, &nx, &nx, &c[1]); f_mul(f, &nx, &nx, &rx); f_add(f, &nx, &nx, &c[2]); f_mul(f, &nx, &nx, &rx); f_add(f, &nx, &nx, &c[3]);
            f_add(f, &dx, &rx, &c[4]); f_mul(f, &dx, &dx, &rx); f_add(f, &dx, &dx, &c[5]);
            f_mul(f, &ny, &c[6], &rx); f_add(f, &ny, &ny, &c[7]); f_mul(f, &ny, &ny, &rx); f_add(f, &ny, &ny, &c[8]); f_mul(f, &ny, &ny, &rx); f_add(f, &ny, &ny, &c[9]);
            f_mul(f, &ny, &ny, &ry);
            f_add(f, &dy, &rx, &c[10]); f_mul(f, &dy, &dy, &rx); f_add(f, &dy, &dy, &c[11]); f_mul(f, &dy, &dy, &rx); f_add(f, &dy, &dy, &c[12]);
            if (f_is_zero(&dx) || f_is_zero(&dy)) finite = 0;
            else { f_inv(f, &t, &dx); f_mul(f, &rx, &nx, &t); f_inv(f, &t, &dy); f_mul(f, &ry, &ny, &t); }
        }
        uint64_t raw[4];
        if (!finite) memset(out + 64 * (size_t)idx, 0, 64);
        else { f_to_raw(f, raw, &rx); memcpy(out + 64 * (size_t)idx, raw, 32); f_to_raw(f, raw, &ry); memcpy(out + 64 * (size_t)idx + 32, raw, 32); }
    }
    return bad ? -3 : 0;
}

/* ------------------------------------------------------------------ N4: sum-check prover rounds (CPU port) */
/* Restates oracle/sumcheck.py (prove_quad / prove_cubic_with_additive_term of Arecibo's SumcheckProof) on Montgomery arrays with OpenMP:
 * the CPU baseline of the N4 sum-check rows.  kind 0: A B (2 polynomials), kind 1: A (B C - D) (4 polynomials). */
typedef struct { int field, kind, k; size_t len; fe *poly[4]; } sc_state;
void *oracle_sc_new(int field_id, int kind, const uint8_t *polys /* k arrays of len canonical elements, back to back */, size_t len) {
    init_fields();
    if (field_id < 0 || field_id > 3 || (kind != 0 && kind != 1)) return NULL;
    const fctx *f = &F[field_id];
    sc_state *st = calloc(1, sizeof *st);
    st->field = field_id; st->kind = kind; st->k = kind == 0 ? 2 : 4; st->len = len;
    for (int j = 0; j < st->k; j++) {
        st->poly[j] = malloc(sizeof(fe) * (len ? len : 1));
        for (size_t i = 0; i < len; i++) { uint64_t raw[4]; memcpy(raw, polys + 32 * ((size_t)j * len + i), 32); f_from_raw(f, &st->poly[j][i], raw); }
    }
    return st;
}
void oracle_sc_free(void *h) {
    sc_state *st = h;
    if (!st) return;
    for (int j = 0; j < st->k; j++) free(st->poly[j]);
    free(st);
}
/* s(0), s(2)[, s(3)] of the current round, canonical */
int oracle_sc_round(void *h, uint8_t *evals_out, int nthreads) {
    sc_state *st = h;
    const fctx *f = &F[st->field];
    const size_t half = st->len / 2;
    const int E = st->kind == 0 ? 2 : 3, K = st->k;
    if (nthreads < 1) nthreads = 1;
    fe *part = calloc((size_t)nthreads * 3, sizeof(fe));
#pragma omp parallel num_threads(nthreads)
    {
        int tid = 0;
#ifdef _OPENMP
        tid = omp_get_thread_num();
#endif
        fe acc[3];
        memset(acc, 0, sizeof acc);
#pragma omp for schedule(static)
        for (long long i = 0; i < (long long)half; i++) {
            fe lo[4], p2[4], p3[4], d, t;
            for (int j = 0; j < K; j++) {
                lo[j] = st->poly[j][i];
                f_sub(f, &d, &st->poly[j][half + i], &lo[j]);
                f_add(f, &p2[j], &st->poly[j][half + i], &d);
                f_add(f, &p3[j], &p2[j], &d);
            }
            if (st->kind == 0) {
                f_mul(f, &t, &lo[0], &lo[1]); f_add(f, &acc[0], &acc[0], &t);
                f_mul(f, &t, &p2[0], &p2[1]); f_add(f, &acc[1], &acc[1], &t);
            } else {
                f_mul(f, &t, &lo[1], &lo[2]); f_sub(f, &t, &t, &lo[3]); f_mul(f, &t, &t, &lo[0]); f_add(f, &acc[0], &acc[0], &t);
                f_mul(f, &t, &p2[1], &p2[2]); f_sub(f, &t, &t, &p2[3]); f_mul(f, &t, &t, &p2[0]); f_add(f, &acc[1], &acc[1], &t);
                f_mul(f, &t, &p3[1], &p3[2]); f_sub(f, &t, &t, &p3[3]); f_mul(f, &t, &t, &p3[0]); f_add(f, &acc[2], &acc[2], &t);
            }
        }
        for (int e = 0; e < 3; e++) part[(size_t)tid * 3 + e] = acc[e];
    }
    for (int e = 0; e < E; e++) {
        fe s;
        memset(&s, 0, sizeof s);
        for (int t = 0; t < nthreads; t++) f_add(f, &s, &s, &part[(size_t)t * 3 + e]);
        uint64_t raw[4];
        f_to_raw(f, raw, &s);
        memcpy(evals_out + 32 * e, raw, 32);
    }
    free(part);
    return 0;
}
/* bind_poly_var_top with the canonical challenge r; the length halves */
int oracle_sc_bind(void *h, const uint8_t r_bytes[32], int nthreads) {
    sc_state *st = h;
    const fctx *f = &F[st->field];
    uint64_t raw[4];
    memcpy(raw, r_bytes, 32);
    if (!raw_reduced(f, raw)) return -2;
    fe r;
    f_from_raw(f, &r, raw);
    const size_t half = st->len / 2;
    for (int j = 0; j < st->k; j++) {
        fe *P = st->poly[j];
#pragma omp parallel for schedule(static) num_threads(nthreads > 0 ? nthreads : 1)
        for (long long i = 0; i < (long long)half; i++) {
            fe d;
            f_sub(f, &d, &P[half + i], &P[i]);
            f_mul(f, &d, &d, &r);
            f_add(f, &P[i], &P[i], &d);
        }
    }
    st->len = half;
    return 0;
}
/* element 0 of every polynomial, canonical (the final evaluations once the length is 1) */
int oracle_sc_heads(void *h, uint8_t *out) {
    sc_state *st = h;
    const fctx *f = &F[st->field];
    for (int j = 0; j < st->k; j++) { uint64_t raw[4]; f_to_raw(f, raw, &st->poly[j][0]); memcpy(out + 32 * j, raw, 32); }
    return 0;
}

int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
