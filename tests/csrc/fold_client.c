/* A plain C99 driver of the fold context (lurk_fold_ctx_*, include/lurk_b200.h): what the Rust side of
 * `Proof::prove_recursively` (src/proof/nova.rs:260-339) would do through bindgen, with no Python, torch or oracle in the
 * process.  It builds a small step circuit that is satisfiable by construction, folds four fresh instances (stage A one step
 * ahead of stage B) and checks protocol-level facts it can verify by itself:
 *   - the device-side verifier check: relaxed R1CS residual 0, commit(W) / commit(E) equal the folded commitments;
 *   - u of the running instance = 1 + r_1 + r_2 + r_3 (the challenges returned by the steps), X = X_0 + sum r_i X_i;
 *   - a checkpoint (get_running) installed into a second context continues to the same running commitments.
 * Without a GPU every compute entry point must fail loudly with LURK_ERR_NOGPU. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "lurk_b200.h"

#define M 48      /* free ("slot") columns */
#define K 20      /* defined columns g_j = s_a(j) * s_b(j) */
#define NW (M + K)
#define ROWS (3 * K)
#define NX 2

static int fail(int code, const char *what) {
    fprintf(stderr, "fold_client: %s (last error: %s)\n", what, lurk_last_error());
    return code;
}
static void put_u64(uint8_t *dst, uint64_t v) { memset(dst, 0, 32); for (int i = 0; i < 8; i++) dst[i] = (uint8_t)(v >> (8 * i)); }
static uint32_t rng_state = 12345;
static uint32_t rnd(void) { rng_state = rng_state * 1664525u + 1013904223u; return rng_state >> 8; }

/* little-endian 256-bit add / multiply-by-small helpers (values stay far below the 254-bit modulus) */
static void add256(uint8_t *acc, const uint8_t *x) { unsigned c = 0; for (int i = 0; i < 32; i++) { unsigned s = acc[i] + x[i] + c; acc[i] = (uint8_t)s; c = s >> 8; } }
static void mul_small(uint8_t *out, const uint8_t *x, uint64_t k) {    /* out = x * k, x < 2^128, k < 2^62 */
    unsigned __int128 carry = 0;
    for (int i = 0; i < 32; i++) { unsigned __int128 t = (unsigned __int128)x[i] * k + carry; out[i] = (uint8_t)t; carry = t >> 8; }
}

int main(void) {
    uint64_t rp[3][ROWS + 1];
    uint32_t col[3][2 * ROWS];
    uint8_t val[3][2 * ROWS * 32];
    int a_of[K], b_of[K];
    size_t nnz[3] = {0, 0, 0};
    for (int j = 0; j < K; j++) { a_of[j] = (int)(rnd() % M); b_of[j] = (int)(rnd() % M); }
    for (int r = 0; r < ROWS; r++) {
        int j = r % K, kind = r / K;       /* 0: definition, 1: the same with other coefficients, 2: linear row */
        for (int m = 0; m < 3; m++) rp[m][r] = nnz[m];
        if (kind < 2) {
            uint64_t l = kind ? 2 : 1, mu = kind ? 3 : 1;
            col[0][nnz[0]] = (uint32_t)a_of[j]; put_u64(val[0] + 32 * nnz[0]++, l);
            col[1][nnz[1]] = (uint32_t)b_of[j]; put_u64(val[1] + 32 * nnz[1]++, mu);
            col[2][nnz[2]] = (uint32_t)(M + j); put_u64(val[2] + 32 * nnz[2]++, l * mu);
        } else {                           /* (s_a + x_0) * u = (s_a + x_0) */
            col[0][nnz[0]] = (uint32_t)a_of[j]; put_u64(val[0] + 32 * nnz[0]++, 1);
            col[0][nnz[0]] = NW + 1; put_u64(val[0] + 32 * nnz[0]++, 1);
            col[1][nnz[1]] = NW; put_u64(val[1] + 32 * nnz[1]++, 1);
            col[2][nnz[2]] = (uint32_t)a_of[j]; put_u64(val[2] + 32 * nnz[2]++, 1);
            col[2][nnz[2]] = NW + 1; put_u64(val[2] + 32 * nnz[2]++, 1);
        }
    }
    for (int m = 0; m < 3; m++) rp[m][ROWS] = nnz[m];

    uint8_t *bases = malloc(64 * 256);
    if (lurk_synthetic_bases(LURK_CURVE_BN254_G1, 0, 256, LURK_FMT_CANONICAL, bases) != LURK_OK) return fail(1, "synthetic bases");
    lurk_msm_ctx *ck = NULL;
    int rc = lurk_msm_ctx_create(LURK_CURVE_BN254_G1, bases, 256, LURK_FMT_CANONICAL, &ck);
    if (lurk_device_count() <= 0) {
        if (rc != LURK_ERR_NOGPU || ck != NULL) return fail(2, "context creation without a GPU must fail loudly");
        lurk_fold_ctx *none = NULL;
        lurk_fold_config bad;
        memset(&bad, 0, sizeof bad);
        if (lurk_fold_ctx_create(&bad, NULL, NULL, &none) != LURK_ERR_ARG || none != NULL) return fail(3, "null arguments must be rejected");
        puts("fold_client ok (no GPU: compute entry points fail loudly)");
        return 0;
    }
    if (rc != LURK_OK) return fail(4, "msm ctx");

    lurk_fold_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.curve_id = LURK_CURVE_BN254_G1; cfg.depth = 2; cfg.n_w = NW; cfg.n_x = NX; cfg.n_rows = ROWS;
    for (int m = 0; m < 3; m++) { cfg.row_ptr[m] = rp[m]; cfg.col[m] = col[m]; cfg.val[m] = val[m]; }
    cfg.fmt = LURK_FMT_CANONICAL; cfg.world = 1; cfg.rank = 0;
    lurk_fold_ctx *ctx = NULL, *ctx2 = NULL;
    if (lurk_fold_ctx_create(&cfg, ck, ck, &ctx) != LURK_OK) return fail(5, "fold ctx");
    cfg.depth = 1;
    if (lurk_fold_ctx_create(&cfg, ck, ck, &ctx2) != LURK_OK) return fail(5, "fold ctx 2");
    lurk_fold_span span = {0, NW, NW, 1};
    if (lurk_fold_ctx_set_spans(ctx, 1, &span) != LURK_OK || lurk_fold_ctx_set_spans(ctx2, 1, &span) != LURK_OK) return fail(6, "spans");

    uint8_t u_want[32], x_want[NX][32], tmp[32];
    memset(u_want, 0, 32); memset(x_want, 0, sizeof x_want);
    u_want[0] = 1;
    uint64_t xs[4][NX];
    uint8_t snap_W[NW * 32], snap_E[ROWS * 32], snap_u[32], snap_X[NX * 32], snap_cw[96], snap_ce[96];
    lurk_fold_result res, res2;
    for (int step = 0; step < 4; step++) {
        int b = step & 1;
        /* the CPU witness generator: fill the pinned buffers of fresh-instance buffer b */
        void *w, *x, *ro;
        size_t bytes;
        lurk_fold_ctx *c = ctx;
        if (lurk_fold_ctx_host_buffer(c, b, LURK_FOLD_BUF_GLUE, &w, &bytes) != LURK_OK || bytes != NW * 32) return fail(7, "glue buffer");
        if (lurk_fold_ctx_host_buffer(c, b, LURK_FOLD_BUF_X2, &x, &bytes) != LURK_OK || bytes != NX * 32) return fail(7, "x2 buffer");
        if (lurk_fold_ctx_host_buffer(c, b, LURK_FOLD_BUF_RO, &ro, &bytes) != LURK_OK || bytes != 24 * 32) return fail(7, "ro buffer");
        uint64_t s[M];
        for (int i = 0; i < M; i++) { s[i] = rnd() & 0xffff; put_u64((uint8_t *)w + 32 * i, s[i]); }
        for (int j = 0; j < K; j++) put_u64((uint8_t *)w + 32 * (M + j), s[a_of[j]] * s[b_of[j]]);
        memset(ro, 0, 24 * 32);
        put_u64((uint8_t *)ro, 0xabcdef);                       /* pp digest */
        for (int k = 0; k < NX; k++) {
            xs[step][k] = rnd();
            put_u64((uint8_t *)x + 32 * k, xs[step][k]);
            put_u64((uint8_t *)ro + 32 * (4 + k), xs[step][k]);  /* U2.X absorbed at positions 4, 5 */
        }
        if (lurk_fold_ctx_stage_a(ctx, b, 0, LURK_FMT_CANONICAL) != LURK_OK) return fail(8, "stage A");
        if (step == 0) {
            if (lurk_fold_ctx_init_running(ctx, b) != LURK_OK) return fail(9, "init running");
        } else if (lurk_fold_ctx_stage_b_launch(ctx, b) != LURK_OK) return fail(10, "stage B");
        if (lurk_fold_ctx_collect(ctx, b, &res, LURK_FMT_CANONICAL) != LURK_OK || res.status != 0) return fail(11, "collect");
        if (step == 0) {
            for (int k = 0; k < NX; k++) put_u64(x_want[k], xs[0][k]);
        } else {
            for (int i = 16; i < 32; i++) if (res.r[i]) return fail(12, "challenge wider than 128 bits");
            add256(u_want, res.r);
            for (int k = 0; k < NX; k++) { mul_small(tmp, res.r, xs[step][k]); add256(x_want[k], tmp); }
        }
        if (step == 2 && lurk_fold_ctx_get_running(ctx, snap_W, snap_E, snap_u, snap_X, snap_cw, snap_ce, LURK_FMT_CANONICAL) != LURK_OK)
            return fail(13, "get running");
        if (step == 3) {
            /* the same fresh instance folded onto the checkpoint in a second context */
            void *w2, *x2, *ro2;
            lurk_fold_ctx_host_buffer(ctx2, 0, LURK_FOLD_BUF_GLUE, &w2, &bytes); memcpy(w2, w, NW * 32);
            lurk_fold_ctx_host_buffer(ctx2, 0, LURK_FOLD_BUF_X2, &x2, &bytes); memcpy(x2, x, NX * 32);
            lurk_fold_ctx_host_buffer(ctx2, 0, LURK_FOLD_BUF_RO, &ro2, &bytes); memcpy(ro2, ro, 24 * 32);
            if (lurk_fold_ctx_set_running(ctx2, snap_W, snap_E, snap_u, snap_X, snap_cw, snap_ce, LURK_FMT_CANONICAL) != LURK_OK) return fail(14, "set running");
            if (lurk_fold_ctx_stage_a(ctx2, 0, 0, LURK_FMT_CANONICAL) != LURK_OK || lurk_fold_ctx_stage_b_launch(ctx2, 0) != LURK_OK ||
                lurk_fold_ctx_collect(ctx2, 0, &res2, LURK_FMT_CANONICAL) != LURK_OK)
                return fail(15, "resumed step");
            if (memcmp(res.r, res2.r, 32) || memcmp(res.comm_T, res2.comm_T, 96) || memcmp(res.running_comm_W, res2.running_comm_W, 96) ||
                memcmp(res.running_comm_E, res2.running_comm_E, 96))
                return fail(16, "resumed context diverged");
        }
    }
    uint8_t u[32], X[NX * 32];
    if (lurk_fold_ctx_get_running(ctx, NULL, NULL, u, X, NULL, NULL, LURK_FMT_CANONICAL) != LURK_OK) return fail(17, "get running");
    if (memcmp(u, u_want, 32)) return fail(18, "u != 1 + sum of the challenges");
    for (int k = 0; k < NX; k++) if (memcmp(X + 32 * k, x_want[k], 32)) return fail(19, "X != X_0 + sum r_i X_i");
    uint64_t bad = 1;
    int okw = 0, oke = 0;
    if (lurk_fold_ctx_check_running(ctx, &bad, &okw, &oke) != LURK_OK || bad != 0 || !okw || !oke) return fail(20, "device-side relaxed R1CS check");
    if (lurk_fold_ctx_check_running(ctx2, &bad, &okw, &oke) != LURK_OK || bad != 0 || !okw || !oke) return fail(21, "resumed context check");
    /* misuse is an error code, never a crash */
    if (lurk_fold_ctx_stage_a(ctx, 7, 0, LURK_FMT_CANONICAL) != LURK_ERR_ARG) return fail(22, "bad buffer index");
    if (lurk_fold_ctx_collect(ctx, 0, &res, LURK_FMT_CANONICAL) != LURK_ERR_ARG) return fail(23, "nothing to collect");
    lurk_fold_ctx_destroy(ctx2);
    lurk_fold_ctx_destroy(ctx);
    lurk_msm_ctx_destroy(ck);
    free(bases);
    puts("fold_client ok");
    return 0;
}
