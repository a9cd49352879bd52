/* A plain C99 client of liblurk_b200.so: what a cgo / bindgen / JNI shim sees.  Built with gcc (not g++) against
 * include/lurk_b200.h by tests/test_host_library.py and run without a GPU: host-only entry points must work, compute
 * entry points must fail with LURK_ERR_NOGPU and a message (no CPU fallback).  With a GPU it checks golden G1 instead. */
#include <stdio.h>
#include <string.h>

#include "lurk_b200.h"

static int fail(int code, const char *what) {
    fprintf(stderr, "c_abi_client: %s (last error: %s)\n", what, lurk_last_error());
    return code;
}

int main(void) {
    if (lurk_version() <= 0) return fail(1, "version");
    /* witness sizes pinned by the reference: src/lem/multiframe.rs:991-1016, :495-498 */
    if (lurk_poseidon_witness_block(LURK_FIELD_BN254_FR, 4) != 293 || lurk_poseidon_witness_block(LURK_FIELD_BN254_FR, 8) != 396 ||
        lurk_poseidon_witness_block(LURK_FIELD_BN254_FR, 3) != 268 || lurk_poseidon_witness_block(LURK_FIELD_BN254_FR, 6) != 343)
        return fail(2, "slot witness sizes");
    if (lurk_bitdecomp_witness_block(LURK_FIELD_BN254_FR) != 354 || lurk_bitdecomp_witness_block(LURK_FIELD_PALLAS_FQ) != 298)
        return fail(3, "bit decomposition sizes");
    if (lurk_poseidon_witness_block(LURK_FIELD_BN254_FR, 5) != 0) return fail(4, "unsupported arity must give 0");

    /* host-side group arithmetic: [1]G + [2]G = [3]G on every curve */
    for (int curve = 0; curve < 4; curve++) {
        uint8_t bases[3 * 64], pts[2 * 96], sum[96];
        if (lurk_synthetic_bases(curve, 0, 3, LURK_FMT_CANONICAL, bases) != LURK_OK) return fail(5, "synthetic bases");
        memset(pts, 0, sizeof pts);
        for (int k = 0; k < 2; k++) {
            memcpy(pts + 96 * k, bases + 64 * k, 64);
            pts[96 * k + 64] = 1; /* z = 1 */
        }
        if (lurk_point_sum(curve, pts, 2, LURK_FMT_CANONICAL, sum) != LURK_OK) return fail(6, "point sum");
        if (memcmp(sum, bases + 128, 64) != 0 || sum[64] != 1) return fail(7, "[1]G + [2]G != [3]G");
    }
    if (lurk_point_sum(7, NULL, 0, LURK_FMT_CANONICAL, NULL) != LURK_ERR_ARG) return fail(8, "bad arguments must be rejected");

    /* N3, host-only pieces: the key length public_params asks for (fib rc = 100: 2^21) and the XOF behind from_label */
    if (lurk_ck_size(1114100, 911900, 0) != ((size_t)1 << 21) || lurk_ck_size(3, 5, 100) != 128) return fail(12, "lurk_ck_size");
    {
        static const uint8_t shake_empty[8] = {0x46, 0xb9, 0xdd, 0x2b, 0x0b, 0xa8, 0x8d, 0x13}; /* SHAKE256(""), FIPS 202 */
        uint8_t xof[8];
        if (lurk_shake256(NULL, 0, xof, sizeof xof) != LURK_OK || memcmp(xof, shake_empty, 8) != 0) return fail(13, "SHAKE256 known answer");
    }

    uint8_t pre[8 * 32], digest[32];
    memset(pre, 0, sizeof pre);
    int rc = lurk_poseidon_hash_batch(LURK_FIELD_BN254_FR, 8, pre, 1, digest);
    if (lurk_device_count() <= 0) {
        if (rc != LURK_ERR_NOGPU || strlen(lurk_last_error()) == 0) return fail(9, "compute without a GPU must fail loudly");
        lurk_msm_ctx *ctx = NULL;
        uint8_t g[64];
        lurk_synthetic_bases(LURK_CURVE_BN254_G1, 0, 1, LURK_FMT_CANONICAL, g);
        if (lurk_msm_ctx_create(LURK_CURVE_BN254_G1, g, 1, LURK_FMT_CANONICAL, &ctx) != LURK_ERR_NOGPU || ctx != NULL)
            return fail(10, "context creation without a GPU must fail loudly");
        uint8_t key[128];
        if (lurk_ck_generate(LURK_CURVE_BN254_G1, (const uint8_t *)"ck", 2, 2, LURK_FMT_CANONICAL, key) != LURK_ERR_NOGPU)
            return fail(14, "key generation without a GPU must fail loudly");
    } else {
        /* golden G1 = H8(0^8), src/coprocessor/trie/mod.rs:932 (big-endian hex 1ca5b207...f35b) */
        static const uint8_t g1_le_tail[4] = {0x07, 0xb2, 0xa5, 0x1c};
        if (rc != LURK_OK || digest[0] != 0x5b || memcmp(digest + 28, g1_le_tail, 4) != 0) return fail(11, "golden G1 on the GPU");
        /* N3: first point of from_label(b"ck") on BN254 G1 (tests/golden/ck_from_label.json: x = 2e9face5...d9a5debd) */
        uint8_t key[128];
        if (lurk_ck_generate(LURK_CURVE_BN254_G1, (const uint8_t *)"ck", 2, 2, LURK_FMT_CANONICAL, key) != LURK_OK || key[0] != 0xbd || key[31] != 0x2e)
            return fail(15, "from_label on the GPU");
    }
    puts("c_abi_client ok");
    return 0;
}
