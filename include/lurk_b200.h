/*
 * lurk_b200.h -- C ABI of liblurk_b200.so: the B200 (sm_100a) implementation of lurk-beta's Nova/SuperNova
 * proving hot path.  Plain pointers and sizes only; no C++/torch types.
 *
 * The reference (argumentcomputer/lurk-beta @ f238d85c) has no FFI of its own: GPU work is delegated to
 * third-party crates via `--features cuda = ["neptune/cuda", "nova/cuda"]` (Cargo.toml:105-110).  Each entry
 * point below names the Rust seam it sits behind (reference file:line); INTEGRATION.md shows the
 * `extern "C"` binding a maintainer adds on the Rust side.
 *
 * Conventions
 *   - Field element: 32 bytes little-endian.  LURK_FMT_CANONICAL = the integer < p, i.e.
 *     ff::PrimeField::to_repr (src/field.rs:72-81).  LURK_FMT_MONTGOMERY = x * 2^256 mod p as 4 x u64, the
 *     in-memory form of pasta_curves (feature repr-c, Cargo.toml:42) and halo2curves field types, so Rust
 *     slices of `F` can be passed without conversion.
 *   - Affine point: x | y (64 bytes); the identity is (0, 0).  Result point: x | y | z (96 bytes) with
 *     z = 1 (finite) or x = y = z = 0 (identity), in the format asked for.
 *   - Ownership: the caller owns every buffer.  Contexts are created/destroyed by paired calls.
 *   - Errors: 0 = LURK_OK, negative = error; lurk_last_error() returns a thread-local message.  The library
 *     never aborts or unwinds (reference error style: Result<_, ProofError>, src/error.rs:8-18).
 *   - Threading: every call is re-entrant; `*_dev` calls are asynchronous on the given CUDA stream
 *     (a cudaStream_t passed as void*; NULL = default stream), host-buffer calls synchronise before returning.
 *   - Non-canonical inputs (>= p) are rejected with LURK_ERR_RANGE by the host-buffer calls (mirrors
 *     from_repr failing, src/field.rs:76-81); `*_dev` calls assume reduced inputs.
 *   - There is no CPU fallback: without a CUDA device every compute call returns LURK_ERR_NOGPU.
 */
#ifndef LURK_B200_H
#define LURK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* LanguageField (src/field.rs:40-50) */
#define LURK_FIELD_BN254_FR 0  /* LanguageField::BN256   = halo2curves::bn256::Fr  (default, all benches) */
#define LURK_FIELD_BN254_FQ 1  /* LanguageField::Grumpkin = grumpkin::Fr = bn256::Fq                       */
#define LURK_FIELD_PALLAS_FQ 2 /* LanguageField::Pallas  = pallas::Scalar                                  */
#define LURK_FIELD_PALLAS_FP 3 /* LanguageField::Vesta   = vesta::Scalar = pallas::Base                    */

/* curves of the Nova curve cycles (src/proof/nova.rs:57-71) */
#define LURK_CURVE_BN254_G1 0 /* base Fq(1), scalars Fr(0) */
#define LURK_CURVE_GRUMPKIN 1 /* base Fr(0), scalars Fq(1) */
#define LURK_CURVE_PALLAS 2   /* base Fp(3), scalars Fq(2) */
#define LURK_CURVE_VESTA 3    /* base Fq(2), scalars Fp(3) */

#define LURK_FMT_CANONICAL 0
#define LURK_FMT_MONTGOMERY 1

#define LURK_OK 0
#define LURK_ERR_ARG (-1)
#define LURK_ERR_CUDA (-2)
#define LURK_ERR_OOM (-3)
#define LURK_ERR_RANGE (-4)
#define LURK_ERR_NOGPU (-5)
#define LURK_ERR_ORDER (-6) /* DAG nodes not topologically ordered */

const char *lurk_last_error(void);
int lurk_version(void);
int lurk_device_count(void);
/* modulus of a field as 32 bytes LE */
int lurk_field_modulus(int field_id, uint8_t out[32]);

/* ---------------------------------------------------------------------------------------------------
 * S1  Poseidon digests.  Replaces PoseidonCache::hash3/hash4/hash6/hash8 = neptune
 *     Poseidon::new_with_preimage(..).hash() (src/hash.rs:180-203) for a batch of independent preimages.
 *     arity in {3,4,6,8}; preimages n*arity elements, digests n elements.
 * ------------------------------------------------------------------------------------------------- */
int lurk_poseidon_hash_batch(int field_id, int arity, const uint8_t *preimages, size_t n, uint8_t *digests);
int lurk_poseidon_hash_batch_mont(int field_id, int arity, const uint8_t *preimages, size_t n, uint8_t *digests);
int lurk_poseidon_hash_batch_dev(int field_id, int arity, const void *d_preimages, size_t n, void *d_digests,
                                 int fmt, void *stream);
/* Constants as PoseidonConstants::new() builds them (src/hash.rs:61-72): R_F, R_P, the t*(R_F+R_P) round
 * constants and the t*t MDS matrix (row-major), canonical form.  Buffers may be NULL to query sizes. */
int lurk_poseidon_constants(int field_id, int arity, int *full_rounds, int *partial_rounds,
                            uint8_t *round_constants, uint8_t *mds);

/* ---------------------------------------------------------------------------------------------------
 * S3  Slot witnesses.  Replaces the per-slot body of generate_slots_witnesses (src/lem/multiframe.rs:520-592):
 *     allocate_slot -> neptune circuit2 poseidon_hash_allocated in witness mode (src/lem/circuit.rs:212-315).
 *     Output block per slot = [preimage (arity) | 3 aux per S-box in Neptune's optimised-round order | digest],
 *     lurk_poseidon_witness_block() elements (= hashN_cost + N, src/lem/multiframe.rs:503-516).
 *     Bit-decomposition slots: [value | aux of AllocatedNum::to_bits_le_strict] (src/lem/circuit.rs:241-243),
 *     lurk_bitdecomp_witness_block() elements (= BIT_DECOMP_*_WITNESS_SIZE, src/lem/multiframe.rs:495-498).
 * ------------------------------------------------------------------------------------------------- */
size_t lurk_poseidon_witness_block(int field_id, int arity);
int lurk_poseidon_witness_batch(int field_id, int arity, const uint8_t *preimages, size_t n, uint8_t *blocks,
                                int fmt);
int lurk_poseidon_witness_batch_dev(int field_id, int arity, const void *d_preimages, size_t n, void *d_blocks,
                                    int fmt, void *stream);
/* In-place form for the step witness: block k is written at element offset d_offsets[k] (u64, device) of d_base.  In the
 * reference every frame's aux is [its slot blocks in slot order | LEM body aux] (synthesize_frames_parallel,
 * src/lem/multiframe.rs:635-712), so the blocks of one slot type are strided by the frame length (9119 on BN256). */
int lurk_poseidon_witness_scatter_dev(int field_id, int arity, const void *d_preimages, size_t n, void *d_base,
                                      const void *d_offsets, int fmt, void *stream);
int lurk_bitdecomp_witness_scatter_dev(int field_id, const void *d_values, size_t n, void *d_base, const void *d_offsets,
                                       int fmt, void *stream);
size_t lurk_bitdecomp_witness_block(int field_id);
int lurk_bitdecomp_witness_batch(int field_id, const uint8_t *values, size_t n, uint8_t *blocks, int fmt);
int lurk_bitdecomp_witness_batch_dev(int field_id, const void *d_values, size_t n, void *d_blocks, int fmt,
                                     void *stream);

/* ---------------------------------------------------------------------------------------------------
 * S2  DAG hydration.  Replaces StoreCore::hydrate_z_cache / hash_ptr_val_unsafe (src/lem/store_core.rs:199-269)
 *     with the preimage layouts of `impl StoreHasher for PoseidonCache` (src/lem/store.rs:29-78).
 *     child[i] < n_atoms refers to atom digest child[i]; otherwise to node (child[i] - n_atoms), which must
 *     precede the referring node (children first).  out_digests: n elements, canonical.
 * ------------------------------------------------------------------------------------------------- */
#define LURK_DAG_TUPLE2 2     /* H4 [t0, d0, t1, d1]                 hash_ptrs len 2 (store.rs:31-36)  */
#define LURK_DAG_TUPLE3 3     /* H6                                   hash_ptrs len 3 (store.rs:37-50)  */
#define LURK_DAG_TUPLE4 4     /* H8                                   hash_ptrs len 4 (store.rs:51-67)  */
#define LURK_DAG_COMPACT 5    /* H4 [d0, t1, d1, d2]                  hash_compact    (store.rs:75-77)  */
#define LURK_DAG_COMMITMENT 6 /* H3 [d0 = secret, t1, d1]            hash_commitment (store.rs:70-73)  */
typedef struct lurk_dag_node {
    uint8_t kind;
    uint8_t reserved;
    uint16_t tag[4];   /* Tag::to_field = F::from(u16) (src/tag.rs:99-101) */
    uint32_t child[4];
} lurk_dag_node;
int lurk_dag_hash(int field_id, const lurk_dag_node *nodes, size_t n, const uint8_t *atom_digests,
                  size_t n_atoms, uint8_t *out_digests);

/* ---------------------------------------------------------------------------------------------------
 * S4  Pedersen commitment = multi-scalar multiplication.  Replaces Arecibo
 *     CommitmentEngineTrait::commit -> DlogGroup::vartime_multiscalar_mul(scalars, bases) (called from
 *     RecursiveSNARK::prove_step, src/proof/nova.rs:287,292; supernova.rs:231-244).  The fixed commitment key
 *     is uploaded once into a context; each call streams scalars.
 * ------------------------------------------------------------------------------------------------- */
typedef struct lurk_msm_ctx lurk_msm_ctx;
/* bases_affine: n * 64 bytes, x || y per point, (0, 0) = identity.  Coordinates >= p or points off the curve are
 * rejected with LURK_ERR_RANGE (what the reference's point deserialisation checks before a key is used).  A context
 * belongs to the device that is current at creation; using it with another device current is LURK_ERR_ARG. */
int lurk_msm_ctx_create(int curve_id, const uint8_t *bases_affine, size_t n, int fmt, lurk_msm_ctx **out);
/* bases already on the current device (n * 64 bytes, Montgomery); the context borrows the pointer */
int lurk_msm_ctx_create_dev(int curve_id, const void *d_bases_mont, size_t n, lurk_msm_ctx **out);
void lurk_msm_ctx_destroy(lurk_msm_ctx *ctx);
/* sum_{i<n} scalars[i] * bases[i], n <= size of the key */
int lurk_msm_ctx_run(lurk_msm_ctx *ctx, const uint8_t *scalars, size_t n, int fmt, uint8_t out_xyz[96]);
int lurk_msm_ctx_run_dev(lurk_msm_ctx *ctx, const void *d_scalars, size_t n, int fmt, uint8_t out_xyz[96],
                         void *stream);
/* Fixed-base acceleration (the key never changes between folds): builds table[w][i] = 2^(c w) * bases[i] once
 * (nwin x n x 64 bytes of HBM; 1.7 GB for a 2^21-point key) so that all windows share one bucket set and a wider
 * window (c = 20) becomes affordable: ~13 instead of 16 bucket additions per scalar.  Results are unchanged.
 * Call before cloning; clones share the table. */
int lurk_msm_ctx_precompute(lurk_msm_ctx *ctx);
/* Asynchronous form: `launch` enqueues the whole commitment on `stream` and returns; `finish` waits for it and
 * produces the point.  One launch may be pending per context; `clone` gives another context on the same resident key
 * (own scratch; the parent must outlive it) so that e.g. commit(W) and commit(T) of one fold overlap. */
int lurk_msm_ctx_launch_dev(lurk_msm_ctx *ctx, const void *d_scalars, size_t n, int fmt, void *stream);
int lurk_msm_ctx_finish(lurk_msm_ctx *ctx, uint8_t out_xyz[96]);
int lurk_msm_ctx_clone(lurk_msm_ctx *ctx, lurk_msm_ctx **out);
/* Measurement hooks: when enabled, every run records CUDA events around the bucket-accumulation kernel (the dominant
 * kernel) on the launching stream; last_profile returns its duration and the number of kernels the run launched. */
int lurk_msm_ctx_set_profiling(lurk_msm_ctx *ctx, int enable);
int lurk_msm_ctx_last_profile(lurk_msm_ctx *ctx, float *accumulate_ms, unsigned *kernel_launches);
/* one-shot convenience (uploads bases every call) */
int lurk_msm(int curve_id, const uint8_t *bases_affine, const uint8_t *scalars, size_t n, int fmt,
             uint8_t out_xyz[96]);
/* Synthetic commitment key: bases_out[i] = [start + i + 1] G for the curve's standard generator, affine, n * 64 bytes
 * (host, multi-threaded).  The reference derives its key by hash-to-curve / powers of tau inside Arecibo
 * (public_params, src/proof/nova.rs:196-216) -- out of scope; this gives benches and tests a deterministic key of
 * distinct points. */
int lurk_synthetic_bases(int curve_id, uint64_t start, size_t n, int fmt, uint8_t *bases_out);
/* host-side sum of `count` result points (the per-GPU partial sums of a sharded commitment key) */
int lurk_point_sum(int curve_id, const uint8_t *points_xyz, size_t count, int fmt, uint8_t out_xyz[96]);

/* ---------------------------------------------------------------------------------------------------
 * S5  Fold helpers on device-resident vectors (Arecibo NIFS::prove / R1CSShape::commit_T /
 *     RelaxedR1CSWitness::fold).  All vectors Montgomery form on the device.
 * ------------------------------------------------------------------------------------------------- */
/* out[i] = a[i] + r * b[i]  (W <- W1 + r W2, E <- E1 + r T).  r: 32 bytes host, Montgomery. out may alias a. */
int lurk_axpy_dev(int field_id, const void *d_a, const void *d_b, const uint8_t r_mont[32], size_t n, void *d_out,
                  void *stream);
/* y = M z for a CSR matrix (row_ptr: rows+1 x u64, col: nnz x u32, val: nnz elements) */
int lurk_spmv_csr_dev(int field_id, const void *d_row_ptr, const void *d_col, const void *d_val, size_t rows,
                      const void *d_z, void *d_y, void *stream);
/* T = az1*bz2 + az2*bz1 - u1*cz2 - u2*cz1 */
int lurk_cross_term_dev(int field_id, const void *d_az1, const void *d_bz1, const void *d_cz1, const void *d_az2,
                        const void *d_bz2, const void *d_cz2, const uint8_t u1_mont[32], const uint8_t u2_mont[32],
                        size_t n, void *d_t, void *stream);
/* element-wise format conversion on the device (LURK_FMT_*), in place allowed */
int lurk_convert_dev(int field_id, const void *d_in, size_t n, int to_fmt, void *d_out, void *stream);

/* ---------------------------------------------------------------------------------------------------
 * K6  Number-theoretic transform (north_star; no call site in the reference -- SURVEY.md D4).
 *     In-place length-2^log_n DFT over the field's 2-adic subgroup, natural order in and out, Montgomery form.
 *     Roots: omega = g^((p-1)/2^s) with g the multiplicative generator of halo2curves / pasta_curves.
 * ------------------------------------------------------------------------------------------------- */
int lurk_ntt_dev(int field_id, void *d_data, int log_n, int inverse, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* LURK_B200_H */
