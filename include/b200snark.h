/*
 * b200snark.h -- C ABI of libb200snark.so, the B200 (sm_100a) backend for the Groth16 prover hot path.
 *
 * The reference (arkworks-rs/snark) has NO FFI: its seam is trait-level.  Each entry point below
 * names the reference interface it sits behind (paths relative to /root/reference):
 *
 *   SNARK::prove / CircuitSpecificSetupSNARK::setup      snark/src/lib.rs:43-54, 84-93
 *   ConstraintSystem::to_matrices()                      relations/src/gr1cs/constraint_system.rs:768-774
 *   instance_assignment() / witness_assignment()         relations/src/gr1cs/constraint_system.rs:193-206
 *   Matrix<F>, mat_vec_mul                               relations/src/utils/matrix.rs:4,26-36
 *   Sr1csAdapter::evaluate_constraint                    relations/src/sr1cs/mod.rs:24-56
 *   SynthesisError (status codes)                        relations/src/utils/error.rs:5-21
 *   (out of tree)  ark-poly Radix2EvaluationDomain::{fft,ifft}, get_coset;
 *                                    ark-ec VariableBaseMSM::msm; ark-groth16 prover / generator
 *
 * DATA CONVENTIONS (what a Rust caller already has in memory):
 *   - Field element: little-endian limbs, MONTGOMERY form with R = 2^(64*N64), N64 = 4 for both
 *     scalar fields and BN254 Fq, 6 for BLS12-381 Fq.  Identical to ark-ff `Fp<MontBackend>`'s in-memory
 *     `BigInt<N>` -- `&[F]` can be passed as is.  "canonical" scalars (ark `into_bigint()`) are accepted
 *     where a `scalars_mont` flag says so.
 *   - G1 affine point: x || y (2 field elements, packed, no padding).  G2 affine: x.c0 || x.c1 || y.c0 ||
 *     y.c1.  The point at infinity is ALL-ZERO bytes ((0,0) is not on either curve).  ark-ec's `Affine`
 *     carries a separate `infinity: bool`; the adapter writes zeros for such points.
 *   - Matrices: CSR per matrix (row_ptr[n_rows+1] u64, col[nnz] u32, coeff[nnz] field elements) built
 *     from `to_matrices()` rows in order; duplicate / unsorted columns are allowed and are summed
 *     (constraint_system.rs:792-804 only filters zeros).  Column 0 is the constant One, columns
 *     1..n_inst-1 the instance variables, n_inst.. the witnesses (relations/src/utils/variable.rs:105-113).
 *   - z = instance_assignment || witness_assignment, z[0] == 1 (relations/src/sr1cs/mod.rs:199-200).
 *
 * OWNERSHIP: caller-owned buffers are only read/written during the call.  `mem` says where a buffer
 * lives: B2S_MEM_HOST (any host pointer; copied through the ctx's pinned staging) or B2S_MEM_DEVICE
 * (a device pointer on the ctx's GPU, e.g. torch tensor storage).  Handles are freed by b2s_*_free.
 *
 * THREADING: one in-flight call per ctx (calls on one ctx are serialised by an internal mutex).
 * ERRORS: never unwinds (reference builds with panic=abort for FFI, Cargo.toml:33); int32 status.
 * NO CPU FALLBACK: every entry point fails with B2S_ERR_NO_DEVICE when no sm_100 GPU is usable.
 */
#ifndef B200SNARK_H
#define B200SNARK_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2S_CURVE_BLS12_381 0
#define B2S_CURVE_BN254 1

#define B2S_MEM_HOST 0
#define B2S_MEM_DEVICE 1

/* status codes; 1..7 mirror SynthesisError (relations/src/utils/error.rs:5-21) */
#define B2S_OK 0
#define B2S_ERR_MISSING_CS 1
#define B2S_ERR_ASSIGNMENT_MISSING 2        /* z / scalar length does not match the uploaded matrices / key */
#define B2S_ERR_DIVISION_BY_ZERO 3
#define B2S_ERR_UNSATISFIABLE 4
#define B2S_ERR_POLYNOMIAL_DEGREE_TOO_LARGE 5 /* domain larger than 2^two_adicity or than the backend limit */
#define B2S_ERR_UNEXPECTED_IDENTITY 6
#define B2S_ERR_MALFORMED_VK 7
#define B2S_ERR_INVALID_ARG 16
#define B2S_ERR_NO_DEVICE 17
#define B2S_ERR_CUDA 18
#define B2S_ERR_OOM 19
#define B2S_ERR_NCCL 20

typedef struct b2s_ctx b2s_ctx;
typedef struct b2s_r1cs b2s_r1cs;   /* device-resident A/B/C in CSR (witness independent; upload once per circuit) */
typedef struct b2s_pk b2s_pk;       /* device-resident Groth16 proving key (or one base-range shard of it) */

/* ---- context ------------------------------------------------------------------------------ */
int32_t b2s_ctx_create(int32_t curve_id, int32_t device_ordinal, b2s_ctx** out);
void b2s_ctx_destroy(b2s_ctx* ctx);
const char* b2s_last_error(const b2s_ctx* ctx);      /* text of the last failure on this ctx */
const char* b2s_version(void);
/* sizes in bytes for the ctx's curve: [0]=Fr, [1]=Fq, [2]=G1 affine, [3]=G2 affine, [4]=G1 xyzz, [5]=G2 xyzz */
int32_t b2s_sizes(const b2s_ctx* ctx, uint32_t out[6]);
/* kernel launches issued by this ctx since creation (bench.py's gpu_launches) */
uint64_t b2s_launch_count(const b2s_ctx* ctx);
/* block until all work queued by this ctx is done */
int32_t b2s_sync(b2s_ctx* ctx);
/* the CUDA stream (cudaStream_t) the ctx launches on, for CUDA-event timing by the harness */
void* b2s_stream(b2s_ctx* ctx);

/* Per-kernel device timing: when enabled, every kernel launch of this ctx is bracketed by CUDA events on
 * the ctx stream.  b2s_profile_report synchronises, writes one line per kernel name
 * ("<name>\t<launches>\t<total_ms>\n", NUL terminated, truncated to cap) and clears the records. */
int32_t b2s_profile_enable(b2s_ctx* ctx, int32_t on);
int32_t b2s_profile_report(b2s_ctx* ctx, char* buf, uint64_t cap);

/* ---- K2: radix-2 NTT over Fr (ark-poly Radix2EvaluationDomain::{fft,ifft}_in_place, get_coset) ----
 * In place on 2^log_n Montgomery-form elements, natural order in and out.
 *   inverse = 0: X[i] = sum_j x[j] (c w^i)^j            inverse = 1: the inverse map (includes 1/N)
 *   coset   = 0: c = 1                                  coset = 1: c = Fr::GENERATOR (7 / 5)          */
int32_t b2s_ntt(b2s_ctx* ctx, void* data, uint32_t log_n, int32_t inverse, int32_t coset, int32_t mem);

/* ---- K4: variable-base MSM (ark-ec VariableBaseMSM::msm / msm_bigint) --------------------------
 * out = sum_i scalars[i] * bases[i], i < n.  Result written to HOST memory as one affine point.
 * scalars_mont = 1: scalars are Montgomery-form Fr (`&[Fr]`); 0: canonical integers (`into_bigint()`). */
int32_t b2s_msm_g1(b2s_ctx* ctx, const void* bases, const void* scalars, uint64_t n, int32_t scalars_mont,
                   int32_t mem, void* out_affine);
int32_t b2s_msm_g2(b2s_ctx* ctx, const void* bases, const void* scalars, uint64_t n, int32_t scalars_mont,
                   int32_t mem, void* out_affine);
/* Shard form for multi-GPU: same sum, returned un-normalised (X,Y,ZZ,ZZZ) to HOST so that ranks can
 * exchange partials (all-gather) and finish with b2s_g{1,2}_sum. */
int32_t b2s_msm_g1_partial(b2s_ctx* ctx, const void* bases, const void* scalars, uint64_t n, int32_t scalars_mont,
                           int32_t mem, void* out_xyzz);
int32_t b2s_msm_g2_partial(b2s_ctx* ctx, const void* bases, const void* scalars, uint64_t n, int32_t scalars_mont,
                           int32_t mem, void* out_xyzz);
/* out_affine = sum of `count` XYZZ points (HOST in, HOST out). */
int32_t b2s_g1_sum(b2s_ctx* ctx, const void* xyzz, uint32_t count, void* out_affine);
int32_t b2s_g2_sum(b2s_ctx* ctx, const void* xyzz, uint32_t count, void* out_affine);

/* ---- K1: R1CS matrices x assignment (Matrix<F>, mat_vec_mul, evaluate_constraint) --------------
 * Upload A, B, C once per circuit.  row_ptr[k] has n_rows+1 entries, col[k]/coeff[k] have row_ptr[k][n_rows]. */
int32_t b2s_r1cs_upload(b2s_ctx* ctx, uint64_t n_rows, uint64_t n_instance, uint64_t n_witness,
                        const uint64_t* const row_ptr[3], const uint32_t* const col[3],
                        const void* const coeff[3], b2s_r1cs** out);
/* The same handle built ON THE DEVICE from the constraint system's flat storage, bypassing to_matrices()
 * (constraint_system.rs:768-804 get_lc + make_row done by kernels):
 *   args[k]      n_rows Variables: the k-th argument of every R1CS constraint (predicate/mod.rs:81-94 argument_lcs[k]);
 *                a Variable is the raw u64 of utils/variable.rs:4-14 (tag << 61 | index; Zero 0, One 1, Instance 2,
 *                Witness 3, SymbolicLc 4)
 *   lc_offsets   n_lcs + 1 entries, lc_vars / lc_coeffs lc_offsets[n_lcs] entries  (gr1cs/lc_map.rs:51-56)
 *   pool         the interner's `vec` (gr1cs/field_interner.rs:19-22): pool_len Montgomery field elements,
 *                pool[0] = ONE, pool[1] = -ONE; lc_coeffs index it
 * The system must be finalized (no LC may refer to another LC), otherwise B2S_ERR_INVALID_ARG.  HOST pointers. */
int32_t b2s_r1cs_upload_lcmap(b2s_ctx* ctx, uint64_t n_rows, uint64_t n_instance, uint64_t n_witness,
                              const uint64_t* const args[3], uint64_t n_lcs, const uint64_t* lc_offsets,
                              const uint64_t* lc_vars, const uint32_t* lc_coeffs, const void* pool,
                              uint32_t pool_len, b2s_r1cs** out);
void b2s_r1cs_free(b2s_ctx* ctx, b2s_r1cs* m);
/* out_k[i] = <M_k row i, z>, i < n_rows; z has n_instance + n_witness elements.  All buffers share `mem`. */
int32_t b2s_spmv(b2s_ctx* ctx, const b2s_r1cs* m, const void* z, int32_t mem, void* out_a, void* out_b, void* out_c);
/* h = LibsnarkReduction::witness_map (SpMV -> 3 iNTT -> 3 coset NTT -> (ab-c)/Z -> coset iNTT), computed with
 * six transforms: Z is constant on the coset and deg C < N, so h = (cosetiNTT(a_coset b_coset) - iNTT(c)) / Z(g) -- the same
 * field elements for every assignment.
 * out_h receives domain_size elements (the top one is 0); domain_size = next_pow2(n_rows + n_instance). */
int32_t b2s_witness_map(b2s_ctx* ctx, const b2s_r1cs* m, const void* z, int32_t mem, void* out_h);
uint64_t b2s_r1cs_domain_size(const b2s_r1cs* m);
/* The same h computed by the DISTRIBUTED schedule of b2s_groth16_prove_group (four-step transforms, SpMV and quotient on
 * column slabs) with 2^log_ranks virtual ranks on this one GPU, the all-to-all replaced by device copies.
 * Exists so that the index algebra of the multi-GPU path is checked bit for bit on a one-GPU box; B2S_ERR_INVALID_ARG
 * when the domain cannot be cut that way (log2(domain) odd, or too few rows per rank). */
int32_t b2s_witness_map_sim(b2s_ctx* ctx, const b2s_r1cs* m, const void* z, int32_t mem, uint32_t log_ranks, void* out_h);

/* ---- Groth16 (ark-groth16 ProvingKey / create_proof_with_reduction) ------------
 * Query vectors are affine point arrays in HOST or DEVICE memory (`mem`); they are copied to the GPU.
 *   a_query, b_g1_query, b_g2_query: n_vars = n_instance + n_witness points
 *   h_query: domain_size - 1 points;  l_query: n_witness points
 * For a base-range shard (multi-GPU) pass the sub-ranges and their offsets; a full key has offsets 0.
 */
typedef struct b2s_pk_desc {
    uint64_t n_instance, n_witness, domain_size;
    const void* alpha_g1;   /* 1 G1 affine */
    const void* beta_g1;
    const void* delta_g1;
    const void* beta_g2;    /* 1 G2 affine */
    const void* delta_g2;
    const void* a_query;    uint64_t a_off, a_len;     /* indices into the full vector held by this shard */
    const void* b_g1_query; uint64_t b1_off, b1_len;
    const void* b_g2_query; uint64_t b2_off, b2_len;
    const void* h_query;    uint64_t h_off, h_len;
    const void* l_query;    uint64_t l_off, l_len;
} b2s_pk_desc;
int32_t b2s_pk_upload(b2s_ctx* ctx, const b2s_pk_desc* desc, int32_t mem, b2s_pk** out);
void b2s_pk_free(b2s_ctx* ctx, b2s_pk* pk);

/* CircuitSpecificSetupSNARK::setup (snark/src/lib.rs:84-93) on the GPU for uploaded matrices.  `trapdoor` = 5 Montgomery
 * Fr on the HOST: tau, alpha, beta, gamma, delta, drawn by the caller from its rng (as ark-groth16's generator does).
 * Returns the device-resident proving key and writes the verifying-key elements to the HOST: alpha_g1 (G1),
 * beta_g2, gamma_g2, delta_g2 (G2), gamma_abc_g1 (n_instance G1 points). */
int32_t b2s_groth16_setup(b2s_ctx* ctx, const b2s_r1cs* m, const void* trapdoor, b2s_pk** out_pk, void* out_alpha_g1,
                          void* out_beta_g2, void* out_gamma_g2, void* out_delta_g2, void* out_gamma_abc_g1);
/* Copy one vector of a device-resident key to the HOST (to serialise a ProvingKey):
 * which = 0 a_query, 1 b_g1_query, 2 b_g2_query, 3 h_query, 4 l_query, 5 [alpha,beta,delta]_g1, 6 [beta,delta]_g2. */
int32_t b2s_pk_query(b2s_ctx* ctx, const b2s_pk* pk, int32_t which, void* out, uint64_t cap_bytes);

/* One proof on one GPU (full key).  z_instance (n_instance, z[0] = 1), z_witness, r, s: Montgomery Fr, HOST.
 * Outputs (HOST): A (G1 affine), B (G2 affine), C (G1 affine) -- `Proof { a, b, c }`. */
int32_t b2s_groth16_prove(b2s_ctx* ctx, const b2s_pk* pk, const b2s_r1cs* m, const void* z_instance,
                          const void* z_witness, const void* r, const void* s, void* out_a_g1, void* out_b_g2,
                          void* out_c_g1);
/* Same, with z = instance || witness already resident on the GPU (n_instance + n_witness elements). */
int32_t b2s_groth16_prove_resident(b2s_ctx* ctx, const b2s_pk* pk, const b2s_r1cs* m, const void* z_dev, const void* r,
                                   const void* s, void* out_a_g1, void* out_b_g2, void* out_c_g1);
/* Shard step for multi-GPU: computes this shard's five MSM partial sums
 *   out_partials = [ h_acc, l_acc, a_acc, b1_acc ] (4 G1 XYZZ) and out_b2_partial (1 G2 XYZZ), HOST.
 * r, s are needed here too: the shard that owns the end of a query range folds r*delta / s*delta into its MSMs. */
int32_t b2s_groth16_prove_shard(b2s_ctx* ctx, const b2s_pk* pk, const b2s_r1cs* m, const void* z_instance,
                                const void* z_witness, const void* r, const void* s, void* out_g1_partials,
                                void* out_g2_partial);
/* Same with z resident on the GPU. */
int32_t b2s_groth16_prove_shard_resident(b2s_ctx* ctx, const b2s_pk* pk, const b2s_r1cs* m, const void* z_dev, const void* r,
                                         const void* s, void* out_g1_partials, void* out_g2_partial);
/* Join: sums the per-shard partials (n_shards x 4 G1 XYZZ, n_shards G2 XYZZ) and applies the r/s epilogue. */
int32_t b2s_groth16_finish(b2s_ctx* ctx, const b2s_pk* pk, const void* g1_partials, const void* g2_partials,
                           uint32_t n_shards, const void* r, const void* s, void* out_a_g1, void* out_b_g2,
                           void* out_c_g1);

/* ---- multi-GPU group (SURVEY 8(b) `b2s_ctx_create(curve, n_gpus)` / 8(e)): one rank per GPU, NCCL inside -------------
 * So that ONE `SNARK::prove` (snark/src/lib.rs:50-54) drives every GPU of a box: each rank (process, or host thread with
 * its own ctx) holds a base-range SHARD of the proving key (b2s_pk_upload with offsets), the same matrices and the same z;
 * the call computes the rank's five MSM partial sums, all-gathers them with NCCL on the ctx stream (device buffers, 1.2 KiB
 * per rank -- EC addition is not an NCCL reduction) and rank 0 joins them and applies the r/s epilogue.
 * NCCL (libnccl.so.2) is bound at run time; B2S_ERR_NCCL if it is missing or fails.  world == 1 needs no NCCL.
 *   b2s_group_unique_id   rank 0 draws the NCCL id (128 bytes) and hands it to the other ranks by the host program's
 *                         own channel (torch.distributed / MPI / a file);
 *   b2s_group_create      collective over all ranks;
 *   b2s_groth16_prove_group[_resident]   collective; the proof is written on rank 0 (outputs may be NULL elsewhere). */
typedef struct b2s_group b2s_group;
#define B2S_GROUP_ID_BYTES 128
int32_t b2s_group_unique_id(uint8_t out[B2S_GROUP_ID_BYTES]);
int32_t b2s_group_create(b2s_ctx* ctx, const uint8_t id[B2S_GROUP_ID_BYTES], int32_t rank, int32_t world, b2s_group** out);
void b2s_group_destroy(b2s_group* group);
int32_t b2s_groth16_prove_group(b2s_group* group, const b2s_pk* pk_shard, const b2s_r1cs* m, const void* z_instance,
                                const void* z_witness, const void* r, const void* s, void* out_a_g1, void* out_b_g2,
                                void* out_c_g1);
int32_t b2s_groth16_prove_group_resident(b2s_group* group, const b2s_pk* pk_shard, const b2s_r1cs* m, const void* z_dev,
                                         const void* r, const void* s, void* out_a_g1, void* out_b_g2, void* out_c_g1);

/* ---- wire format: CanonicalSerialize::serialize_compressed of group elements / Proof --------
 * (snark/src/lib.rs:25-36 bounds).  BLS12-381: zcash/IETF big-endian form, 48 B (G1) / 96 B (G2); BN254: ark-ec
 * SWFlags little-endian form, 32 B / 64 B.  HOST affine Montgomery points in, bytes out; `cap` = size of `out`. */
int32_t b2s_serialize_g1_compressed(b2s_ctx* ctx, const void* affine, uint32_t count, uint8_t* out, uint64_t cap);
int32_t b2s_serialize_g2_compressed(b2s_ctx* ctx, const void* affine, uint32_t count, uint8_t* out, uint64_t cap);
/* Proof { a, b, c } -> a || b || c (192 B on BLS12-381, 128 B on BN254). */
int32_t b2s_proof_serialize_compressed(b2s_ctx* ctx, const void* a_g1, const void* b_g2, const void* c_g1, uint8_t* out,
                                       uint64_t cap);

/* serialize_uncompressed of the same types: x || y in the curve's byte / component order (96 / 192 B on BLS12-381 with only
 * the infinity bit in byte 0; 64 / 128 B on BN254 with both SWFlags in the last byte). */
int32_t b2s_serialize_g1_uncompressed(b2s_ctx* ctx, const void* affine, uint32_t count, uint8_t* out, uint64_t cap);
int32_t b2s_serialize_g2_uncompressed(b2s_ctx* ctx, const void* affine, uint32_t count, uint8_t* out, uint64_t cap);
int32_t b2s_proof_serialize_uncompressed(b2s_ctx* ctx, const void* a_g1, const void* b_g2, const void* c_g1, uint8_t* out,
                                         uint64_t cap);
/* ark-groth16 key framing (snark/src/lib.rs:25-36: ProvingKey / VerifyingKey are CanonicalSerialize):
 *   VerifyingKey = alpha_g1 || beta_g2 || gamma_g2 || delta_g2 || Vec(gamma_abc_g1)       (Vec = u64 LE length + elements)
 *   ProvingKey   = VerifyingKey || beta_g1 || delta_g1 || Vec(a_query) || Vec(b_g1_query) || Vec(b_g2_query) ||
 *                  Vec(h_query) || Vec(l_query)
 * The vk elements are HOST affine points (what b2s_groth16_setup wrote); the proving key is the device-resident FULL key,
 * streamed through the GPU in chunks (canonical form and sign bits on the device, byte order on the host).
 * compressed = 1 / 0 selects serialize_compressed / serialize_uncompressed.  The *_size functions give the exact length. */
uint64_t b2s_vk_serialized_size(const b2s_ctx* ctx, uint64_t n_gamma_abc, int32_t compressed);
int32_t b2s_vk_serialize(b2s_ctx* ctx, const void* alpha_g1, const void* beta_g2, const void* gamma_g2, const void* delta_g2,
                         const void* gamma_abc_g1, uint64_t n_gamma_abc, int32_t compressed, uint8_t* out, uint64_t cap);
uint64_t b2s_pk_serialized_size(const b2s_ctx* ctx, const b2s_pk* pk, uint64_t vk_len, int32_t compressed);
int32_t b2s_pk_serialize(b2s_ctx* ctx, const b2s_pk* pk, const uint8_t* vk_bytes, uint64_t vk_len, int32_t compressed,
                         uint8_t* out, uint64_t cap);

/* ---- setup helper: fixed-base batch multiplication -------------------------
 * out[i] = scalars[i] * G (the curve's standard generator), affine, i < n.  Used to build proving keys
 * (a_query[j] = A_j(tau) G1, ...) and synthetic bases on the GPU.  scalars: Fr; all buffers share `mem`. */
int32_t b2s_fixed_base_g1(b2s_ctx* ctx, const void* scalars, uint64_t n, int32_t scalars_mont, int32_t mem, void* out);
int32_t b2s_fixed_base_g2(b2s_ctx* ctx, const void* scalars, uint64_t n, int32_t scalars_mont, int32_t mem, void* out);

/* ---- element-wise polynomial kernels: universal-setup (Marlin-style) path, SURVEY 8(f) row 4 -----
 * The reference declares that path as a trait only (UniversalSetupSNARK, snark/src/lib.rs:107-133: universal_setup / index,
 * then SNARK::prove, lib.rs:50-54); ark-marlin's AHP is generic over ark-poly-commit's `PolynomialCommitment` and ark-poly's
 * `EvaluationDomain`.  A binding accelerates those two seams: commit / open = b2s_msm_g1 over the SRS (itself
 * b2s_fixed_base_g1 of the powers of tau), every fft / ifft / coset form = b2s_ntt, matrix products = b2s_spmv, and the
 * element-wise arithmetic in between (ark-poly `Evaluations` / `DensePolynomial` operators, ark-ff `batch_inversion`) = the
 * three entry points below.  Vectors: Fr in Montgomery form, `n` elements, all in `mem`; scalars (s, c, z): ONE Montgomery Fr
 * on the HOST.  Device-memory calls are queued on the ctx stream and return without synchronising (b2s_poly_eval
 * synchronises: it returns a value).
 *   b2s_poly_op   op 0: out = a * b   1: a + b   2: a - b   3: a * s   4: a + s   5: 1 / a with 0 -> 0 (batched inversion;
 *                 not in place).  b is ignored for ops 3-5, s for ops 0-2 and 5.  out may alias a or b for ops 0-4.
 *   b2s_poly_geom out[i] = c * s^i                      (domain elements, coset points, shifted powers)
 *   b2s_poly_eval *out = sum_i coeffs[i] z^i            (out: one Montgomery Fr on the HOST) */
int32_t b2s_poly_op(b2s_ctx* ctx, int32_t op, const void* a, const void* b, const void* s, void* out, uint64_t n, int32_t mem);
int32_t b2s_poly_geom(b2s_ctx* ctx, const void* c, const void* s, uint64_t n, int32_t mem, void* out);
int32_t b2s_poly_eval(b2s_ctx* ctx, const void* coeffs, uint64_t n, const void* z, int32_t mem, void* out);

/* ---- element-wise field kernels (unit tests of the device arithmetic; K3 building blocks) -------
 * op: 0 mul, 1 add, 2 sub, 3 inverse(a), 4 neg(a), 5 to_mont(a), 6 from_mont(a), 7 sqr(a).
 * field: 0 = Fq, 1 = Fr of the ctx's curve.  HOST buffers of `count` elements. */
int32_t b2s_field_op(b2s_ctx* ctx, int32_t field, int32_t op, const void* a, const void* b, void* out, uint64_t count);
/* group: 1 or 2.  op: 0 mixed add a+b, 1 general add, 2 double a, 3 k*a (k: one canonical Fr per point).
 * HOST affine in / HOST affine out, `count` points. */
int32_t b2s_group_op(b2s_ctx* ctx, int32_t group, int32_t op, const void* a, const void* b, const void* k, void* out,
                     uint64_t count);

#ifdef __cplusplus
}
#endif
#endif /* B200SNARK_H */
