/* plonk_b200 - C ABI of the B200-native backend for the dusk-plonk prover hot path.
 *
 * Each entry point replaces one crate-private function of the reference (paths relative to the
 * dusk-network/plonk checkout):
 *
 *   pb200_ntt / pb200_ntt_dev        EvaluationDomain::{fft, ifft, coset_fft, coset_ifft}
 *                                    src/fft/domain.rs:166-232  (best_fft :383-422)
 *   pb200_msm_g1 / pb200_msm_g1_dev  CommitKey::commit -> msm_variable_base + Commitment::from
 *                                    src/commitment_scheme/kzg10/key.rs:376-388,
 *                                    src/commitment_scheme/kzg10/commitment.rs:89-93
 *   pb200_srs_upload                 CommitKey { powers_of_g } made resident in HBM once per Prover
 *                                    src/commitment_scheme/kzg10/key.rs:36-41
 *   pb200_g1_compress                G1Affine::to_bytes as used by Commitment::to_bytes
 *                                    src/commitment_scheme/kzg10/commitment.rs:95-101
 *   pb200_prover_* / pb200_prove     Prover::new / Prover::prove (PlonkVersion::V3)
 *                                    src/compiler/prover.rs:53-115, 415-761
 *
 * Data layout (identical to the reference's in-memory layout):
 *   Fr  (BlsScalar)  4 x u64 little-endian limbs, Montgomery form R = 2^256        -> 32 bytes
 *   G1 affine        x then y, each 6 x u64 little-endian limbs, Montgomery R=2^384 -> 96 bytes;
 *                    the identity is encoded as x = y = 0.
 *
 * All functions return 0 on success or a negative pb200_status.  There is no CPU fallback: if no
 * CUDA device is usable every call fails with PB200_ERR_CUDA (the Rust shim turns that into a
 * panic, because the reference's Error enum has no device variant and the NTT functions are
 * infallible - src/error.rs:21-120, src/fft/domain.rs:394).
 * All entry points are thread safe; host-pointer variants are synchronous, *_dev variants enqueue
 * on the given CUDA stream (a cudaStream_t passed as void*, NULL = the calling thread's default
 * pb200 stream) and return without synchronising.
 */
#ifndef PLONK_B200_H
#define PLONK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  PB200_OK = 0,
  PB200_ERR_CUDA = -1,            /* CUDA runtime failure (no device, OOM, launch error) */
  PB200_ERR_INVALID_DOMAIN = -2,  /* log_n >= 32: Error::InvalidEvalDomainSize (domain.rs:132-137) */
  PB200_ERR_DEGREE_TOO_LARGE = -3,/* Error::PolynomialDegreeTooLarge (key.rs:362-370) */
  PB200_ERR_INVALID_ARG = -4,
  PB200_ERR_UNSATISFIED = -5,     /* Error::CircuitUnsatisfied (quotient_poly.rs:132-134) */
  PB200_ERR_NOT_READY = -6,
  /* -7 .. -9: circuit front end, plonk_b200_composer.h */
  PB200_ERR_POINT_MALFORMED = -10 /* dusk_bytes::Error::InvalidData / Error::PointMalformed: a G1 encoding that is
                                     not canonical, not on the curve or not in the prime-order subgroup */
} pb200_status;

typedef struct pb200_srs pb200_srs_t;
typedef struct pb200_prover pb200_prover_t;

/* ---- process / device ------------------------------------------------------------------- */
/* Selects the device for this process.  Idempotent for the same device; ONE device per process: the
 * twiddle / coset tables, per-thread streams and pinned staging buffers are created on the first device
 * used, so a second call naming another device returns PB200_ERR_INVALID_ARG (multi-GPU hosts run one
 * process per GPU, as bench.py does under torchrun). */
int pb200_init(int device);
const char* pb200_last_error(void);         /* thread-local description of the last failure */
int pb200_device_sync(void);
/* Number of kernels launched by this library since process start (for bench.py's gpu_launches). */
uint64_t pb200_launch_count(void);

/* ---- NTT -------------------------------------------------------------------------------- */
/* `batch` vectors; vector b reads in + b*in_stride (in_len elements, zero padded / truncated to
 * n = 2^log_n exactly as Vec::resize does at domain.rs:174) and writes n elements at
 * out + b*out_stride.  inverse: 0 = fft, 1 = ifft (includes the 1/n scaling, domain.rs:187-196).
 * coset: 0 = plain, 1 = coset variant (distribute_powers with GENERATOR = 7 before a forward
 * transform, with 7^-1 after an inverse one; domain.rs:198-232).  Strides are in elements. */
int pb200_ntt(const uint64_t* in, size_t in_len, uint64_t* out, uint32_t log_n, int inverse,
              int coset, uint32_t batch, size_t in_stride, size_t out_stride);
int pb200_ntt_dev(const uint64_t* d_in, size_t in_len, uint64_t* d_out, uint32_t log_n,
                  int inverse, int coset, uint32_t batch, size_t in_stride, size_t out_stride,
                  void* stream);

/* ---- KZG commit key + MSM ---------------------------------------------------------------- */
/* raw_points: n_points x 96 bytes (layout above) = CommitKey::powers_of_g. */
int pb200_srs_upload(const uint8_t* raw_points, size_t n_points, pb200_srs_t** out);
/* Same with an explicit bucket-window width (bits); 0 = automatic (what pb200_srs_upload picks from
 * n_points).  The ranks of a point-sharded MSM (pb200_msm_g1_allgather*) must use ONE width for all slices:
 * take pb200_msm_window_for(largest slice) on every rank. */
int pb200_srs_upload_window(const uint8_t* raw_points, size_t n_points, int window_bits, pb200_srs_t** out);
int pb200_msm_window_for(size_t n_points);    /* the automatic choice for a key of n_points */
int pb200_srs_window(const pb200_srs_t* srs); /* the width a key was uploaded with */
void pb200_srs_free(pb200_srs_t* srs);
size_t pb200_srs_len(const pb200_srs_t* srs);

/* batch commitments sum_i scalars[b][i] * powers_of_g[i], i < n_scalars (n_scalars may be smaller
 * than the key: zip semantics of msm_variable_base).  n_scalars > pb200_srs_len() returns
 * PB200_ERR_DEGREE_TOO_LARGE.  out_affine: batch x 96 bytes, normalised affine points. */
int pb200_msm_g1(const pb200_srs_t* srs, const uint64_t* scalars, size_t n_scalars, uint32_t batch,
                 size_t stride, uint64_t* out_affine);
int pb200_msm_g1_dev(const pb200_srs_t* srs, const uint64_t* d_scalars, size_t n_scalars,
                     uint32_t batch, size_t stride, uint64_t* out_affine_host, void* stream);
/* Partial MSM over the point range [first, first + n_scalars) of the key, for sharding one large
 * MSM across GPUs by points; the caller adds the per-rank results. */
int pb200_msm_g1_range(const pb200_srs_t* srs, size_t first, const uint64_t* scalars,
                       size_t n_scalars, uint64_t* out_affine);

/* One MSM whose points are partitioned across the GPUs of a box (BASELINE.json configs[3], SURVEY.md
 * section 8e-ii): every rank calls this with its slice of the commit key (uploaded with
 * pb200_srs_upload) and the matching slice of each scalar vector (host memory, `batch` vectors of
 * stride `stride`).  The per-rank partial results - one affine point per batch entry - are exchanged
 * with a single ncclAllGather on `nccl_comm` (an ncclComm_t of `n_ranks` ranks created by the caller;
 * NCCL is looked up in the process at run time) and added locally in rank order, so every rank
 * receives the same batch x 96-byte result.  This is the only collective on the path.  (Since round 2 the
 * exchange is the device-resident one described under pb200_msm_g1_allgather_dev; this entry point only adds
 * the host -> device copy of the scalars.) */
int pb200_msm_g1_allgather(const pb200_srs_t* srs_slice, const uint64_t* scalars_slice, size_t n_scalars,
                           uint32_t batch, size_t stride, void* nccl_comm, int n_ranks, uint64_t* out_affine);
/* Same with the scalar slice already resident in HBM.  The exchange is device to device on `stream`
 * (NULL = the calling thread's pb200 stream), straight behind the reduction kernels: what travels is each
 * rank's per-digit bucket sums ((ndig + 1) x 192 bytes per batch entry plus a 16-byte header), the per-digit
 * sums are added across ranks and ONE Horner + affine normalisation finishes the result on every rank; the
 * call synchronises the stream once.  A rank whose local part fails still joins the collective (flagged in
 * its header), so its peers return PB200_ERR_CUDA instead of hanging; slices uploaded with different window
 * widths are reported as PB200_ERR_INVALID_ARG on every rank.  An empty slice (n_scalars = 0) is allowed. */
int pb200_msm_g1_allgather_dev(const pb200_srs_t* srs_slice, const uint64_t* d_scalars_slice, size_t n_scalars,
                               uint32_t batch, size_t stride, void* nccl_comm, int n_ranks,
                               uint64_t* out_affine_host, void* stream);

/* The host tail of that exchange on its own (no GPU needed): `parts` holds n_parts x batch records of
 * *words_per_entry 32-bit words - the (ndig + 1) XYZZ digit sums D_0 .. D_{ndig-1}, sum A_G a partial MSM with
 * window width `window_bits` leaves (part-major) - and out_affine receives batch x 96 bytes:
 *   R = sum_parts [ A + g * sum_j 2^shift_j D_j ],  added digit by digit, then one Horner and one shared inversion.
 * With parts = NULL only *words_per_entry is written. */
int pb200_msm_combine_parts(const uint32_t* parts, int n_parts, int window_bits, uint32_t batch,
                            uint64_t* out_affine, size_t* words_per_entry);

/* PublicParameters::setup with explicit secrets (srs.rs:61-100): out[i] = [g_scalar * x^i] G1, as
 * n_points x 96-byte raw points.  Test/bench helper - a real SRS comes from a ceremony. */
int pb200_srs_setup_from_secret(const uint64_t* x, const uint64_t* g_scalar, size_t n_points,
                                uint8_t* out_raw);

/* CommitKey::from_slice / PublicParameters::from_slice (key.rs:319-326, srs.rs:163-178): decodes
 * n_points 48-byte compressed points (G1Affine::from_bytes: zcash encoding, with the on-curve and -
 * when check_subgroup != 0, as the reference always does - prime-order subgroup checks) into the
 * 96-byte raw layout pb200_srs_upload / pb200_prover_new take.  The square roots and the subgroup
 * checks run on the GPU, one thread per point.  PB200_ERR_POINT_MALFORMED names the first bad point. */
int pb200_g1_decompress(const uint8_t* compressed, size_t n_points, int check_subgroup, uint8_t* out_raw);
/* The raw ("unchecked", fast-loading) key formats of the reference - SURVEY.md section 8 row f4.
 *   CommitKey::to_raw_var_bytes (key.rs:215-229) = u64 little-endian point count, then per point
 *   G1Affine::to_raw_bytes of dusk-bls12_381 0.14: PB200_G1_RAW_SIZE = 97 bytes = x then y as 6 + 6 little-endian
 *   u64 Montgomery limbs (the first 96 bytes ARE this library's raw layout) and one byte that is 1 for the
 *   identity.  That crate is not vendored under the reference checkout and the reference holds no golden bytes
 *   for this format, so the record layout is restated from the crate's published source, not pinned by a vector.
 *   PublicParameters::to_raw_var_bytes (srs.rs:114-119) = OpeningKey::to_bytes (PB200_OPENING_KEY_BYTES = 240: a
 *   compressed G1 and two compressed G2 points, verifier material this library never reads) followed by the above.
 * checked = 0 mirrors from_slice_unchecked (key.rs:242-256, srs.rs:132-146: no point validation, at most the
 * announced count of whole records); checked = 1 mirrors CommitKey::from_raw_var_bytes (key.rs:258-298: exact
 * length, is_on_curve & is_torsion_free per point - run on the GPU; PB200_ERR_POINT_MALFORMED names the first
 * bad point).  out_raw receives n_points x 96 bytes for pb200_srs_upload / pb200_prover_new. */
#define PB200_G1_RAW_SIZE 97
#define PB200_OPENING_KEY_BYTES 240
int pb200_raw_commit_key_points(const uint8_t* bytes, size_t len, int checked, size_t* n_points);
int pb200_commit_key_from_raw_var_bytes(const uint8_t* bytes, size_t len, int checked, uint8_t* out_raw);
/* 48-byte compressed encoding of one affine point given in the 96-byte raw layout. */
int pb200_g1_compress(const uint64_t* affine_raw, uint8_t out48[48]);
/* out = a + b for two points in the 96-byte raw layout (host-side helper for multi-GPU reduction). */
int pb200_g1_add_affine(const uint64_t* a_raw, const uint64_t* b_raw, uint64_t* out_raw);

/* ---- device-resident prover (Prover::new / Prover::prove, PlonkVersion::V3) ------------------ */
/* Circuit description = what Compiler::preprocess reads from the Composer (src/compiler.rs:132-170):
 *   selectors: 11 columns x n_constraints Fr in the order q_m, q_l, q_r, q_o, q_f, q_c, q_arith,
 *              q_range, q_logic, q_fixed_group_add, q_variable_group_add (column-major);
 *   wires:     4 columns (a, b, c, d) x n_constraints witness indices;
 *   srs_raw:   PublicParameters' commit key, 96-byte raw points; it is trimmed here exactly as
 *              pp.trim(next_pow2(constraints + 6)) does (compiler.rs:121-124, srs.rs:188-196).
 * Preprocessing (15 iNTT, 15 commitments, 16 coset NTTs, sigma evaluations) runs on the GPU and
 * the prover key stays resident in HBM. */
int pb200_prover_new(const uint8_t* label, size_t label_len, size_t n_constraints,
                     const uint64_t* selectors, const uint32_t* wires, size_t n_witnesses,
                     const uint8_t* srs_raw, size_t n_srs_points, pb200_prover_t** out);
/* Prover::try_from_bytes (src/compiler/prover.rs:265-350) for the bytes of Prover::to_bytes (:236-263): six
 * big-endian u64 (label, prover-key, commit-key, verifier-key lengths, size, constraints), the label,
 * ProverKey::to_var_bytes (src/proof_system/widget.rs:347-445: n, the byte size of one Evaluations, then for
 * each of the 15 polynomials its coefficient count, canonical 32-byte coefficients and its 8n coset
 * evaluations, then the linear and vanishing-polynomial evaluations), the commit key in the raw format above
 * and VerifierKey::to_bytes (widget.rs:84-111: n and 15 compressed commitments).  The polynomials go to HBM in
 * coefficient form and the commitments are taken as stored, so none of the 15 iNTTs / 15 MSMs of preprocessing
 * runs; the 8n evaluations are recomputed by 16 coset NTTs on the device (faster than moving 17 x 8n scalars
 * over PCIe), so the serialized ones are skipped, not read.  Errors as the reference: too short ->
 * PB200_ERR_INVALID_ARG (NotEnoughBytes), inconsistent sizes / non-canonical scalars -> PB200_ERR_POINT_MALFORMED
 * (InvalidData), an invalid commit-key point -> PB200_ERR_POINT_MALFORMED (checked as try_from_bytes does).
 * A serialized Prover does not hold the circuit's wiring (the reference re-runs the circuit per proof and reads
 * the wires from that composer), so `wires` (4 x constraints witness indices) and n_witnesses come with it. */
int pb200_prover_from_bytes(const uint8_t* bytes, size_t len, const uint32_t* wires, size_t n_witnesses,
                            pb200_prover_t** out);
void pb200_prover_free(pb200_prover_t* prover);
/* 15 compressed commitments (verifier-key material) in the order of `selectors` then s_sigma_1..4. */
int pb200_prover_commitments(const pb200_prover_t* prover, uint8_t* out_15x48);
/* One proof.  witnesses: the Composer's witness table after running the circuit (n_witnesses Fr);
 * pi_idx / pi_vals: sorted public-input positions and values (Composer::public_input_indexes /
 * public_inputs, src/composer.rs:465-480); blinders: the 14 BlsScalar::random draws of
 * prove_inner in RNG order a0,a1,b0,b1,c0,c1,d0,d1, z0,z1,z2, b12,b13,b14 (prover.rs:154-161,
 * 503, 553-555) - the RNG belongs to the caller, as in the reference API.
 * out_proof: Proof::to_bytes, 1008 bytes (src/proof_system/proof.rs:137-162).
 * Returns PB200_ERR_UNSATISFIED for Error::CircuitUnsatisfied. */
int pb200_prove(const pb200_prover_t* prover, const uint64_t* witnesses, size_t n_witnesses,
                const uint64_t* pi_idx, const uint64_t* pi_vals, size_t n_pi,
                const uint64_t* blinders, uint8_t* out_proof);
/* Same with the witness table already resident in HBM (pi_* and blinders stay host pointers);
 * n_witnesses is the length of the device table and must equal the compiled circuit's.
 * Both calls return PB200_ERR_INVALID_ARG when n_pi > 0 and pi_idx / pi_vals is NULL, when a position is
 * outside the circuit, or when pi_idx is not strictly increasing (the reference's BTreeMap cannot hold
 * duplicates).  Largest circuit: 2^28 gates (the quotient domain 8n must stay below 2^32, domain.rs:132-137). */
int pb200_prove_dev(const pb200_prover_t* prover, const uint64_t* d_witnesses, size_t n_witnesses,
                    const uint64_t* pi_idx, const uint64_t* pi_vals, size_t n_pi,
                    const uint64_t* blinders, uint8_t* out_proof, void* stream);

/* ---- measurement helpers ----------------------------------------------------------------- */
/* In-library CUDA-event timing of the MSM bucket-accumulation phase (k_msm_accumulate, the dominant
 * kernel of a proof, plus the two heavy-bucket kernels behind it) on its launching stream: enable resets
 * the counters; read returns the summed time, the G1 mixed additions executed (one per non-zero window
 * digit), the launch count and the MSM points processed. */
int pb200_profile_enable(int on);
/* The prover gives its dense MSMs one lane per bucket (no merge additions) as soon as two proofs are in flight on
 * it, and splits buckets over lanes for a proof that is alone (latency).  pb200_throughput_mode(1) makes single
 * proofs use the in-flight launch shape too, so that a kernel can be timed alone in the shape it has under load;
 * pb200_throughput_mode(0) restores the automatic choice.  Results never depend on it. */
int pb200_throughput_mode(int on);
int pb200_profile_read(double* accumulate_ms, uint64_t* accumulate_adds, uint64_t* accumulate_launches,
                       uint64_t* msm_points);
/* pb200_profile_read counts the DENSE MSMs (at least a quarter of the window digits non-zero: polynomial
 * coefficients); sparse ones - the wire-value commitments against the Lagrange-form key, ~1 digit per scalar,
 * mostly the over-long-bucket kernels - are accumulated separately so that they do not blur the dominant
 * kernel's rate. */
int pb200_profile_read_sparse(double* accumulate_ms, uint64_t* accumulate_adds, uint64_t* accumulate_launches,
                              uint64_t* msm_points);
/* Lagrange form of a commit key: out[j] = [L_j(x)]G = (1/n) sum_i w^(-ij) points[i] for the first
   n = 2^k points of CommitKey::powers_of_g (reference src/commitment_scheme/kzg10/key.rs:36-41), by an
   inverse NTT over group elements on the device.  Host buffers, n affine points of 96 bytes each (the
   layout of pb200_srs_upload).  With it a polynomial can be committed through its evaluations on the
   domain (sum_i p(w^i) out[i]) instead of its coefficients; the prover does not use it yet.
   PB200_ERR_INVALID_DOMAIN unless n is a power of two. */
int pb200_g1_lagrange_key(const uint64_t* points, size_t n, uint64_t* out);
/* Register-only IMAD.WIDE microbenchmark: returns achieved 32x32+64 multiply-adds per second. */
int pb200_imad_peak(double* mads_per_sec);
/* Dependent chains of carry-chained Fp (381-bit) Montgomery products, 1024 threads per SM: achieved products
 * per second - the ceiling the G1 kernels are measured against (IMAD.WIDE.U32.X, the carry-in/out form every
 * multi-limb product needs, issues at half the rate of the carry-free IMAD.WIDE that pb200_imad_peak times). */
int pb200_fp_product_peak(double* products_per_sec);
/* Elementwise Fr / Fp Montgomery products on the device (kernel self-test of the arithmetic). */
int pb200_selftest_fr_mul(const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n);
int pb200_selftest_fp_mul(const uint64_t* a, const uint64_t* b, uint64_t* out, size_t n);
/* The three Fp product forms the G1 formulas use: out[0..n) = a*b, out[n..2n) = a^2,
   out[2n..3n) = a*b - c*d (6 x u64 Montgomery limbs each). */
int pb200_selftest_fp_ops(const uint64_t* a, const uint64_t* b, const uint64_t* c, const uint64_t* d,
                          uint64_t* out, size_t n);

#ifdef __cplusplus
}
#endif
#endif /* PLONK_B200_H */
