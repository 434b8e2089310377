/*
 * b200zk.h -- C ABI of libb200zk.so: the B200 (sm_100a) backend for the one data-parallel hot
 * path under the Halo2/KZG prover and verifier of input-output-hk/plutus-halo2-verifier-gen:
 * BLS12-381 G1 multi-scalar multiplication and the Fr NTT of the evaluation domain.
 *
 * The reference has no FFI of its own for this path: its arithmetic lives in the crates
 * midnight-proofs =0.8.0 / midnight-curves =0.3.0 (-> blst), reached through the generic bound
 *   PCS: ExtractPCS + PolynomialCommitmentScheme<Scalar, Commitment = G1Projective>
 *   (/root/reference/src/plutus_gen/mod.rs:77,116,181; alias `type KZG = KZGCommitmentScheme<Bls12>`
 *   at /root/reference/examples/simple_mul.rs:31).
 * Each entry point below names the upstream function whose body it replaces and the
 * reference call sites that reach it.  INTEGRATION.md shows the Rust `extern "C"` block and
 * the PCS shim that binds them.
 *
 * Conventions
 *   - every function returns int32_t: 0 = OK, negative = error class (B200ZK_ERR_*);
 *     b200zk_last_error() returns the text of the calling thread's last failure.
 *   - no exceptions cross the boundary; entry points may be called concurrently from any host thread
 *     (rayon callers): one process drives every GPU bound by b200zk_init_devices; per GPU two calls
 *     are in flight at a time, each on its own stream, and no lock is held while a call waits for the GPU.
 *   - host-buffer MSM / NTT entry points use all bound GPUs; every other host-buffer entry point runs on
 *     the calling thread's selected GPU (b200zk_set_device); "_dev" entry points run on the GPU that owns
 *     their pointers.
 *   - there is NO CPU fallback: without a usable B200 every call fails with B200ZK_ERR_NO_DEVICE.
 *
 * Wire formats (little-endian throughout)
 *   Fr  (midnight_curves::Fq, the 255-bit scalar field):
 *        B200ZK_FMT_CANONICAL  32 bytes, integer; values >= r are reduced on read, as the proof
 *                              wire format does (/root/reference/aiken-verifier/aiken_halo2/lib/transcript.ak:158-179)
 *        B200ZK_FMT_MONT       32 bytes = 4 x u64 limbs of a*2^256 mod r: the in-memory form of
 *                              blst_fr / midnight_curves::Fq, for a zero-copy Rust shim
 *   G1 affine (midnight_curves::G1Affine): x || y, 48 bytes each
 *        B200ZK_FMT_CANONICAL  integers < p;  B200ZK_FMT_MONT  6 x u64 limbs of a*2^384 mod p
 *                              (blst_p1_affine)
 *        the identity is (0, 0) in both (/root/reference/plinth-verifier/plutus-halo2/src/Plutus/Crypto/Halo2/CompressUncompress.hs:72)
 *   compressed G1: 48 bytes big-endian x with flag bits 0x80 compressed / 0x40 infinity /
 *        0x20 y-is-larger (/root/reference/aiken-verifier/aiken_halo2/lib/bls_utils.ak:17-49)
 *   "_dev" entry points take device pointers (16-byte aligned) and a cudaStream_t passed as void*.
 */
#ifndef B200ZK_H
#define B200ZK_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200ZK_OK 0
#define B200ZK_ERR_INVALID_ARG (-1)
#define B200ZK_ERR_CUDA (-2)
#define B200ZK_ERR_NO_DEVICE (-3)
#define B200ZK_ERR_BAD_HANDLE (-4)
#define B200ZK_ERR_OOM (-5)
#define B200ZK_ERR_NOT_INIT (-6)
#define B200ZK_ERR_BAD_POINT (-7) /* a base is not a canonical point of the curve */

#define B200ZK_FMT_CANONICAL 0u
#define B200ZK_FMT_MONT 1u
/* OR-ed into the `fmt` of b200zk_bases_register*: keep only the points themselves.  By default a table of
 * n >= 1024 points is expanded at registration into W rows 2^(c*w) * P_i (W = ceil(256/c) windows), which
 * costs (W-1) * n * 96 bytes of HBM once and removes the per-window bucket sets and the window combine
 * from every later MSM against it.                                                                     */
#define B200ZK_BASES_NO_WINDOW_TABLES 0x100u
/* With several GPUs bound: partition the table by point range (every GPU holds n / G points; each MSM is reduced
 * to one partial sum per GPU and the partials meet on the first GPU through peer stores over NVLink), or keep a
 * full copy on every GPU (the columns of a batch are dealt out, no exchange).  Default: replicate tables of up to
 * 2 GiB including window rows (n <= 2^20), partition larger ones.                                           */
#define B200ZK_BASES_SHARD 0x200u
#define B200ZK_BASES_REPLICATE 0x400u

/* NTT flags */
#define B200ZK_NTT_INVERSE_SCALE 1u /* multiply the result by 1/n (caller passes omega^-1) */
#define B200ZK_NTT_COSET_IN 2u      /* x[i] *= shift^i before the transform  (coeff_to_extended) */
#define B200ZK_NTT_COSET_OUT 4u     /* y[i] *= shift^i after the transform   (extended_to_coeff, shift = g^-1) */
#define B200ZK_NTT_MONT 8u          /* data is in Montgomery form (in and out); omega/shift stay canonical */

/* ---- lifetime -------------------------------------------------------------------------- */
/* Binds the process to the listed GPUs (1..16 CUDA ordinals; NULL = 0 .. n-1).  The reference prover is ONE process
 * (/root/reference/examples/simple_mul.rs:39-141; commits at :62,72): every later host-buffer MSM / NTT call
 * fans out over all of them.  With more than one GPU every pair must have peer access (NVLink / NVSwitch).
 * A second call is a no-op until b200zk_shutdown.                                                            */
int32_t b200zk_init_devices(const int32_t *device_ids, int32_t n_devices);
/* One GPU (also what each rank of a one-process-per-GPU job calls).  device < 0 selects the current device. */
int32_t b200zk_init(int32_t device);
int32_t b200zk_shutdown(void);
int32_t b200zk_device_count(void);
/* Selects, for the calling thread, which bound GPU (index into the b200zk_init_devices list) serves the entry
 * points that neither fan out nor take a device pointer (b200zk_dev_alloc, b200zk_msm_g1_adhoc, gate-program
 * creation, ...).  Default 0.                                                                               */
int32_t b200zk_set_device(int32_t index);
int32_t b200zk_last_error(char *buf, size_t len);
/* name, SM count and compute capability of the bound device, e.g. "NVIDIA B200 sm_100 148SM" */
int32_t b200zk_device_info(char *buf, size_t len);
/* pinned host memory, so that the host<->device copies inside the entry points run at link speed (a 2^24-point MSM:
 * 77 ms end to end from pinned scalars, 101 ms from pageable ones; either way large scalar vectors are streamed in two
 * pieces so that most of the copy runs under the first piece's bucket accumulation) */
int32_t b200zk_host_alloc(void **out, size_t bytes);
int32_t b200zk_host_free(void *p);

/* ---- base tables ------------------------------------------------------------------------
 * Replaces nothing; this is the device residency of ParamsKZG::{g, g_lagrange}
 * (/root/reference/src/kzg_params.rs:33-80): upload once, commit many times.
 * stride_bytes = distance between consecutive points in the source (>= 96, multiple of 4; 0 = 96),
 * which lets a shim pass a slice of a wider Rust struct without repacking.               */
int32_t b200zk_bases_register(const uint8_t *g1_affine, uint64_t n, uint32_t fmt, uint32_t stride_bytes,
                              uint64_t *out_handle);
int32_t b200zk_bases_register_dev(const void *d_g1_affine, uint64_t n, uint32_t fmt, uint32_t stride_bytes,
                                  uint64_t *out_handle);
int32_t b200zk_bases_release(uint64_t handle);
/* copies points [start, start+n) of a table back to the host in canonical wire format */
int32_t b200zk_bases_read(uint64_t handle, uint64_t start, uint64_t n, uint8_t *out_affine);

/* ---- G1 MSM -----------------------------------------------------------------------------
 * Replaces midnight_curves G1Projective::multi_exp / blst p1s_mult_pippenger under
 *   KZGCommitmentScheme::commit_lagrange  (advice, lookup, permutation, instance columns:
 *       /root/reference/examples/simple_mul.rs:62,72; /root/reference/src/circuits/atms_circuit.rs:247,296;
 *       /root/reference/examples/schnorr.rs:115)
 *   KZGCommitmentScheme::commit           (h pieces, random poly, f and pi of multi_open:
 *       order pinned by /root/reference/src/plutus_gen/extraction/pcs/kzg.rs:55-79)
 * result = sum_{i<n} scalars[i] * bases[offset + i], affine canonical wire format (96 bytes).  */
int32_t b200zk_msm_g1(uint64_t bases, uint64_t offset, const uint8_t *scalars, uint64_t n, uint32_t scalar_fmt,
                      uint8_t out_affine[96]);
/* `batch` scalar vectors of the same length n against one table: the prover's pattern
 * (one MSM per committed column); scalars = batch*n elements, out = batch*96 bytes.          */
int32_t b200zk_msm_g1_batch(uint64_t bases, uint64_t offset, const uint8_t *scalars, uint64_t n, uint32_t batch,
                            uint32_t scalar_fmt, uint8_t *out_affine);
/* The same with one pointer per column: halo2 keeps every polynomial in its own Vec, so the shim passes the Vecs'
 * addresses and no host-side gather is needed.                                                              */
int32_t b200zk_msm_g1_batch_ptrs(uint64_t bases, uint64_t offset, const uint8_t *const *scalars, uint64_t n, uint32_t batch,
                                 uint32_t scalar_fmt, uint8_t *out_affine);
/* How a table is laid out over the bound GPUs: shard k holds points [start[k], start[k] + n[k]) on GPU index device[k]
 * (a replicated table reports every shard as [0, n)).  Arrays of `cap` entries; any output may be NULL.           */
int32_t b200zk_bases_layout(uint64_t bases, uint32_t cap, uint32_t *out_n_shards, int32_t *out_device, uint64_t *out_start,
                            uint64_t *out_n, uint32_t *out_replicated);
/* One full-length MSM against a table partitioned over several GPUs, with the scalars already resident: slice k (the
 * scalars of shard k's points, 16-byte aligned) lives in the HBM of shard k's GPU.  The result comes back to the host. */
int32_t b200zk_msm_g1_sharded_dev(uint64_t bases, const void *const *d_scalar_slices, uint32_t n_slices, uint32_t scalar_fmt,
                                  uint8_t out_affine[96]);
/* Bases that are not resident: the verifier's DualMSM left/right sums
 * (Guard::verify / DualMSM::check, /root/reference/examples/simple_mul.rs:98-102,
 * /root/reference/examples/ivc.rs:196; batch_verify, /root/reference/src/circuits/schnorr_circuit.rs:224).     */
int32_t b200zk_msm_g1_adhoc(const uint8_t *g1_affine, uint32_t point_fmt, const uint8_t *scalars, uint32_t scalar_fmt,
                            uint64_t n, uint8_t out_affine[96]);
/* Device-resident variant: scalars already in HBM, result (Montgomery affine, 96 bytes) written
 * to d_out_mont and/or (canonical) to d_out_canon; asynchronous on `stream`.                  */
int32_t b200zk_msm_g1_dev(uint64_t bases, uint64_t offset, const void *d_scalars, uint64_t n, uint32_t batch,
                          uint32_t scalar_fmt, void *d_out_mont, void *d_out_canon, void *stream);
/* Point-range sharded MSM (one process per GPU): each rank computes the sum over its slice of the table
 * and leaves it un-normalised (extended Jacobian X,Y,ZZ,ZZZ in Montgomery form, 192 bytes) in HBM; the
 * ranks all-gather the 192-byte partials over NVLink and every rank folds them with one call.  Skipping
 * the per-rank affine normalisation takes one field inversion off the critical path.                    */
int32_t b200zk_msm_g1_partial_dev(uint64_t bases, uint64_t offset, const void *d_scalars, uint64_t n, uint32_t scalar_fmt,
                                  void *d_out_xyzz, void *stream);
int32_t b200zk_g1_sum_partials_dev(const void *d_partials_xyzz, uint32_t n, void *d_out_mont, void *d_out_canon,
                                   void *stream);
/* The same exchange without a collective library (one process per GPU): rank 0 creates the meeting point in its HBM
 * and hands the 64-byte CUDA-IPC handle to the other ranks, which map it.  b200zk_msm_g1_xchg_dev then runs the local
 * MSM, stores the XYZZ partial into rank 0's memory from inside its last kernel and bumps an arrival counter with a
 * system-scope atomic; the rank that arrives last folds the partials and normalises; every rank's call ends with a
 * short kernel that waits (bounded, ~4 s) for that result and copies it into d_out_*.  Every part must make the same
 * sequence of calls.  b200zk_xchg_status reports whether a wait ever timed out (a peer never arrived).           */
int32_t b200zk_xchg_create(uint32_t n_parts, uint8_t out_ipc_handle[64], uint64_t *out_xchg);
int32_t b200zk_xchg_open(const uint8_t ipc_handle[64], uint32_t n_parts, uint32_t part, uint64_t *out_xchg);
int32_t b200zk_xchg_close(uint64_t xchg);
int32_t b200zk_xchg_status(uint64_t xchg, uint32_t *out_timed_out);
int32_t b200zk_msm_g1_xchg_dev(uint64_t bases, uint64_t offset, const void *d_scalars, uint64_t n, uint32_t scalar_fmt,
                               uint64_t xchg, void *d_out_mont, void *d_out_canon, void *stream);
/* host-buffer form: this rank's slice of the scalars streams up in two pieces like b200zk_msm_g1; synchronous */
int32_t b200zk_msm_g1_xchg(uint64_t bases, uint64_t offset, const uint8_t *scalars, uint64_t n, uint32_t scalar_fmt,
                           uint64_t xchg, uint8_t out_affine[96]);
/* Sum of n affine points: combines the per-GPU partial results of a point-range sharded MSM
 * (Montgomery affine in HBM, e.g. the all-gathered d_out_mont of every rank).                */
int32_t b200zk_g1_sum_dev(const void *d_points_mont, uint32_t n, void *d_out_mont, void *d_out_canon, void *stream);

/* ---- Fr NTT -----------------------------------------------------------------------------
 * Replaces midnight_proofs best_fft / EvaluationDomain::{lagrange_to_coeff, coeff_to_lagrange,
 * coeff_to_extended, extended_to_coeff} (constructed inside keygen/create_proof; explicit at
 * /root/reference/examples/ivc.rs:109, /root/reference/src/circuits/ivc_circuit.rs:305).
 * X[k] = sum_i x[i] omega^(ik), natural order in and out, in place; n = 2^log_n.
 * omega and coset_shift are canonical 32-byte Fr; coset_shift may be NULL without COSET flags. */
int32_t b200zk_ntt_fr(uint8_t *data, uint32_t log_n, const uint8_t omega[32], uint32_t flags,
                      const uint8_t coset_shift[32]);
/* `batch` polynomials stored back to back (batch * 2^log_n elements) */
int32_t b200zk_ntt_fr_batch(uint8_t *data, uint32_t batch, uint32_t log_n, const uint8_t omega[32], uint32_t flags,
                            const uint8_t coset_shift[32]);
/* one pointer per polynomial (separate Vecs); with several GPUs bound the polynomials are dealt out in blocks */
int32_t b200zk_ntt_fr_batch_ptrs(uint8_t *const *data, uint32_t batch, uint32_t log_n, const uint8_t omega[32], uint32_t flags,
                                 const uint8_t coset_shift[32]);
int32_t b200zk_ntt_fr_dev(void *d_data, uint32_t batch, uint32_t log_n, const uint8_t omega[32], uint32_t flags,
                          const uint8_t coset_shift[32], void *stream);
/* One long transform over all bound GPUs (2, 4, 8 or 16).  b200zk_ntt_fr does this by itself for log_n >= 23: four-step
 * decomposition n = R * C (R = 2^log_r <= 2^11); GPU g transforms the columns [g C/G, (g+1) C/G) and stores row k1 of the
 * intermediate matrix, twiddled, straight into the HBM of the GPU that owns it (the one exchange, peer stores from inside the
 * kernel); then every GPU transforms its R/G rows.  From host memory both transpositions ride on the strided H2D / D2H copies.
 * The resident form takes and leaves the blocks in HBM: d_in[g] = x[i1 * C + g C/G + u] as [R][C/G]; d_out[g] (may be d_in[g])
 * receives X[g R/G + k1 + R k2] as [C][R/G].  Synchronous.  b200zk_ntt_sharded_layout reports log_r and log_c.            */
int32_t b200zk_ntt_sharded_layout(uint32_t log_n, uint32_t *out_log_r, uint32_t *out_log_c);
int32_t b200zk_ntt_fr_sharded_dev(void *const *d_in, void *const *d_out, uint32_t n_parts, uint32_t log_n, const uint8_t omega[32],
                                  uint32_t flags, const uint8_t coset_shift[32]);

/* ---- encodings --------------------------------------------------------------------------
 * ZCash compressed form of an affine canonical point: what the transcript absorbs and the proof
 * carries (/root/reference/aiken-verifier/aiken_halo2/lib/transcript.ak:62-83,121-156).  Host-side byte logic. */
int32_t b200zk_g1_compress(const uint8_t affine[96], uint8_t out[48]);

/* =============================================================================================
 * Components next to the hot path: columns stay resident in HBM between the NTTs and
 * the commitments.  All Fr vectors below are DEVICE arrays of 32-byte elements in Montgomery form (the
 * in-memory form of midnight_curves::Fq; b200zk_fr_convert_dev converts); host scalars are canonical.
 * ============================================================================================= */

/* ---- device memory, for a shim that does not link the CUDA runtime ------------------------------ */
int32_t b200zk_dev_alloc(void **out, size_t bytes);
int32_t b200zk_dev_free(void *p);
int32_t b200zk_dev_upload(void *d_dst, const void *src, size_t bytes);   /* synchronous */
int32_t b200zk_dev_download(void *dst, const void *d_src, size_t bytes); /* synchronous, after all streams */

/* ---- batched G1 decompression: the front end of (batched) verification ----------------------------
 * Every proof carries its commitments as 48-byte compressed points
 * (/root/reference/aiken-verifier/aiken_halo2/lib/transcript.ak:62-83; y = (x^3+4)^((p+1)/4),
 * /root/reference/plinth-verifier/plutus-halo2/src/Plutus/Crypto/Halo2/CompressUncompress.hs:70-100).
 * status[i]: 0 ok, 1 not a compressed encoding, 2 bad infinity encoding, 3 x >= p, 4 x not on the curve.
 * The host variant returns B200ZK_ERR_BAD_POINT if any status is non-zero (status may be NULL);
 * rejected points are written as (0,0).  No subgroup check, like the in-tree uncompress.             */
int32_t b200zk_g1_decompress_batch(const uint8_t *compressed, uint64_t n, uint8_t *out_affine, uint32_t *status);
int32_t b200zk_g1_decompress_dev(const void *d_compressed, uint64_t n, void *d_out_mont, void *d_out_canon, void *d_status,
                                 void *stream);

/* ---- SRS generation (ParamsKZG::unsafe_setup behind /root/reference/src/kzg_params.rs:33-80) -------
 * g[i] = s^i * G and g_lagrange[i] = L_i(s) * G, L_i the Lagrange basis of the 2^k-th roots of unity
 * generated by omega; both written as packed Montgomery affine points (96 bytes = blst_p1_affine), ready
 * for b200zk_bases_register_dev.  Either output may be NULL.  Fixed-base multiplication: 32 byte windows. */
int32_t b200zk_srs_generate_dev(const uint8_t s[32], uint32_t k, const uint8_t omega[32], void *d_g_mont,
                                void *d_g_lagrange_mont, void *stream);
/* out[i] = scalars[i] * G */
int32_t b200zk_g1_fixed_mul_dev(const void *d_scalars, uint32_t scalar_fmt, uint64_t n, void *d_out_mont, void *stream);
/* packed Montgomery affine points in HBM -> canonical wire format on the host */
int32_t b200zk_g1_export_dev(const void *d_points_mont, uint64_t n, uint8_t *out_affine);

/* ---- Fr vectors: the polynomial side of multi_open and of the permutation / lookup arguments ------- */
int32_t b200zk_fr_convert_dev(const void *d_in, void *d_out, uint64_t n, uint32_t to_mont, void *stream);
/* out[r * cols + c] = base^((row0 + r) * c), Montgomery form: the twiddle block omega^(i2 * k1) between the two
 * halves of a four-step transform (the multi-GPU NTT of dist.py multiplies by it with b200zk_fr_pointwise_dev) */
int32_t b200zk_fr_power_table_dev(const uint8_t base[32], uint64_t row0, uint64_t rows, uint64_t cols, void *d_out, void *stream);
/* `batch` polynomials of n_in coefficients -> n_out coefficients each, zero padded: the staging step of
 * EvaluationDomain::coeff_to_extended when the coefficients already live in HBM (d_out must not overlap d_in) */
int32_t b200zk_fr_extend_dev(const void *d_in, uint64_t n_in, void *d_out, uint64_t n_out, uint32_t batch, void *stream);
/* op 0: out = a*b, 1: a+b, 2: a-b, 3: a*scalar, 4: out += a*b   (in place allowed) */
int32_t b200zk_fr_pointwise_dev(uint32_t op, const void *d_a, const void *d_b, const uint8_t scalar[32], void *d_out,
                                uint64_t n, void *stream);
/* out = sum_k coeffs[k] * polys[k]: the x1 / x2 / x4 linear combinations of multi_open
 * (/root/reference/src/plutus_gen/extraction/pcs/kzg.rs:55-79); d_polys is a HOST array of device pointers */
int32_t b200zk_fr_lincomb_dev(const void *const *d_polys, const uint8_t *coeffs, uint32_t count, void *d_out, uint64_t n,
                              void *stream);
/* elementwise inverse, zeros stay zero (halo2 batch_invert); in place allowed */
int32_t b200zk_fr_batch_invert_dev(const void *d_in, void *d_out, uint64_t n, void *stream);
/* out[i] = init * prod_{j<i} in[j] (exclusive) or prod_{j<=i} (inclusive): the grand products
 * z(omega X) = z(X) * num / den of the permutation and lookup arguments; init NULL = 1; in place allowed */
int32_t b200zk_fr_running_product_dev(const void *d_in, void *d_out, uint64_t n, const uint8_t init[32], uint32_t inclusive,
                                      void *stream);
/* One sweep: eval = p(z) and quot = (p(X) - p(z)) / (X - z) (n-1 coefficients) for p of n coefficients.
 * Either output may be NULL.  d_quot must not overlap d_coeffs.                                        */
int32_t b200zk_fr_kate_div_dev(const void *d_coeffs, uint64_t n, const uint8_t z[32], void *d_quot, void *d_eval, void *stream);

/* ---- gate programs: evaluation of the quotient's numerator over the extended coset domain ------------
 * (upstream plonk evaluation of h(X); consumed as the vanishing commitments,
 * /root/reference/src/plutus_gen/extraction/data/extraction_steps/proof.rs:76-80; the gate expressions are the ones the
 * reference walks at /root/reference/src/plutus_gen/extraction/mod.rs:81-102).
 * A program is a register machine of n_instr instructions of 4 x u32 { op | dst << 8, src a, src b, src c }:
 *   src = kind << 28 | payload;  kind 0: consts[payload]; 1: register[payload];
 *                                kind 2: columns[payload >> 12] at row + rotations[payload & 0xfff] * 2^(log_ext - log_n)
 *   op 0 add, 1 sub, 2 mul, 3 neg(a), 4 double(a), 5 square(a), 6 muladd a*b + c, 7 mov(a); at most 48 registers.
 * One launch evaluates the program at every row of the 2^log_ext domain; the last destination register, times
 * t_inv[row mod 2^log_period] when t_inv is given (1/(X^n - 1) on the coset), is written (or added) to out. */
#define B200ZK_GATE_MAX_REGS 48u
int32_t b200zk_gate_program_create(const uint32_t *program, uint32_t n_instr, const uint8_t *consts, uint32_t n_consts,
                                   const int32_t *rotations, uint32_t n_rotations, const uint8_t *t_inv, uint32_t log_period,
                                   uint32_t n_columns, uint32_t log_n, uint32_t log_ext, uint64_t *out_handle);
int32_t b200zk_gate_program_set_const(uint64_t handle, uint32_t index, const uint8_t value[32]); /* challenges */
int32_t b200zk_gate_program_run_dev(uint64_t handle, const void *const *d_columns, void *d_out, uint32_t accumulate,
                                    void *stream);
int32_t b200zk_gate_program_release(uint64_t handle);

/* =============================================================================================
 * The flows above the kernels, composed inside the library (one call per opening / per batch of proofs).
 * ============================================================================================= */

/* ---- Fiat-Shamir transcript: CardanoFriendlyBlake2b ------------------------------------------------
 * Unkeyed blake2b-256 over the whole absorbed history: 0x01 || item per absorbed scalar (32 B LE) or point (48 B
 * compressed), 0x00 per squeeze; challenge = LE(H) + LE(H(H)) * 2^256 mod r
 * (/root/reference/src/plutus_gen/adjusted_types/mod.rs:30-72; /root/reference/aiken-verifier/aiken_halo2/lib/transcript.ak:19-106).
 * Host-side byte logic (no GPU needed).  A verifier "reads" a proof item by absorbing the same bytes.          */
int32_t b200zk_transcript_new(uint64_t *out_transcript);
int32_t b200zk_transcript_free(uint64_t transcript);
int32_t b200zk_transcript_common_scalar(uint64_t transcript, const uint8_t scalar[32]);
int32_t b200zk_transcript_common_point(uint64_t transcript, const uint8_t compressed[48]);
int32_t b200zk_transcript_squeeze(uint64_t transcript, uint8_t out_challenge[32]);

/* ---- KZGCommitmentScheme::multi_open, prover side, polynomials resident in HBM ------------------------
 * The halo2 multi-open argument in the message order the reference pins
 * (/root/reference/src/plutus_gen/extraction/pcs/kzg.rs:55-79): squeeze x1, x2; write the commitment of
 * f = sum_s x2^s (q_s - r_s) / Z_s; squeeze x3; write q_s(x3) for every point set; squeeze x4; write the commitment pi of
 * (final - final(x3)) / (X - x3), final = sum_s x4^s q_s + x4^S f.  q_s = sum_j x1^j p_{s,j} over the polynomials opened at
 * point set s; sets and their order follow precompute_intermediate_sets (/root/reference/src/plutus_gen/extraction/pcs/mod.rs:36-109).
 * d_polys: HOST array of device pointers, n Montgomery coefficients each, all on one GPU on which `bases` (the monomial
 * table g) is resident.  Query i opens polynomial query_poly[i] at query_points[32 i ..].  The transcript is advanced exactly
 * as the verifier will advance it.  out_proof receives f (48 B) || q evals (32 B each) || pi (48 B); *out_len its length.  */
int32_t b200zk_h2mo_open_dev(uint64_t bases, uint64_t transcript, const void *const *d_polys, uint32_t n_polys, uint64_t n,
                             const uint32_t *query_poly, const uint8_t *query_points, uint32_t n_queries, uint8_t *out_proof,
                             size_t cap, size_t *out_len);

/* ---- KZGCommitmentScheme::multi_prepare, verifier side -------------------------------------------------
 * Replays the transcript over the opening proof, runs the scalar pipeline of
 * /root/reference/aiken-verifier/aiken_halo2/lib/halo2_kzg.ak:46-171 (q_eval_sets, f_eval, v) on the host and returns a guard:
 * the two lazy sums  left = pi,  right = sum_s x4^s sum_j x1^j C_{s,j} + x4^S f - v G + x3 pi  (halo2_kzg.ak:15-44), to be
 * accepted iff e(left, [s]G2) == e(right, G2).  commitments: 48-byte compressed points; query i says commitment
 * query_commitment[i] is claimed to evaluate to query_evals[32 i ..] at query_points[32 i ..].  out_scalars (optional, 192 B):
 * x1, x2, x3, x4, f_eval, v.                                                                                       */
int32_t b200zk_h2mo_prepare(uint64_t transcript, const uint8_t *commitments, uint32_t n_commitments, const uint32_t *query_commitment,
                            const uint8_t *query_points, const uint8_t *query_evals, uint32_t n_queries, const uint8_t *proof,
                            size_t proof_len, uint64_t *out_guard, uint8_t *out_scalars);
/* The scalar pipeline alone with the four challenges given (x1 || x2 || x3 || x4): the known-answer surface of the reference's
 * own tests (/root/reference/plinth-verifier/plutus-halo2/src/Plutus/Crypto/Halo2/Halo2MultiOpenMSM.hs:26-42).  out_q_eval_sets:
 * set after set, the points of a set in ascending canonical order.                                                  */
int32_t b200zk_h2mo_scalars(uint32_t n_commitments, const uint32_t *query_commitment, const uint8_t *query_points,
                            const uint8_t *query_evals, uint32_t n_queries, const uint8_t challenges[128], const uint8_t *proof_q_evals,
                            uint32_t n_sets, uint8_t *out_q_eval_sets, size_t cap_q_eval_sets, uint8_t out_f_eval[32], uint8_t out_v[32]);
/* Guard::verify / batch_verify up to the pairing (/root/reference/examples/simple_mul.rs:98-102,
 * /root/reference/src/circuits/schnorr_circuit.rs:224-229): left and right sums of sum_i c_i * guard_i (challenges = NULL: one
 * guard, c = 1).  Every proof point is decompressed on the GPU in one batch and each side is one ad-hoc MSM, split by point
 * range over the bound GPUs with the partial sums meeting on the first.                                              */
int32_t b200zk_guard_eval(const uint64_t *guards, uint32_t n_guards, const uint8_t *challenges, uint8_t out_left[96],
                          uint8_t out_right[96]);
int32_t b200zk_guard_free(uint64_t guard);

/* ---- synthetic inputs and self-test (bench / tests) ------------------------------------------
 * bases P_i = a_i*G with a_i = splitmix64(seed + start + i), written as packed Montgomery affine. */
int32_t b200zk_g1_synth_bases_dev(uint64_t seed, uint64_t start, uint64_t n, void *d_out_mont, void *stream);
/* elementwise field ops on device arrays, for limb-for-limb parity tests of the arithmetic layer:
 * op 0 mul, 1 add, 2 sub, 3 inverse(a); field 0 = Fr (32 B), 1 = Fp (48 B); canonical in/out.    */
int32_t b200zk_selftest_field(uint32_t field, uint32_t op, const uint8_t *a, const uint8_t *b, uint8_t *out,
                              uint64_t count);
/* integer-pipe micro-benchmark: kind 0 = independent IMAD.WIDE.U32, (1 was removed: ptxas hoisted half of its
 * pairs, so it timed 32-bit IMADs), 2 = Fp Montgomery
 * multiplications, 3 = XYZZ mixed additions, 4 = Fr multiplications, 5 = carry-chained IMAD.WIDE.U32.X rows
 * (as the Montgomery multiplier issues them), 6 = DFMA, 7 = IMAD.WIDE with carry-out only, 8 = IMAD.WIDE
 * paired 1:1 with IADD, 9 = IMAD.HI.U32 alone, 10 = 32-bit IMAD alone, 11 = the unfused IMAD + IMAD.HI.U32 pair with an
 * immediate multiplier, 12 = Fr multiplications at 16 warps / SM (the NTT kernel's occupancy), 13 / 14 = the arithmetic of an NTT
 * butterfly (add, sub, product; no memory) at 16 / 64 warps per SM.  *out_ops_per_s receives limb-MACs (0,1,5,7,8), multiplications (2,4,12), butterflies (13,14), additions (3)
 * or FMAs (6) per second over all SMs.                                                            */
int32_t b200zk_microbench(uint32_t kind, uint32_t iters, double *out_ops_per_s, double *out_ms);
/* number of kernels this library has launched since init (bench.py's gpu_launches counter) */
uint64_t b200zk_launch_count(void);
/* Phase timing of the last MSM / NTT call, taken with CUDA events on the launching stream (no effect on the
 * work itself).  kind: 1 = MSM with phases {sort (digits, scan, scatter, tasks), accumulate, tail (collapse,
 * bucket reduce, window combine)}; 2 = NTT with one phase per pass.  Synchronises on the last event.        */
int32_t b200zk_set_profiling(uint32_t enable);
int32_t b200zk_get_profile(uint32_t *kind, double *phase_ms, uint32_t cap, uint32_t *n_phases, uint32_t *msm_window_bits,
                           uint32_t *msm_windows);
/* overrides for experiments: window_bits bits 0..7 = MSM window bits (0 = automatic; also used for tables
 * registered afterwards), bits 8..14 = 1 + accumulate-kernel code variant (0 = default), bit 15 = do not
 * build window tables; smax = max entries per task (0 = automatic)                                      */
int32_t b200zk_set_msm_tuning(uint32_t window_bits, uint32_t smax);

#ifdef __cplusplus
}
#endif
#endif /* B200ZK_H */
