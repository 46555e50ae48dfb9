/*
 * tcb200.h — C ABI of the B200-native BLS12-381 threshold-crypto engine (libtcb200.so).
 *
 * This is the drop-in boundary for the ONE hot path of poanetwork/threshold_crypto
 * (SURVEY.md §8): every entry point is the batched form of a reference method and is what
 * a Rust `extern "C"` block in a patched `threshold_crypto` (or a shim crate) would bind —
 * see INTEGRATION.md for the Rust-side declarations.  Citations are file:line under
 * /root/reference.
 *
 * Conventions
 *   - Plain pointers and sizes only; the caller owns every buffer; the library keeps
 *     nothing after a call returns.  One tcb_ctx per calling thread (calls on one ctx are
 *     serialised).  There is NO CPU fallback: without a CUDA device tcb_init fails.
 *   - Return value: 0 = ok, < 0 = CUDA / argument failure (text via tcb_last_error).
 *     Per-item scheme outcomes go to status[]: 0 ok, 1 NotEnoughShares (normally caught
 *     by the caller, src/lib.rs:731-733), 2 DuplicateEntry (src/lib.rs:763; unreachable in
 *     the reference because of its by-value filter, kept for ABI completeness),
 *     3 invalid encoding (field element >= modulus).
 *   - G1 points: EXTERNAL pairing 0.16 *uncompressed* affine encoding, 96 B = x || y
 *     big-endian; G2 points: 192 B = x.c1 || x.c0 || y.c1 || y.c0; infinity = byte 0 bit
 *     0x40 set, everything else zero (src/lib.rs:89,238-240 use the same encoding).
 *     Points must be valid group elements (the Rust types guarantee it).
 *   - Scalars (Fr): canonical (non-Montgomery) 4 x u64 little-endian = 32 B, the bytes
 *     SerdeSecret / FieldWrap emit (src/serde_impl.rs:109,296).
 *   - Messages: one concatenated byte buffer + (n+1) u64 offsets.
 *   - `_dev` variants take DEVICE pointers (on ctx's first device) and a cudaStream_t
 *     passed as void*; they only enqueue work.  They exist so that a caller that already
 *     holds its batch in HBM (and bench.py's kernel-only timing) skips the copies.
 */
#ifndef TCB200_H
#define TCB200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct tcb_ctx tcb_ctx;

/* Engines of the pairing check (tcb_set_engine); both put one item on a quad of lanes and return the same booleans: */
#define TCB_ENGINE_QUAD_REG 1  /* round-1 kernel: Miller loop + final exponentiation fused, Fp12 register-resident, exchanges by warp shuffles */
#define TCB_ENGINE_QUAD_SMEM 2 /* default: Miller loop (k_miller_quad) and final exponentiation (k_final_exp_sm) with their operands staged
                                  in shared memory (dot-product form, TMA bulk input staging), f through HBM between the two kernels */
#define TCB_ENGINE_QUAD_SMEM_REGFE 3 /* the shared-memory Miller loop followed by round 1's register-engine final exponentiation
                                  (k_final_exp_quad): 32.2 instead of 28.5 ms per 2^16 and 4.1 GB instead of 0.35 GB of DRAM write-back;
                                  kept as the self-test reference and for A/B measurements */

int tcb_init(tcb_ctx **ctx, const int *device_ids, int n_devices);
void tcb_free(tcb_ctx *ctx);
const char *tcb_last_error(const tcb_ctx *ctx);
int tcb_set_engine(tcb_ctx *ctx, int engine);
/* How tcb_verify_batch (PublicKey::verify, src/lib.rs:115-117) forms the message point.  0 (default): the hash kernel stops at
 * Q0 = [3 (x^2 - 1)] H(m) (two of the three 64-bit multiplications of the exact cofactor clearing) and the check is
 * e(pk, Q0) == e([3 (x^2 - 1)] g1, sig) with the constant multiple of the generator — 3 (x^2 - 1) is a unit mod r, so this is
 * the same boolean as e(pk, H(m)) == e(g1, sig).  1: the exact H(m) of tcb_hash_g2_batch and the plain generator (A/B
 * measurement, cross-check in the tests).  tcb_hash_g2_batch / tcb_sign_batch always produce the exact point. */
int tcb_set_verify_hash(tcb_ctx *ctx, int exact);
/* hash_g2 (tcb_hash_g2_batch and inside tcb_verify_batch): 0 (default) two kernels — SHA3 / ChaCha / candidate search / square
 * root with one THREAD per item (k_hash_g2_point), cofactor clearing on lane pairs (k_g2_clear); 1 the one-kernel lane-pair
 * version (k_hash_g2), in which the two Fp powers of the square root run redundantly on both lanes.  Same outputs. */
int tcb_set_hash_algo(tcb_ctx *ctx, int algo);
/* Tuning knob of the multi-scalar multiplication behind combine / decrypt / lincomb: partial sums per
 * item (shared doublings vs. parallelism).  0 (default) = chosen per call from the batch shape. */
int tcb_set_msm_groups(tcb_ctx *ctx, size_t groups);
/* Algorithm of that multi-scalar multiplication: 0 (default) Straus with shared doublings and mixed additions from
 * affine per-share tables, 1 batch-affine pairwise tree + Horner (28 % fewer multiplications, but slower on B200:
 * bound by global-memory latency at 2 warps per scheduler), 2 one scalar multiplication per share, 3 the G2 accumulation with one
 * item per thread, 4 the G2 accumulation on shared-memory cells at 4 blocks/SM (g2sm.cuh: faster below ~4 k items, slower at 2^14),
 * 5 = 0 without the "spill" layout, 6 = 0 with the spill layout forced.  (Spill: when one unit per item leaves unit slots of the
 * single wave idle — 2^14 items on 18 944 slots — the last share of every item moves to the spare units, a few items each, so the
 * longest unit gets shorter; chosen automatically by 0.)  Same outputs; 1-6 exist for measurement and tests. */
int tcb_set_msm_algo(tcb_ctx *ctx, int algo);
/* Commitment::evaluate: units (coefficient blocks) per evaluation point.  0 (default) = chosen from the batch size (small batches
 * are split so that they fill the GPU), 1 = never split, k > 1 = force k.  Same outputs. */
int tcb_set_eval_split(tcb_ctx *ctx, size_t units);
/* number of kernel launches issued through ctx since tcb_init (bench.py's gpu_launches) */
uint64_t tcb_launch_count(const tcb_ctx *ctx);

/* e(a,b) == e(c,d).  PublicKey::verify_g2 (src/lib.rs:108-110) with (a,b,c,d) =
 * (pk, hash, g1, sig); PublicKeyShare::verify_decryption_share (:182-186) with
 * (share, H(U,V), pk_i, W); Ciphertext::verify (:508-512) with (g1, W, U, H(U,V)).
 * c_g1 == NULL means the G1 generator for every item. */
int tcb_verify_g2_batch(tcb_ctx *, size_t n, const uint8_t *a_g1, const uint8_t *b_g2,
                        const uint8_t *c_g1, const uint8_t *d_g2, uint8_t *ok);
/* hash_g2 (src/lib.rs:691-694) */
int tcb_hash_g2_batch(tcb_ctx *, size_t n, const uint8_t *msgs, const uint64_t *off, uint8_t *out_g2);
/* hash_g1_g2 (src/lib.rs:697-707): H(msg-or-its-SHA3 || compressed(g1)); used by
 * Ciphertext::verify and verify_decryption_share together with tcb_verify_g2_batch */
int tcb_hash_g1_g2_batch(tcb_ctx *, size_t n, const uint8_t *g1, const uint8_t *msgs, const uint64_t *off,
                         uint8_t *out_g2);
/* PublicKey::verify / PublicKeyShare::verify (src/lib.rs:115-117,177-179) */
int tcb_verify_batch(tcb_ctx *, size_t n, const uint8_t *pk_g1, const uint8_t *sig_g2,
                     const uint8_t *msgs, const uint64_t *off, uint8_t *ok);
/* SecretKey::sign / SecretKeyShare::sign (src/lib.rs:379-381,447-449) */
int tcb_sign_batch(tcb_ctx *, size_t n, const uint8_t *sk, const uint8_t *msgs, const uint64_t *off,
                   uint8_t *out_g2);
/* SecretKey::sign_g2 (src/lib.rs:372-374) */
int tcb_sign_g2_batch(tcb_ctx *, size_t n, const uint8_t *sk, const uint8_t *h_g2, uint8_t *out_g2);
/* interpolate::<G2> behind PublicKeySet::combine_signatures (src/lib.rs:608-615,719-767).
 * x_fr: n*(t+1) scalars, already i+1 (into_fr_plus_1, :769-773); shares: n*(t+1) G2. */
int tcb_combine_g2_batch(tcb_ctx *, size_t n, size_t t, const uint8_t *x_fr, const uint8_t *shares_g2,
                         uint8_t *out_g2, uint8_t *status);
/* interpolate::<G1> (decryption shares) */
int tcb_combine_g1_batch(tcb_ctx *, size_t n, size_t t, const uint8_t *x_fr, const uint8_t *shares_g1,
                         uint8_t *out_g1, uint8_t *status);
/* SecretKeyShare::decrypt_share_no_verify (src/lib.rs:460-462): out = sk_i * U_i */
int tcb_decrypt_share_batch(tcb_ctx *, size_t n, const uint8_t *sk, const uint8_t *u_g1, uint8_t *out_g1);
/* PublicKeySet::decrypt (src/lib.rs:618-626): interpolate::<G1> then xor_with_hash (:710-715) */
int tcb_decrypt_batch(tcb_ctx *, size_t n, size_t t, const uint8_t *x_fr, const uint8_t *shares_g1,
                      const uint8_t *v, const uint64_t *v_off, uint8_t *out, uint8_t *status);
/* Commitment::evaluate (src/poly.rs:497-508) at n points of one commitment of degree deg */
int tcb_commitment_eval_batch(tcb_ctx *, size_t deg, const uint8_t *coeff_g1, size_t n,
                              const uint8_t *x_fr, uint8_t *out_g1);
/* SecretKey::public_key / Poly::commitment (src/lib.rs:367-369, src/poly.rs:372-377): g1 * c */
int tcb_g1_mul_gen_batch(tcb_ctx *, size_t n, const uint8_t *sk, uint8_t *out_g1);

/* SURVEY §8(f) row 3 — linear combinations out_i = sum_{k<m} s_{i,k} P_{i,k} (canonical Fr scalars, points of
 * order r).  BivarCommitment::row / evaluate (src/poly.rs:693-726) call it with the power products
 * x^j (resp. x^i y^j) as scalars; Commitment addition / Poly::commitment of sums reduce to it as well. */
int tcb_g1_lincomb_batch(tcb_ctx *, size_t n, size_t m, const uint8_t *scalars_fr, const uint8_t *pts_g1, uint8_t *out_g1);
int tcb_g2_lincomb_batch(tcb_ctx *, size_t n, size_t m, const uint8_t *scalars_fr, const uint8_t *pts_g2, uint8_t *out_g2);

/* SURVEY §8(f) row 4 — Fr-side Poly algebra on canonical little-endian Fr coefficients (32 B each, constant term first).
 * Poly::evaluate (src/poly.rs:358-369) of ONE polynomial of degree deg at n points, and the product of n pairs of polynomials of
 * degrees da and db (impl Mul for Poly, src/poly.rs:173-194; out holds da + db + 1 coefficients per item, no zero trimming).
 * rc -10 = a scalar >= r. */
int tcb_poly_eval_batch(tcb_ctx *, size_t deg, const uint8_t *coeff_fr, size_t n, const uint8_t *x_fr, uint8_t *out_fr);
int tcb_poly_mul_batch(tcb_ctx *, size_t n, size_t da, const uint8_t *a_fr, size_t db, const uint8_t *b_fr, uint8_t *out_fr);

/* SURVEY §8(f) row 2 — PublicKey::encrypt_with_rng (src/lib.rs:128-137) with the random scalars r
 * drawn by the caller (Fr::random stays in the Rust shim): u = g1*r, v = xor_with_hash(pk*r, msg),
 * w = hash_g1_g2(u, v)*r.  v_out has the layout of msgs (same offsets). */
int tcb_encrypt_batch(tcb_ctx *, size_t n, const uint8_t *pk_g1, const uint8_t *r_fr, const uint8_t *msgs,
                      const uint64_t *off, uint8_t *u_out_g1, uint8_t *v_out, uint8_t *w_out_g2);

/* SURVEY §8(f) row 1 — batched wire-format codecs.  Compressed encodings of SURVEY App. B (48 B / 96 B);
 * decompression is the CHECKED decode of PublicKey::from_bytes / Signature::from_bytes
 * (src/lib.rs:140-146,246-252) and serde `projective::deserialize` (src/serde_impl.rs:187-218):
 * flags, x < p, on the curve and in the r-order subgroup; status[i] = 0 ok | 3 invalid. */
int tcb_g1_compress_batch(tcb_ctx *, size_t n, const uint8_t *unc_g1, uint8_t *out48);
int tcb_g2_compress_batch(tcb_ctx *, size_t n, const uint8_t *unc_g2, uint8_t *out96);
int tcb_g1_decompress_batch(tcb_ctx *, size_t n, const uint8_t *in48, uint8_t *out_g1, uint8_t *status);
int tcb_g2_decompress_batch(tcb_ctx *, size_t n, const uint8_t *in96, uint8_t *out_g2, uint8_t *status);

/* Device-resident variants (inputs/outputs in HBM of ctx's first device; enqueue only). */
int tcb_verify_g2_batch_dev(tcb_ctx *, void *stream, size_t n, const uint8_t *a_g1, const uint8_t *b_g2,
                            const uint8_t *c_g1, const uint8_t *d_g2, uint8_t *ok);
int tcb_hash_g2_batch_dev(tcb_ctx *, void *stream, size_t n, const uint8_t *msgs, const uint64_t *off,
                          uint8_t *out_g2);
/* The two halves of tcb_verify_g2_batch_dev as separate launches (EXTERNAL pairing 0.16 Engine::miller_loop over the two pairs
 * (a,b), (-c,d), then final_exponentiation(..) == 1): f_out / f_in hold tcb_miller_value_bytes() per item (opaque, device-side
 * layout), enc_ok one byte per item (0 = some coordinate >= p).  bench.py times the two kernels through these. */
int tcb_miller_loop_batch_dev(tcb_ctx *, void *stream, size_t n, const uint8_t *a_g1, const uint8_t *b_g2,
                              const uint8_t *c_g1, const uint8_t *d_g2, void *f_out, uint8_t *enc_ok);
int tcb_final_exp_is_one_batch_dev(tcb_ctx *, void *stream, size_t n, const void *f_in, const uint8_t *enc_ok, uint8_t *ok);
size_t tcb_miller_value_bytes(void);
/* The first step of tcb_verify_batch_dev on its own: the message points the verifier pairs with pk, i.e. (by default, see
 * tcb_set_verify_hash) Q0_i = [3 (x^2 - 1)] hash_g2(msg_i), as uncompressed affine G2.  The matching first argument of the other
 * pairing is tcb_verifier_generator (the 96-byte encoding of [3 (x^2 - 1)] g1, or of g1 after tcb_set_verify_hash(ctx, 1)):
 *   tcb_verify_batch == tcb_verify_g2_batch(a = pk, b = these points, c = that generator for every item, d = sig).
 * bench.py times the hash kernels of the verify step through this entry point. */
int tcb_verifier_hash_g2_batch_dev(tcb_ctx *, void *stream, size_t n, const uint8_t *msgs, const uint64_t *off, uint8_t *out_g2);
int tcb_verifier_generator(const tcb_ctx *, uint8_t *out_g1);
int tcb_verify_batch_dev(tcb_ctx *, void *stream, size_t n, const uint8_t *pk_g1, const uint8_t *sig_g2,
                         const uint8_t *msgs, const uint64_t *off, uint8_t *ok);
int tcb_sign_batch_dev(tcb_ctx *, void *stream, size_t n, const uint8_t *sk, const uint8_t *msgs,
                       const uint64_t *off, const uint8_t *h_g2, uint8_t *out_g2);
int tcb_combine_g2_batch_dev(tcb_ctx *, void *stream, size_t n, size_t t, const uint8_t *x_fr,
                             const uint8_t *shares_g2, uint8_t *out_g2, uint8_t *status);
int tcb_combine_g1_batch_dev(tcb_ctx *, void *stream, size_t n, size_t t, const uint8_t *x_fr,
                             const uint8_t *shares_g1, uint8_t *out_g1, uint8_t *status);
int tcb_decrypt_batch_dev(tcb_ctx *, void *stream, size_t n, size_t t, const uint8_t *x_fr,
                          const uint8_t *shares_g1, const uint8_t *v, const uint64_t *v_off,
                          uint64_t v_total, uint8_t *out, uint8_t *status);
int tcb_g1_mul_batch_dev(tcb_ctx *, void *stream, size_t n, const uint8_t *sk, const uint8_t *pts_g1 /*NULL = generator*/,
                         uint8_t *out_g1);
int tcb_commitment_eval_batch_dev(tcb_ctx *, void *stream, size_t deg, const uint8_t *coeff_g1, size_t n,
                                  const uint8_t *x_fr, uint8_t *out_g1);

/* Self-test / measurement helpers (used by tests and bench.py, not part of the drop-in surface). */
/* Runs the PTX Montgomery multiply, dot2, add, sub against the portable CIOS on n random
 * pairs on the device; returns the number of mismatches (0 = pass) or < 0 on CUDA failure. */
int tcb_selftest_fp(tcb_ctx *, size_t n, uint64_t seed);
/* Runs the Miller loop of both engines on the caller's n items (host buffers, c_g1 may be NULL), then both final-exponentiation
 * kernels on the result, and returns the number of differing 32-bit words of the Miller values, the final-exponentiation values and
 * the flags (0 = bit-identical) or < 0 on CUDA failure. */
int tcb_selftest_miller(tcb_ctx *, size_t n, const uint8_t *a_g1, const uint8_t *b_g2, const uint8_t *c_g1, const uint8_t *d_g2);
/* Integer-MAC roofline probe: dependent-free IMAD.WIDE.U32 chains on every SM; writes the
 * achieved 32x32->64 multiply-accumulates per second. */
int tcb_probe_imad(tcb_ctx *, double *macs_per_sec);
/* Fp multiplication throughput (independent chains, all SMs): Montgomery muls per second. */
int tcb_probe_fpmul(tcb_ctx *, double *muls_per_sec);

#ifdef __cplusplus
}
#endif
#endif
