/* b381.h -- C ABI of the B200 BLS12-381 batch-verification engine (libb381.so).
 *
 * This is the drop-in boundary for the hot path of phoreproject/bls: the Go packages g1pubs /
 * g2pubs keep their exported API and forward the pairing / aggregation work here through cgo
 * (INTEGRATION.md shows the binding).  Each entry point names the reference function(s) it
 * replaces (file:line under the reference tree).
 *
 * Data layout (identical to the reference's in-memory structs, so Go can pass slices without
 * conversion):
 *   b381_fp     6 x u64 limbs, least-significant first, Montgomery form R = 2^384, canonical
 *               in [0, Q)                                   (FQ / FQRepr, fq.go:12-14, fqrepr.go:13-14)
 *   b381_fp2    c0 || c1                                    (FQ2, fq2.go:13-16)
 *   b381_fp12   c0.c0, c0.c1, c0.c2, c1.c0, c1.c1, c1.c2    (FQ12 -> FQ6 -> FQ2, flattened;
 *               the Go side flattens the two *FQ6 pointers of fq12.go:9-12 before the call)
 *   b381_g1_affine / b381_g2_affine   x, y, infinity flag + Go's struct padding (g1.go:10-14,
 *               g2.go:12-16): 104 / 200 bytes
 *   b381_g1_jac / b381_g2_jac         x, y, z Jacobian, infinity <=> z == 0 (g1.go:252-256,287-289)
 *   b381_scalar 4 x u64 limbs, least-significant first, canonical integer < r (FRRepr of
 *               FR.ToRepr(), fr.go:316-329)
 *
 * Conventions: every function returns B381_OK (0) or a negative error code and never aborts;
 * verification outcomes are written to out-parameters.  A ctx is bound to one CUDA device and
 * may be used by one host thread at a time.  Functions without the _dev suffix take HOST
 * pointers and do the host<->device copies themselves; _dev functions take DEVICE pointers,
 * enqueue on the ctx stream and return without synchronising (call b381_sync).
 * A pair with P or Q at infinity contributes the factor 1 to a Miller product (the reference
 * panics on such input: pairing.go:17-26, :53-56).
 */
#ifndef B381_H
#define B381_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { uint64_t l[6]; } b381_fp;
typedef struct { b381_fp c0, c1; } b381_fp2;
typedef struct { b381_fp2 c[6]; } b381_fp12;
typedef struct { b381_fp x, y; uint8_t infinity; uint8_t pad[7]; } b381_g1_affine;
typedef struct { b381_fp2 x, y; uint8_t infinity; uint8_t pad[7]; } b381_g2_affine;
typedef struct { b381_fp x, y, z; } b381_g1_jac;
typedef struct { b381_fp2 x, y, z; } b381_g2_jac;
typedef struct { uint64_t l[4]; } b381_scalar;

/* G2Prepared (g2.go:639-642): the 68 line-coefficient triples G2AffineToPrepared (g2.go:650-801) computes for a G2 point, in
 * the reference's order (63 doubling steps and 5 addition steps interleaved as the bits of blsX >> 1 dictate), then the
 * infinity flag + padding.  An infinite point has all-zero coefficients (the reference leaves the slice empty). */
typedef struct { b381_fp2 coeffs[68][3]; uint8_t infinity; uint8_t pad[7]; } b381_g2_prepared;

typedef struct b381_ctx b381_ctx;

enum {
    B381_OK = 0,
    B381_ERR_ARG = -1,      /* null pointer, bad size, bad group offsets */
    B381_ERR_CUDA = -2,     /* CUDA runtime error; see b381_last_error */
    B381_ERR_NOMEM = -3,    /* device or host allocation failed */
    B381_ERR_NO_DEVICE = -4 /* no usable CUDA device: there is no CPU fallback */
};

/* ---- context --------------------------------------------------------------------------------- */
int b381_init(int device, b381_ctx **out);
void b381_free(b381_ctx *ctx);
const char *b381_last_error(const b381_ctx *ctx);
/* run on an existing cudaStream_t (e.g. the caller's framework stream; NULL = CUDA's legacy default
 * stream); b381_use_own_stream returns to the non-blocking stream the ctx created in b381_init */
int b381_set_stream(b381_ctx *ctx, void *cuda_stream);
int b381_use_own_stream(b381_ctx *ctx);
int b381_sync(b381_ctx *ctx);
/* Which of the three schedules of the SAME arithmetic (bit-identical results) the pairing entry points launch:
 * AUTO picks by batch size (small batches: the warp-cooperative VM, low latency; large batches: one pairing per thread,
 * k_miller_loop / k_final_exp, the fastest at 2^16 on B200).  DUO (two lanes per pairing, k_duo_*: half the tower state per
 * lane, a third of the DRAM traffic, 9 % slower) and QUAD (four lanes, k_quad_*) are complete alternative schedules kept
 * for A/B measurements and as further implementations the parity tests run (profiles/r02_experiments.md).  The reference has one
 * schedule (pairing.go:16-129); this call has no reference counterpart.  Env B381_PATH=auto|thread|vm|quad|duo sets the
 * initial value. */
enum { B381_PATH_AUTO = -1, B381_PATH_THREAD = 0, B381_PATH_VM = 1, B381_PATH_QUAD = 2, B381_PATH_DUO = 3 };
int b381_set_kernel_path(b381_ctx *ctx, int path);
/* Bit length of the weights r_i of the random-linear-combination checks (b381_verify_rlc_dev, b381_verify_with_domain_rlc_batch,
 * b381_verify_rlc_partial_dev): default 255 (any canonical scalar); 64 or 128 is what a verifier needs (soundness error
 * 2^-bits) and shortens the weighted sums accordingly.  A weight of zero or of more than `bits` bits makes the check false. */
int b381_set_rlc_weight_bits(b381_ctx *ctx, int bits);
/* number of kernels this ctx has launched so far */
uint64_t b381_launch_count(const b381_ctx *ctx);

/* device buffers for callers without their own CUDA runtime (resident public-key tables) */
int b381_dev_alloc(b381_ctx *ctx, size_t bytes, void **dptr);
int b381_dev_free(b381_ctx *ctx, void *dptr);
int b381_h2d(b381_ctx *ctx, void *dptr, const void *host, size_t bytes);
int b381_d2h(b381_ctx *ctx, void *host, const void *dptr, size_t bytes);

/* ---- pairing --------------------------------------------------------------------------------- */
/* out[i] = bls.Pairing(p[i], q[i])                                        (pairing.go:132-136) */
int b381_pairing_batch(b381_ctx *ctx, const b381_g1_affine *p, const b381_g2_affine *q, size_t n,
                       b381_fp12 *out);
int b381_pairing_batch_dev(b381_ctx *ctx, const b381_g1_affine *d_p, const b381_g2_affine *d_q,
                           size_t n, b381_fp12 *d_out);
/* out[i] = bls.MillerLoop({p[i], G2AffineToPrepared(q[i])})        (pairing.go:16-75, g2.go:650-801) */
int b381_miller_loop_batch(b381_ctx *ctx, const b381_g1_affine *p, const b381_g2_affine *q, size_t n,
                           b381_fp12 *out);
/* A stream of batches from host memory: out[i] = Pairing(p[i], q[i]) for n pairs, `batch` pairs per launch, the copies of
 * neighbouring batches overlapped with the kernels (double buffers on the device, a second stream of the ctx).  Same values as
 * b381_pairing_batch; with page-locked host memory only the first copy in and the last copy out are not hidden.  This is the
 * host-buffer form of a caller that verifies batch after batch (pairing.go:131-138 once per pair in the reference). */
int b381_pairing_batch_stream(b381_ctx *ctx, const b381_g1_affine *p, const b381_g2_affine *q, size_t n, size_t batch,
                              b381_fp12 *out);
int b381_miller_loop_batch_dev(b381_ctx *ctx, const b381_g1_affine *d_p, const b381_g2_affine *d_q,
                               size_t n, b381_fp12 *d_out);
/* out[i] = bls.FinalExponentiation(in[i]); ok[i] = 0 where the reference returns nil (in[i] == 0),
 * out[i] is then 1                                                         (pairing.go:79-129) */
int b381_final_exp_batch(b381_ctx *ctx, const b381_fp12 *in, size_t n, b381_fp12 *out, uint8_t *ok);
int b381_final_exp_batch_dev(b381_ctx *ctx, const b381_fp12 *d_in, size_t n, b381_fp12 *d_out,
                             uint8_t *d_ok);
/* Prepared G2 points.  b381_g2_prepare_batch = n x G2AffineToPrepared (g2.go:650-801); b381_miller_loop_prepared_batch:
 * out[i] = MillerLoop([]MillerLoopItem{{p[i], prep[prep_idx ? prep_idx[i] : i]}}) (pairing.go:4-7,16-75): the same Fq12 as
 * b381_miller_loop_batch on the unprepared point, 1 760 Fq multiplications cheaper per pair.  A MillerLoop over several items is
 * the product of the one-item values (the shared accumulator of pairing.go:40-69 computes exactly that product).  prep_idx
 * lets many pairs share one prepared point (the hash of a message many committees signed; the generator in g2pubs); nprep is
 * the number of prepared points behind `prep`. */
int b381_g2_prepare_batch(b381_ctx *ctx, const b381_g2_affine *q, size_t n, b381_g2_prepared *prep);
int b381_g2_prepare_batch_dev(b381_ctx *ctx, const b381_g2_affine *d_q, size_t n, b381_g2_prepared *d_prep);
int b381_miller_loop_prepared_batch(b381_ctx *ctx, const b381_g1_affine *p, const b381_g2_prepared *prep, size_t nprep,
                                    const uint32_t *prep_idx, size_t n, b381_fp12 *out);
int b381_miller_loop_prepared_batch_dev(b381_ctx *ctx, const b381_g1_affine *d_p, const b381_g2_prepared *d_prep,
                                        const uint32_t *d_prep_idx, size_t n, b381_fp12 *d_out);
/* Generalised bls.CompareTwoPairings (pairing.go:140-147): for every group g,
 *   ok[g] = FinalExponentiation(prod_{i in [group_off[g], group_off[g+1])} MillerLoop(p[i], q[i])) == 1.
 * group_off has ngroups+1 non-decreasing entries, group_off[0] == 0, group_off[ngroups] == npairs.
 * CompareTwoPairings(P1,Q1,P2,Q2) is the group {(P1,Q1), (-P2,Q2)}; g1pubs.Verify
 * (g1pubs/bls.go:165-168) is the group {(G1One, sig), (-pub, H(m))}; VerifyAggregate
 * (g1pubs/bls.go:252-282) is one group of n+1 pairs.                                           */
int b381_pairing_product_is_one(b381_ctx *ctx, const b381_g1_affine *p, const b381_g2_affine *q,
                                size_t npairs, const uint32_t *group_off, size_t ngroups, uint8_t *ok);
int b381_pairing_product_is_one_dev(b381_ctx *ctx, const b381_g1_affine *d_p, const b381_g2_affine *d_q,
                                    size_t npairs, const uint32_t *d_group_off, size_t ngroups,
                                    uint8_t *d_ok);

/* ---- aggregation ----------------------------------------------------------------------------- */
/* out = sum of n affine G1 points == AggregatePublicKeys (g1pubs/bls.go:192-198) /
 * g2pubs.AggregateSignatures (g2pubs/bls.go:165-177).  The result is returned normalised
 * (z = 1, or the canonical zero (0,1,0) of g1.go:269): it Equal()s the reference's fold, whose
 * Jacobian coordinates depend on the order of additions.                                       */
int b381_g1_sum(b381_ctx *ctx, const b381_g1_affine *p, size_t n, b381_g1_jac *out);
int b381_g1_sum_dev(b381_ctx *ctx, const b381_g1_affine *d_p, size_t n, b381_g1_jac *d_out);
/* same over G2 == g1pubs.AggregateSignatures (g1pubs/bls.go:177-183) / g2pubs.AggregatePublicKeys */
int b381_g2_sum(b381_ctx *ctx, const b381_g2_affine *p, size_t n, b381_g2_jac *out);
int b381_g2_sum_dev(b381_ctx *ctx, const b381_g2_affine *d_p, size_t n, b381_g2_jac *d_out);
/* out = sum_i k[i] * p[i]  (Pippenger bucket method).  The reference has no MSM; this equals the
 * fold of G1Affine.MulFR (g1.go:80-90) results with G1Projective.Add (g1.go:400-482), normalised
 * as above.                                                                                    */
int b381_g1_msm(b381_ctx *ctx, const b381_g1_affine *p, const b381_scalar *k, size_t n, b381_g1_jac *out);
int b381_g1_msm_dev(b381_ctx *ctx, const b381_g1_affine *d_p, const b381_scalar *d_k, size_t n,
                    b381_g1_jac *d_out);
/* The same multi-scalar multiplication over G2 (weighted signature aggregation, g1pubs/bls.go:177-183 with weights;
 * the reference folds G2Affine.MulFR results, g2.go:92-102): out = sum_i k[i] * p[i], normalised (z = 1). */
int b381_g2_msm(b381_ctx *ctx, const b381_g2_affine *p, const b381_scalar *k, size_t n, b381_g2_jac *out);
int b381_g2_msm_dev(b381_ctx *ctx, const b381_g2_affine *d_p, const b381_scalar *d_k, size_t n, b381_g2_jac *d_out);
/* Bucket-sharded MSM for multi-GPU (one process per GPU): this rank accumulates only the windows
 * w with w % nranks == rank and writes its partial sum (Jacobian, already weighted by 2^(c*w)) to
 * *d_partial.  The caller all-gathers the nranks partials and folds them with b381_g1_fold_dev.  */
int b381_g1_msm_shard_dev(b381_ctx *ctx, const b381_g1_affine *d_p, const b381_scalar *d_k, size_t n,
                          int rank, int nranks, b381_g1_jac *d_partial);
/* The same call with the device time of its phases (measurement; synchronises the stream): phase_ms[0..4] = sort of the point
 * indices by bucket (histogram, scan, scatter), chunk partial sums, in-bucket chunk tree, bucket reduction (running-sum
 * segments + window sums), window combine.  phase_ms is HOST memory, 5 floats. */
int b381_g1_msm_shard_phases_dev(b381_ctx *ctx, const b381_g1_affine *d_p, const b381_scalar *d_k, size_t n,
                                 int rank, int nranks, b381_g1_jac *d_partial, float *phase_ms);
/* out = normalised sum of n Jacobian points (G1Projective.Add fold, g1.go:400-482) */
int b381_g1_fold_dev(b381_ctx *ctx, const b381_g1_jac *d_parts, size_t n, b381_g1_jac *d_out);

/* ---- signature-scheme glue ------------------------------------------------------------------- */
/* Batch of g1pubs VerifyAggregateCommon (g1pubs/bls.go:287-290) over a resident key registry:
 * attestation a uses public keys registry[key_idx[key_off[a] .. key_off[a+1])], the aggregate
 * signature sig[a] and the pre-hashed message point msg_hash[msg_idx[a]] (= HashG2(m) /
 * HashG2WithDomain(m, d), computed by the Go host).  ok[a] = e(G1One, sig[a]) == e(sum pk, H(m)). */
int b381_verify_aggregate_common_batch_dev(b381_ctx *ctx, const b381_g1_affine *d_registry,
                                           const uint32_t *d_key_idx, const uint32_t *d_key_off,
                                           const b381_g2_affine *d_sig, const b381_g2_affine *d_msg_hash,
                                           const uint32_t *d_msg_idx, size_t nattest, size_t nkeys, size_t nmsg, uint8_t *d_ok);
/* The same batch as ONE boolean (batch verification by a random linear combination, grouped by message):
 * *d_ok = [ e(-G1One, sum_a r_a sig[a]) * prod_m e(sum_{a: msg_idx[a] = m} r_a pk_a, msg_hash[m]) == 1 ], which holds iff every
 * attestation verifies, except with probability ~2^-bits for weights r_a drawn independently by the VERIFIER after the batch
 * is fixed (b381_set_rlc_weight_bits; a zero or over-wide weight, an index outside the tables, an empty committee, an
 * infinite signature or aggregate key make it false).  nmsg + 1 Miller loops and one FinalExponentiation per batch instead of
 * two and one per attestation (pairing.go:16-129); on a false result the caller locates the offender with the per-attestation
 * call above.  nmsg <= 2^22.  No reference counterpart: the reference verifies one aggregate at a time (g1pubs/bls.go:287-297). */
int b381_verify_aggregate_common_rlc_dev(b381_ctx *ctx, const b381_g1_affine *d_registry,
                                         const uint32_t *d_key_idx, const uint32_t *d_key_off,
                                         const b381_g2_affine *d_sig, const b381_g2_affine *d_msg_hash,
                                         const uint32_t *d_msg_idx, const b381_scalar *d_r, size_t nattest, size_t nkeys,
                                         size_t nmsg, uint8_t *d_ok);
/* The same batch as the attestations arrive on the wire: compressed aggregate signatures (n x 96 bytes) and 32-byte
 * message hashes (nmsg DISTINCT messages, msg_idx[a] selects one) with one 8-byte domain.  On the device:
 * DeserializeSignature with the subgroup check (g1pubs/bls.go:38-45), HashG2WithDomain per distinct message
 * (g2.go:1041-1085), key aggregation, CompareTwoPairings.  ok[a] = 0 also when the signature fails to deserialise or is
 * the point at infinity.  The registry holds validated keys (b381_g1_decompress_batch with the subgroup check). */
int b381_verify_aggregate_common_with_domain_batch_dev(b381_ctx *ctx, const b381_g1_affine *d_registry,
                                                       const uint32_t *d_key_idx, const uint32_t *d_key_off,
                                                       const uint8_t *d_sig96, const uint8_t *d_msg32, size_t nmsg,
                                                       const uint8_t *d_domain8, const uint32_t *d_msg_idx, size_t nattest,
                                                       size_t nkeys, uint8_t *d_ok);

/* ---- wire formats and scalar multiplication (the callers on either side of the verification path) ------------ */
/* DecompressG1 (check_subgroup != 0, g1.go:185-195) / DecompressG1Unchecked (g1.go:199-227) over n x 48 bytes;
 * DecompressG2 (g2.go:219-229) / DecompressG2Unchecked (g2.go:232-265) over n x 96 bytes (x.c1 first).
 * status[i]: 0 ok; 1 "unexpected compression mode"; 2 "unexpected information in compressed infinity";
 * 3 "point not on curve"; 4 "not in correct subgroup" (the reference's error strings).  out[i] is the point for
 * status 0 and 4 and the canonical zero (0, 1, infinity) otherwise.  A coordinate >= Q decodes as 0 exactly as
 * FQReprToFQ does (fq.go:49-56).  This is what g1pubs.DeserializePublicKey / DeserializeSignature and their g2pubs
 * mirrors run per key (g1pubs/bls.go:38-58,91-111). */
enum { B381_POINT_OK = 0, B381_POINT_ERR_MODE = 1, B381_POINT_ERR_INFINITY = 2, B381_POINT_ERR_NOT_ON_CURVE = 3,
       B381_POINT_ERR_SUBGROUP = 4 };
int b381_g1_decompress_batch(b381_ctx *ctx, const uint8_t *in, size_t n, int check_subgroup, b381_g1_affine *out,
                             uint8_t *status);
int b381_g1_decompress_batch_dev(b381_ctx *ctx, const uint8_t *d_in, size_t n, int check_subgroup,
                                 b381_g1_affine *d_out, uint8_t *d_status);
int b381_g2_decompress_batch(b381_ctx *ctx, const uint8_t *in, size_t n, int check_subgroup, b381_g2_affine *out,
                             uint8_t *status);
int b381_g2_decompress_batch_dev(b381_ctx *ctx, const uint8_t *d_in, size_t n, int check_subgroup,
                                 b381_g2_affine *d_out, uint8_t *d_status);
/* CompressG1 (g1.go:230-249) / CompressG2 (g2.go:268-289): n affine points -> n x 48 / n x 96 bytes
 * (Serialize of keys and signatures, g1pubs/bls.go:18-24,67-70). */
int b381_g1_compress_batch(b381_ctx *ctx, const b381_g1_affine *in, size_t n, uint8_t *out);
int b381_g1_compress_batch_dev(b381_ctx *ctx, const b381_g1_affine *d_in, size_t n, uint8_t *d_out);
int b381_g2_compress_batch(b381_ctx *ctx, const b381_g2_affine *in, size_t n, uint8_t *out);
int b381_g2_compress_batch_dev(b381_ctx *ctx, const b381_g2_affine *d_in, size_t n, uint8_t *d_out);
/* out[i] = k[i * k_stride] * p[i * p_stride] as an affine point: G1Affine.MulFR / G2Affine.MulFR followed by
 * ToAffine (g1.go:80-90,322-340; g2.go:92-102,365-386).  Strides are 0 or 1: p_stride 0 multiplies ONE base by n
 * scalars (PrivToPub = sk * G1One, g1pubs/bls.go:144-146), k_stride 0 multiplies n points by ONE scalar (Sign of n
 * hashed messages with one key, g1pubs/bls.go:132-141).  Scalars are canonical integers < r. */
int b381_g1_mul_batch(b381_ctx *ctx, const b381_g1_affine *p, size_t p_stride, const b381_scalar *k, size_t k_stride,
                      size_t n, b381_g1_affine *out);
int b381_g1_mul_batch_dev(b381_ctx *ctx, const b381_g1_affine *d_p, size_t p_stride, const b381_scalar *d_k,
                          size_t k_stride, size_t n, b381_g1_affine *d_out);
int b381_g2_mul_batch(b381_ctx *ctx, const b381_g2_affine *p, size_t p_stride, const b381_scalar *k, size_t k_stride,
                      size_t n, b381_g2_affine *out);
int b381_g2_mul_batch_dev(b381_ctx *ctx, const b381_g2_affine *d_p, size_t p_stride, const b381_scalar *d_k,
                          size_t k_stride, size_t n, b381_g2_affine *d_out);
/* The same products for points that are KNOWN to lie in G1 / G2 (the generator, a hash to the curve, a key or signature that
 * passed the subgroup check): the 255-bit ladder is replaced by the endomorphism ladders (G1: k = k0 + k1 x^2 over
 * (x, y) and (beta x, -y); G2: base-|x| digits over psi), 1.5x / 1.9x faster, same affine result bit for bit.  For a
 * point outside the r-torsion the result is NOT k * p: use b381_g{1,2}_mul_batch there.  This is what PrivToPub and
 * Sign (g1pubs/bls.go:132-146, g2pubs/bls.go:126-140) need. */
int b381_g1_mul_subgroup_batch(b381_ctx *ctx, const b381_g1_affine *p, size_t p_stride, const b381_scalar *k,
                               size_t k_stride, size_t n, b381_g1_affine *out);
int b381_g1_mul_subgroup_batch_dev(b381_ctx *ctx, const b381_g1_affine *d_p, size_t p_stride, const b381_scalar *d_k,
                                   size_t k_stride, size_t n, b381_g1_affine *d_out);
int b381_g2_mul_subgroup_batch(b381_ctx *ctx, const b381_g2_affine *p, size_t p_stride, const b381_scalar *k,
                               size_t k_stride, size_t n, b381_g2_affine *out);
int b381_g2_mul_subgroup_batch_dev(b381_ctx *ctx, const b381_g2_affine *d_p, size_t p_stride, const b381_scalar *d_k,
                                   size_t k_stride, size_t n, b381_g2_affine *d_out);

/* HashG2WithDomain (g2.go:1041-1085) over a batch: out[i] = the affine form of HashG2WithDomain(msg32[i],
 * domain8[i * domain_stride]) (the reference returns the same point unnormalised).  domain_stride is 0 (one 8-byte
 * domain for the batch) or 1.  This is the message point of VerifyWithDomain / SignWithDomain /
 * VerifyAggregate[Common]WithDomain (g1pubs/bls.go:138-141,171-174,294-311). */
int b381_hash_g2_with_domain_batch(b381_ctx *ctx, const uint8_t *msg32, const uint8_t *domain8, size_t domain_stride,
                                   size_t n, b381_g2_affine *out);
int b381_hash_g2_with_domain_batch_dev(b381_ctx *ctx, const uint8_t *d_msg32, const uint8_t *d_domain8,
                                       size_t domain_stride, size_t n, b381_g2_affine *d_out);

/* g1pubs.VerifyWithDomain (g1pubs/bls.go:171-174) for n independent (public key, message hash, signature) triples
 * given in WIRE format: pub48 = n x 48 bytes (PublicKey.Serialize), sig96 = n x 96 bytes (Signature.Serialize),
 * msg32 = n x 32 bytes, domain8 = 8 bytes (domain_stride 0) or n x 8 bytes (domain_stride 1).  Per item the device
 * runs DeserializePublicKey + DeserializeSignature (decompression and subgroup checks, g1pubs/bls.go:38-58,91-111),
 * HashG2WithDomain, and CompareTwoPairings(G1One, sig, pub, H) (pairing.go:140-147).  ok[i] = 1 iff everything
 * succeeded and the pairing equation holds; a key or signature that fails to deserialise, or is the point at
 * infinity (where the reference panics), yields 0. */
int b381_verify_with_domain_batch(b381_ctx *ctx, const uint8_t *pub48, const uint8_t *msg32, const uint8_t *domain8,
                                  size_t domain_stride, const uint8_t *sig96, size_t n, uint8_t *ok);
int b381_verify_with_domain_batch_dev(b381_ctx *ctx, const uint8_t *d_pub48, const uint8_t *d_msg32,
                                      const uint8_t *d_domain8, size_t domain_stride, const uint8_t *d_sig96, size_t n,
                                      uint8_t *d_ok);

/* HashG1 (hash.go:320-331) / HashG2 (hash.go:404-411) over a batch of variable-length messages packed back to back:
 * message i is msgs[msg_off[i] .. msg_off[i+1]), msg_off has n + 1 entries starting at 0.  hash_to_field (hp / hp2,
 * hash.go:41-113), the simplified SWU maps (g1.go:628-714, g2.go:933-1031), iso11 / iso3 and ClearH / clearH2
 * (hash.go:185-203,282-309,341-389) all run on the device.  These are the message points of g2pubs.Verify / Sign
 * (HashG1) and g1pubs.Verify / Sign / VerifyAggregate (HashG2). */
int b381_hash_g1_batch(b381_ctx *ctx, const uint8_t *msgs, const uint64_t *msg_off, size_t n, b381_g1_affine *out);
int b381_hash_g1_batch_dev(b381_ctx *ctx, const uint8_t *d_msgs, const uint64_t *d_msg_off, size_t n, b381_g1_affine *d_out);
int b381_hash_g2_batch(b381_ctx *ctx, const uint8_t *msgs, const uint64_t *msg_off, size_t n, b381_g2_affine *out);
int b381_hash_g2_batch_dev(b381_ctx *ctx, const uint8_t *d_msgs, const uint64_t *d_msg_off, size_t n, b381_g2_affine *d_out);
/* g1pubs.Verify (g1pubs/bls.go:165-168) and g2pubs.Verify (g2pubs/bls.go:159-162) for n independent wire-format
 * triples: like b381_verify_with_domain_batch, with HashG2(msg) / HashG1(msg) as the message point.  g1pubs: 48-byte
 * keys, 96-byte signatures; g2pubs: 96-byte keys, 48-byte signatures. */
int b381_g1pubs_verify_batch(b381_ctx *ctx, const uint8_t *pub48, const uint8_t *msgs, const uint64_t *msg_off,
                             const uint8_t *sig96, size_t n, uint8_t *ok);
int b381_g1pubs_verify_batch_dev(b381_ctx *ctx, const uint8_t *d_pub48, const uint8_t *d_msgs, const uint64_t *d_msg_off,
                                 const uint8_t *d_sig96, size_t n, uint8_t *d_ok);
int b381_g2pubs_verify_batch(b381_ctx *ctx, const uint8_t *pub96, const uint8_t *msgs, const uint64_t *msg_off,
                             const uint8_t *sig48, size_t n, uint8_t *ok);
int b381_g2pubs_verify_batch_dev(b381_ctx *ctx, const uint8_t *d_pub96, const uint8_t *d_msgs, const uint64_t *d_msg_off,
                                 const uint8_t *d_sig48, size_t n, uint8_t *d_ok);

/* Random-linear-combination batch verification: ONE boolean for n (public key, message point, signature) triples,
 *   *ok = [ prod_i e(r_i pk_i, H_i) * e(-G1One, sum_i r_i sig_i) == 1 ],
 * i.e. every g1pubs.Verify* equation e(G1One, sig_i) == e(pk_i, H_i) (g1pubs/bls.go:165-174, pairing.go:140-147) holds,
 * up to a false-accept probability of about 2^-bits(r_i) for independent uniformly random weights r_i chosen by the
 * caller AFTER the signatures are fixed (64-bit weights recommended; r is an array of canonical scalars, upper limbs
 * zero).  n + 1 Miller loops and a single final exponentiation instead of n checks; the reference has no batch
 * verification, this is the addition SURVEY.md 8d names for the attestation workload.  *ok = 0 when any key or
 * signature is the point at infinity or fails to deserialise.  The _rlc_batch forms take wire bytes like
 * b381_verify_with_domain_batch. */
int b381_verify_rlc_dev(b381_ctx *ctx, const b381_g1_affine *d_pub, const b381_g2_affine *d_msg_point,
                        const b381_g2_affine *d_sig, const b381_scalar *d_r, size_t n, uint8_t *d_ok);
int b381_verify_with_domain_rlc_batch(b381_ctx *ctx, const uint8_t *pub48, const uint8_t *msg32, const uint8_t *domain8,
                                      size_t domain_stride, const uint8_t *sig96, const b381_scalar *r, size_t n,
                                      uint8_t *ok);
int b381_verify_with_domain_rlc_batch_dev(b381_ctx *ctx, const uint8_t *d_pub48, const uint8_t *d_msg32,
                                          const uint8_t *d_domain8, size_t domain_stride, const uint8_t *d_sig96,
                                          const b381_scalar *d_r, size_t n, uint8_t *d_ok);

/* Multi-GPU form of the random-linear-combination check: every rank forms the factor of ITS triples,
 *   partial = prod_i MillerLoop(r_i pk_i, H_i) * MillerLoop(-G1One, sum_i r_i sig_i)      (no final exponentiation),
 * the ranks all-gather the 576-byte partials, and each finishes with b381_fp12_product_final_exp_is_one_dev.
 * *d_valid = 0 when one of the rank's keys or signatures is the point at infinity.  b381_miller_product_dev is the
 * same building block for arbitrary pairs: prod_i MillerLoop(p[i], q[i]) (pairing.go:16-75 with len(items) = n). */
int b381_verify_rlc_partial_dev(b381_ctx *ctx, const b381_g1_affine *d_pub, const b381_g2_affine *d_msg_point,
                                const b381_g2_affine *d_sig, const b381_scalar *d_r, size_t n, b381_fp12 *d_partial,
                                uint8_t *d_valid);
int b381_miller_product_dev(b381_ctx *ctx, const b381_g1_affine *d_p, const b381_g2_affine *d_q, size_t npairs,
                            b381_fp12 *d_out);
int b381_fp12_product_final_exp_is_one_dev(b381_ctx *ctx, const b381_fp12 *d_parts, size_t n, uint8_t *d_ok);

/* ---- measurement -------------------------------------------------------------------------------- */
/* Integer-pipe roofline probe: launches blocks x threads threads that each issue iters * 8
 * independent IMAD.WIDE.U32 (the 32x32->64 multiply-accumulate an Fq multiplication is made of).
 * d_out needs blocks*threads u32.  Time it with events on the ctx stream.                        */
int b381_imad_probe_dev(b381_ctx *ctx, uint32_t *d_out, int blocks, int threads, int iters);
/* The same roofline through the engine's own Fq multiplier: every thread runs a dependent chain of
 * iters Montgomery multiplications held in registers (300 carry-chained IMAD.WIDE.U32 each); threads <= 256.
 * The faster of the two probes is the denominator bench.py reports against.                         */
int b381_fpmul_probe_dev(b381_ctx *ctx, uint32_t *d_out, int blocks, int threads, int iters);

/* Test hook of the warp-cooperative VM behind the pairing entry points (bls_b200/csrc/vm.cuh): runs an encoded
 * program (host image: nsteps * lanes * 64 bytes; constants: nconsts * 96 bytes, Montgomery form) over n units
 * whose per-unit elements live in the device arrays d_seg[0..3] with the given strides, then synchronises.
 * tests/test_gpu_vm.py drives random programs through it against the big-integer emulator.              */
int b381_vm_exec_dev(b381_ctx *ctx, const void *code, int lanes, int nsteps, int nslots, const void *consts, int nconsts,
                     void *const d_seg[4], const size_t stride[4], size_t n);

/* test hook: ONE operation of the device build of the field / tower / group-law routines over n operand pairs in host
 * memory, so the reference's known-answer tests (fq_test.go:189-207, fq2_test.go:71-246, g1_test.go:62-104) and edge operands
 * reach the CUDA code directly (tests/test_gpu_ops.py).  family: 0 Fq (48 B) 1 Fq2 (96 B) 2 Fq6 (288 B) 3 Fq12 one element per
 * thread (576 B) 4 Fq12 on four lanes (576 B) 5 G1 / 6 G2 group law (affine PODs in, normalised Jacobian out); op codes in
 * bls_b200/csrc/testops.cuh.  out2 (n x 96 B) and ok (n bytes) may be NULL. */
int b381_test_op(b381_ctx *ctx, int family, int op, uint64_t arg, const void *a, const void *b, void *out, void *out2, uint8_t *ok,
                 size_t n);

#ifdef __cplusplus
}
#endif
#endif /* B381_H */
