/*
 * ringo_cuda.h -- C ABI of libringo_b200.so, the sm_100a implementation of
 * Ringo-SNARK's polynomial-arithmetic hot path.
 *
 * The reference (sp301415/ringo-snark) is a pure-Go library with no FFI layer; the
 * seams this library plugs into are Go interfaces / method sets.  Every entry point
 * below names the reference symbol it replaces (file:line relative to the reference
 * root).  A Go maintainer binds them with cgo (INTEGRATION.md shows the shim).
 *
 * Conventions
 *   - All integers little-endian.  A field element is L64 uint64 limbs in MONTGOMERY
 *     form (R = 2^(64 L64)), fully reduced -- byte-identical to the reference's
 *     `type Uint [L]uint64` (jindo/internal/zp/element.go:24-37), so the Go shim copies
 *     element memory verbatim.
 *   - A "batch" is n_polys x rank elements, array-of-structs, poly-major, device
 *     resident.  Host<->device copies happen only in rs_batch_upload/download.
 *   - Every call returns RS_OK (0) or a negative code; rs_last_error() gives the text.
 *     Shape/state misuse that makes the reference panic (math/bigpoly/poly.go:106-121,
 *     base_op.go:135,146,183,201) returns RS_ERR_SHAPE / RS_ERR_STATE; the shim re-panics.
 *   - Calls are asynchronous on the library's current CUDA stream (rs_set_stream);
 *     rs_batch_download and rs_sync synchronise.  One context per goroutine, like the
 *     reference's SafeCopy convention (jindo/prover.go:327-338).
 *   - There is no CPU fallback: without a CUDA device every compute call fails with
 *     RS_ERR_CUDA.
 */
#ifndef RINGO_CUDA_H
#define RINGO_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RS_OK 0
#define RS_ERR_ARG (-1)         /* null pointer / bad argument */
#define RS_ERR_UNSUPPORTED (-2) /* limb count or rank outside the compiled kernels */
#define RS_ERR_SHAPE (-3)       /* "inconsistent input(s)" panics of the reference */
#define RS_ERR_STATE (-4)       /* IsNTT-domain misuse panics of the reference */
#define RS_ERR_CUDA (-5)        /* CUDA runtime failure (message has the cudaError) */
#define RS_ERR_NTT (-6)         /* "NTT not supported" (ntt.go:35-37): 2N does not divide q-1 */

typedef struct rs_field rs_field;
typedef struct rs_ntt rs_ntt;
typedef struct rs_batch rs_batch;

/* ---- runtime -------------------------------------------------------------------------- */
int rs_init(int device);                    /* cudaSetDevice + capability check (needs sm_100) */
int rs_set_stream(void *cuda_stream);       /* stream used by subsequent calls (NULL = default) */
int rs_sync(void);                          /* cudaStreamSynchronize(current stream) */
/* stream plumbing for callers that run independent pieces side by side (the compiled Buckler prover issues the Commits of a round
 * on a few side streams): create / destroy a non-blocking stream, read the calling thread's current one, and make `waiter`
 * wait -- on the device, no host wait -- for everything enqueued on `signaller` so far */
void *rs_get_stream(void);
int rs_stream_create(void **stream_out);
void rs_stream_destroy(void *stream);
int rs_stream_wait(void *waiter, void *signaller);
const char *rs_last_error(void);
const char *rs_version(void);
/* number of kernels this library has launched since load (bench.py "gpu_launches") */
unsigned long long rs_launch_count(void);

/* roofline denominator: measured sustained multiply-add rate of the integer pipe, in G/s
 * (wide != 0: IMAD.WIDE.U32 32x32+64, the instruction the Montgomery product is built from) */
int rs_bench_imad_peak(int wide, double *gops);

/* pinned host staging the Go shim gathers []E into (cgo cannot pass []*Uint) */
void *rs_pinned_alloc(size_t bytes);
void rs_pinned_free(void *p);

/* ---- field: bignum.Uint[E] (math/bignum/bignum.go:8-31) -------------------------------- */
/* Builds qInvNeg, R, R^2 exactly like the generated zp packages (element.go:72, :784).
 * modulus_le: L64 limbs of an odd prime q.  Supported L64: 2, 4, 7, 14 (every shipped zp). */
int rs_field_create(const uint64_t *modulus_le, int L64, rs_field **out);
void rs_field_destroy(rs_field *f);
int rs_field_limbs(const rs_field *f);                        /* Limb() element.go:1776-1778 */
/* host-side scalar helpers (used by the shim for challenges / constants; value-level
 * identical to SetUint64 / Mul / toMont / fromMont of element.go:94-99, :470-479, :604-653) */
int rs_field_to_mont(const rs_field *f, uint64_t *out, const uint64_t *canonical_le);
int rs_field_from_mont(const rs_field *f, uint64_t *out, const uint64_t *mont_le);
/* Mul / Add / Sub (element.go:470-479, :400-445) on single Montgomery elements; out may alias an input */
int rs_field_mul(const rs_field *f, uint64_t *out, const uint64_t *a, const uint64_t *b);
int rs_field_add(const rs_field *f, uint64_t *out, const uint64_t *a, const uint64_t *b);
int rs_field_sub(const rs_field *f, uint64_t *out, const uint64_t *a, const uint64_t *b);
/* SetBytes (element.go:891-915): len <= 128 bytes read as a big-endian integer, reduced mod q, Montgomery form out --
 * how buckler turns its 32-byte SHA-256 challenges into field elements (prover.go:248, :273-274, :308, :340);
 * Marshal (element.go:873-883): the canonical value as 8 L64 big-endian bytes (transcript bindings, prover.go:220, :236). */
int rs_field_set_bytes(const rs_field *f, const uint8_t *be_bytes, size_t len, uint64_t *out_mont);
int rs_field_marshal(const rs_field *f, const uint64_t *mont, uint8_t *out_be);

/* ---- device batches: []*bigpoly.Poly (math/bigpoly/poly.go:11-28) ----------------------- */
int rs_batch_alloc(rs_field *f, size_t n_polys, size_t rank, rs_batch **out);
/* wrap caller-owned device memory (e.g. a torch tensor) of n_polys*rank*L64 uint64 */
int rs_batch_wrap(rs_field *f, void *device_ptr, size_t n_polys, size_t rank, rs_batch **out);
void rs_batch_free(rs_batch *b);
void *rs_batch_device_ptr(rs_batch *b);
size_t rs_batch_polys(const rs_batch *b);
size_t rs_batch_rank(const rs_batch *b);
int rs_batch_upload(rs_batch *b, size_t poly0, size_t n_polys, const uint64_t *host_aos);
int rs_batch_download(const rs_batch *b, size_t poly0, size_t n_polys, uint64_t *host_aos);
/* same copy without the trailing stream synchronisation (host_aos must be pinned; the caller
 * synchronises with rs_sync) -- lets uploads, transforms and downloads of different chunks overlap
 * on different streams */
int rs_batch_download_async(const rs_batch *b, size_t poly0, size_t n_polys, uint64_t *host_aos);
/* element-range copies (EncodeTo zero-padding, Coeffs[:N+1] slices, remLo/remHi moves;
 * buckler/encoder.go:32-38, buckler/prover.go:152, :443-456) */
int rs_batch_copy(rs_batch *dst, size_t dst_poly, size_t dst_off,
                  const rs_batch *src, size_t src_poly, size_t src_off, size_t n_elems);
int rs_batch_zero(rs_batch *dst, size_t poly, size_t off, size_t n_elems);   /* Poly.Clear poly.go:79-84 */

/* ---- NTT: bigpoly.transformer (math/bigpoly/ntt.go:11-14) ------------------------------- */
/* NewCyclicTransformer (ntt.go:26-95) when negacyclic == 0, NewCyclotomicTransformer
 * (ntt.go:153-203) otherwise: same generator search, same bit-reversed twiddle tables. */
int rs_ntt_create(rs_field *f, int logN, int negacyclic, rs_ntt **out);
void rs_ntt_destroy(rs_ntt *t);
int rs_ntt_rank_log2(const rs_ntt *t);
/* copies the (Montgomery) twiddle tables back to the host: tw, twinv are N elements each */
int rs_ntt_tables(const rs_ntt *t, uint64_t *tw, uint64_t *twinv, uint64_t *rankinv);
/* FwdNTTTo / InvNTTTo (ntt.go:98-136, :206-244) on n_polys polynomials of rank N.
 * dst may alias src (in place).  Batches must have rank == N (else RS_ERR_SHAPE). */
int rs_ntt_forward(rs_ntt *t, rs_batch *dst, size_t dst_poly, const rs_batch *src, size_t src_poly, size_t n_polys);
int rs_ntt_inverse(rs_ntt *t, rs_batch *dst, size_t dst_poly, const rs_batch *src, size_t src_poly, size_t n_polys);
/* Measurement hook for bench.py: runs the transform `reps` times in place on batch b and
 * returns, per kernel launch ("pass") of one transform, the mean device time in ms measured
 * with CUDA events on the current stream.  *n_passes in: capacity of pass_ms, out: passes. */
int rs_ntt_profile(rs_ntt *t, rs_batch *b, size_t n_polys, int inverse, int reps, float *pass_ms, int *n_passes);
/* Sub-range transforms used by buckler.Encoder.EncodeTo (buckler/encoder.go:32-38): the
 * rank-N transformer acts on the first N coefficients of each poly of a larger-rank batch;
 * coefficients [N, dst rank) of dst are zeroed when zero_pad != 0. */
int rs_ntt_inverse_embed(rs_ntt *t, rs_batch *dst, size_t dst_poly, const rs_batch *src, size_t src_poly,
                         size_t n_polys, int zero_pad);

/* ---- element-wise ops: baseOperator (math/bigpoly/base_op.go:49-171, vec.go:9-121) ------ */
int rs_vec_add(rs_batch *out, size_t op, const rs_batch *a, size_t ap, const rs_batch *b, size_t bp, size_t n_polys);
int rs_vec_sub(rs_batch *out, size_t op, const rs_batch *a, size_t ap, const rs_batch *b, size_t bp, size_t n_polys);
int rs_vec_neg(rs_batch *out, size_t op, const rs_batch *a, size_t ap, size_t n_polys);
int rs_vec_mul(rs_batch *out, size_t op, const rs_batch *a, size_t ap, const rs_batch *b, size_t bp, size_t n_polys);
int rs_vec_mul_add(rs_batch *out, size_t op, const rs_batch *a, size_t ap, const rs_batch *b, size_t bp, size_t n_polys);
int rs_vec_mul_sub(rs_batch *out, size_t op, const rs_batch *a, size_t ap, const rs_batch *b, size_t bp, size_t n_polys);
/* scalar c: one Montgomery element on the host */
int rs_vec_scalar_mul(rs_batch *out, size_t op, const rs_batch *a, size_t ap, const uint64_t *c, size_t n_polys);
int rs_vec_scalar_mul_add(rs_batch *out, size_t op, const rs_batch *a, size_t ap, const uint64_t *c, size_t n_polys);
int rs_vec_scalar_mul_sub(rs_batch *out, size_t op, const rs_batch *a, size_t ap, const uint64_t *c, size_t n_polys);

/* ---- Buckler prover kernels -------------------------------------------------------------- */
/* One term of an arithmetic constraint: coeff * pw * prod(w)  (buckler/constraint.go:33-69).
 * Terms of one constraint are contiguous and share `constraint`. */
#define RS_TERM_MAX_W 8
typedef struct rs_term {
    int32_t constraint;          /* constraint index (terms grouped, ascending) */
    int32_t coeff;               /* index into coeffs[] (Montgomery scalars, host) */
    int32_t pw;                  /* public-witness poly index in `pw` batch, or -1 */
    int32_t n_w;                 /* number of witness factors (<= RS_TERM_MAX_W) */
    int32_t w[RS_TERM_MAX_W];    /* witness poly indices in `w` batch */
} rs_term;
/* linCheckVec (buckler/prover.go:409-413): out[poly][i] = c^i for i < n, c a Montgomery element (host) */
int rs_vec_powers(rs_batch *out, size_t poly, size_t n, const uint64_t *c_mont);
/* Encoder.RandEncodeTo's blinding step (buckler/encoder.go:50-54) for n polynomials of a batch at once, after
 * rs_ntt_inverse_embed: coefficient N <- blind[i] (a MustSetRandom element, Montgomery limbs, host), coefficient 0 -= it. */
int rs_rand_encode_blind(rs_batch *b, size_t poly0, size_t n, size_t N, const uint64_t *blind_mont);

/* Prover.evalCircuit (buckler/prover.go:355-379) fused into one pass:
 * out[j] = sum_c batch_const * sum_{terms of c} coeff * pw[j] * prod w[j]. */
int rs_eval_circuit(rs_batch *out, size_t out_poly, const rs_batch *pw, const rs_batch *w,
                    const rs_term *terms, size_t n_terms, const uint64_t *coeffs, size_t n_coeffs,
                    const uint64_t *batch_const);
/* CyclicEvaluator.QuoRemByVanishing (math/bigpoly/cyclic.go:18-37): p = quo*(X^N-1) + rem. */
int rs_quorem_vanishing(rs_batch *quo, size_t quo_poly, rs_batch *rem, size_t rem_poly,
                        const rs_batch *p, size_t p_poly, size_t N);
/* CyclotomicEvaluator.AutTo (math/bigpoly/cyclotomic.go:29-82); out must not alias p. */
int rs_aut(rs_batch *out, size_t out_poly, const rs_batch *p, size_t p_poly, size_t n_polys, long idx, int is_ntt);
/* nttChecker.TransposeTo pre-step (buckler/linear.go:38-41): out[i] = v[N-1-i] * N. */
int rs_reverse_scale(rs_batch *out, size_t out_poly, const rs_batch *v, size_t v_poly, const uint64_t *scale);
/* Poly.Evaluate (math/bigpoly/poly.go:64-76) of the first n_elems coefficients at x; result to host */
int rs_poly_evaluate(const rs_batch *p, size_t poly, size_t n_elems, const uint64_t *x, uint64_t *result);
/* the closing loop of jindo.Prover.Evaluate (jindo/prover.go:318-321): n polynomials (batches[i], polys[i], first lens[i]
 * coefficients) at the same point in one launch and one copy; results: n x L64 Montgomery limbs (host) */
int rs_poly_evaluate_many(const rs_batch *const *batches, const size_t *polys, const size_t *lens, size_t n, const uint64_t *x,
                          uint64_t *results);

/* ---- multi-GPU plumbing (SURVEY 8b "rs_allgather_commitments", 8e) ---------------------------------------------------
 * One process per GPU.  The only exchanges of the path are tiny: per-shard commitments / evaluations into the host
 * Fiat-Shamir transcript, and the InCommit slices of a column-sharded Commit (ringo_jindo.h).  NCCL is loaded at run time.
 * A communicator can come from the caller (any ncclComm_t) or be made here: rank 0 calls rs_nccl_unique_id, ships the 128
 * bytes to the other ranks over its own control plane, and every rank calls rs_nccl_comm_create after rs_init(device). */
int rs_nccl_unique_id(uint8_t *id128);
int rs_nccl_comm_create(const uint8_t *id128, int rank, int world, void **comm_out);
void rs_nccl_comm_destroy(void *comm);
/* every rank contributes n words and receives world * n words, rank-major (host buffers) */
int rs_allgather_u64(void *comm, int rank, int world, const uint64_t *host_in, size_t n, uint64_t *host_out);

#ifdef __cplusplus
}
#endif
#endif /* RINGO_CUDA_H */
