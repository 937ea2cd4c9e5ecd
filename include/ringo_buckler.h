/*
 * ringo_buckler.h -- C ABI of the compiled Buckler prover (libringo_b200.so).
 *
 * Replaces buckler.Prover.Prove (buckler/prover.go:42-353) as ONE call: the round structure, the gnark-crypto
 * fiat-shamir transcript over SHA-256 (prover.go:113-121), the SHAKE128 transcript of jindo.Prover.Evaluate
 * (jindo/prover.go:221-300) and the launch sequence run in C++ inside the library; every polynomial operation is a
 * kernel of ringo_cuda.h / ringo_jindo.h on device-resident batches.  A Go caller keeps what only Go can do:
 *   - buckler.Compile (reflection walker, compile.go) -> the plain-data circuit description below,
 *   - prvWalk + the big.Int witness decompositions (prover.go:52-111) -> pw / w as Montgomery elements,
 *   - all randomness (crypto/rand MustSetRandom draws and the Jindo tapes), explicit inputs exactly like
 *     rs_jindo_tape, so that a proof is a deterministic function of its inputs and can be compared bit for bit.
 * The projection checker (prover.go:166-207; unused by every example and test of the reference) is not covered.
 */
#ifndef RINGO_BUCKLER_H
#define RINGO_BUCKLER_H

#include "ringo_jindo.h"
#include "ringo_sampler.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rs_buckler rs_buckler;

/* one LinearChecker with its constraints (buckler/linear.go:12-73, context.go linCheckConstraints) */
typedef struct rs_lin_checker {
    int32_t kind;            /* 0: nttChecker (linear.go:20-43); 1: autChecker (linear.go:46-73) */
    int32_t is_ntt;          /* autChecker.isNTT */
    int64_t idx_inv;         /* autChecker: idx^-1 mod 2N, the index TransposeTo applies (linear.go:69-73) */
    const int32_t *pairs;    /* n_pairs x (wOut, wIn) witness ids */
    size_t n_pairs;
} rs_lin_checker;

/* what buckler.Compile leaves in Prover.ctx, flattened (buckler/context.go:11-66) */
typedef struct rs_buckler_circuit {
    size_t rank, emb_rank;               /* ctx.rank, polyEval rank (compile.go:180-186) */
    size_t w_cnt, pw_cnt;
    size_t arith_max_rank, sum_max_rank; /* ctx.arithCheckMaxRank / sumCheckMaxRank */
    const rs_term *arith_terms; size_t n_arith_terms;       /* ctx.arithConstraints, one rs_term per monomial */
    const uint64_t *arith_coeffs; size_t n_arith_coeffs;    /* Montgomery elements */
    const rs_term *sum_terms; size_t n_sum_terms;           /* ctx.sumCheckConstraints (empty: no sum check) */
    const uint64_t *sum_coeffs; size_t n_sum_coeffs;
    const rs_lin_checker *lin; size_t n_lin;                /* ctx.linCheckers in order (empty: no linear check) */
} rs_buckler_circuit;

/* the random draws of one Prove, in the reference's order */
typedef struct rs_buckler_rand {
    const uint64_t *blind;          /* w_cnt elements: RandEncodeTo's MustSetRandom (encoder.go:50-54), Montgomery */
    const uint64_t *lin_mask;       /* sumCheckMask(2 rank) (prover.go:381-397): 2*rank elements, Montgomery */
    const uint64_t *lin_mask_sum;   /* its maskSum, one element */
    const uint64_t *sum_mask;       /* sumCheckMask(sumCheckMaxRank): sum_max_rank elements */
    const uint64_t *sum_mask_sum;
    const rs_jindo_tape *tapes;     /* one tape per polyProver.Commit, in call order */
    size_t n_tapes;
    int32_t tapes_pinned;           /* != 0: the tapes' host arrays are pinned memory (rs_pinned_alloc, rs_tape_builder) and are read in
                                       place; otherwise they are staged through the prover's own pinned arena first */
} rs_buckler_rand;

/* buckler.Proof (buckler/entities.go) as caller-owned host buffers; sizes follow from the circuit and the Jindo
 * parameters (batch = rs_buckler_batch):
 *   witness       batch x out_msis polys over ringQOut (n_qout*d words each)   Proof.Witness[i].Value
 *   evals         batch x L64 Montgomery limbs                                  Proof.Evals
 *   partial       cols polys over ringQ; partial_mask 1; encode rows; mlwe mlwe+in_msis; incommit (cols+1)*in_msis (ringQOut)
 *   lin_mask_sum / sum_mask_sum / eval_point: L64 Montgomery limbs each (eval_point is the evalPoint challenge, for tests) */
typedef struct rs_buckler_proof {
    uint64_t *witness;
    uint64_t *lin_mask_sum, *sum_mask_sum;
    uint64_t *evals;
    uint64_t *partial, *partial_mask, *encode, *mlwe, *incommit;
    uint64_t *eval_point;
} rs_buckler_proof;

/* NewProver half of buckler.Compile (compile.go:150-200): transformers (Encoder.ntt at rank, polyEval at emb_rank, the
 * nttChecker's negacyclic transformer), persistent work batches.  `j` is the Jindo prover built for
 * jindo.NewParameters(q, jindo rank, batch); crs = CommitKey.WriteRawTo bytes (entities.go:75-77).  The circuit arrays are copied. */
int rs_buckler_create(rs_field *f, rs_jindo *j, const rs_buckler_circuit *c, const uint8_t *crs, size_t crs_len, rs_buckler **out);
void rs_buckler_destroy(rs_buckler *b);
/* number of commitments of a proof = Context.batch() (context.go) */
size_t rs_buckler_batch(const rs_buckler *b);
/* Lets the prover draw its own randomness, as the reference's Prove does: an rs_tape_builder (ringo_sampler.h) with n_threads
 * worker threads and one slot per Commit samples the COSAC rows, the mask column's noise and every MustSetRandom element from OS
 * randomness; the encoder's TwinCDT rows and the data columns' MLWE noise come from the Jindo prover's device streams, which the
 * caller must have set (rs_jindo_set_twin_stream, rs_jindo_set_mlwe_stream).  q_le: the field modulus, L64 limbs.  After this,
 * rs_buckler_prove accepts rand == NULL. */
int rs_buckler_enable_sampling(rs_buckler *b, const rs_jindo_noise *noise, const uint64_t *q_le, int n_threads);
/* Prover.Prove.  pw / w: pw_cnt x rank and w_cnt x rank Montgomery elements (host), decomposition witnesses filled in.
 * phase_ms (optional, 16 doubles): wall milliseconds of [0] encode+NTT, [1] commits, [2] arithCheck, [3] linCheck, [4] sumCheck, [5] Evaluate,
 * [6] tape staging, [7] total; Evaluate's parts: [8] batch combination of the openings, [9] Partial, [10] InCommit download + transcript, [11] responses,
 * [12] Poly.Evaluate of the committed vectors.  The stream is synchronised at every boundary when phase_ms is given. */
/* on != 0: the self-sampled proofs take their COSAC rows from the Jindo prover's device sampler (rs_jindo_set_cosac_stream must
 * have been called) instead of the host worker threads, which then only draw the MustSetRandom elements and the mask noise */
int rs_buckler_set_device_cosac(rs_buckler *b, int on);
/* rand == NULL: self-sampled (see rs_buckler_enable_sampling). */
int rs_buckler_prove(rs_buckler *b, const uint64_t *pw, const uint64_t *w, const rs_buckler_rand *rand,
                     rs_buckler_proof *proof, double *phase_ms);

/* Verifier.Verify (buckler/verifier.go:29-315) for a proof in the layout rs_buckler_prove writes.  pw: the pw_cnt x rank public
 * witnesses (Montgomery, the two-norm base / mask rows filled in as Prove's caller fills them); in_com_dcmp_two_nm / res_two_nm:
 * Parameters.inComDcmpTwoNm / resTwoNm of the Jindo parameters.  Both transcripts are recomputed from the proof; the Jindo evaluation
 * proof is checked by rs_jindo_verify; Encode(.).Evaluate(x) of the public witnesses and the linear-check vectors runs on the device.
 * *result: 0 = accepted, otherwise the first failing check (1 polyVerifier.Verify, 2 arithmetic, 3 / 4 linear, 5 / 6 sum check). */
int rs_buckler_verify(rs_buckler *b, const uint64_t *pw, const rs_buckler_proof *proof, double in_com_dcmp_two_nm, double res_two_nm,
                      int *result);

/* the two hash functions of the transcripts, exported for the shim's tests (FIPS 180-4 / FIPS 202) */
void rs_sha256(const uint8_t *data, size_t len, uint8_t out32[32]);
void rs_shake128(const uint8_t *data, size_t len, uint8_t *out, size_t out_len);

#ifdef __cplusplus
}
#endif
#endif /* RINGO_BUCKLER_H */
