/*
 * ringo_jindo.h -- C ABI for the Jindo polynomial-commitment hot path (libringo_b200.so).
 *
 * Replaces the inner loops of jindo.Prover.Commit / Evaluate (jindo/prover.go:45-324), the
 * encoder (jindo/encoder.go:113-201), the RNS reconstruction / cut-off pipeline
 * (jindo/rns.go:76-114, prover.go:164-176, :190-200) and the lattigo ring.Ring calls they make
 * (NTT/INTT, MForm/IMForm, MulCoeffsMontgomeryThenAdd, Add, MulScalarThenSub).
 *
 * Everything that is a *convention* of the host library stays an INPUT, so this library embeds
 * no assumption about lattigo or Go's math package:
 *   - RNS primes and their 2d-th primitive roots psi (lattigo ring.SubRing),
 *   - the commit key polynomials (jindo.NewCommitKey, entities.go:21-61),
 *   - the encoder's float64 constants deltaInv (encoder.go:50-67),
 *   - the TwinCDT tables (math/csprng/gaussian_twin_cdt.go:13-35) and all randomness:
 *     raw uniform uint64 draws for TwinCDT rows, already-sampled integers for the rows the
 *     reference samples with data-dependent rejection loops (COSAC / rounded Gaussian),
 *     and the MustSetRandom field elements of the first/last/mask rows.
 * The Go shim produces them with the reference's own csprng; tests produce them with the oracle.
 *
 * Ring polynomials use lattigo's layout: Coeffs[limb][coeff] uint64, limb-major.
 */
#ifndef RINGO_JINDO_H
#define RINGO_JINDO_H

#include "ringo_cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

#define RS_JINDO_MAX_LIMBS 4
#define RS_JINDO_MAX_EXP 64

typedef struct rs_jindo rs_jindo;
typedef struct rs_opening rs_opening;

/* jindo.Parameters (jindo/params.go:53-124) -- the fields the prover's arithmetic needs */
typedef struct rs_jindo_params {
    int32_t d;                 /* ring degree (256) */
    int32_t slots;             /* params.slots */
    int32_t exp;               /* params.ecd.exp */
    uint64_t base;             /* params.ecd.base */
    int32_t rows, cols;
    int32_t in_msis, mlwe, out_msis;
    int32_t log_in_cut, log_out_cut;
    int32_t n_q, n_qout;                       /* RNS limbs of ringQ / ringQOut */
    uint64_t q[RS_JINDO_MAX_LIMBS];            /* ringQ.SubRings[l].Modulus */
    uint64_t qout[RS_JINDO_MAX_LIMBS];
    uint64_t psi_q[RS_JINDO_MAX_LIMBS];        /* primitive 2d-th root used by ringQ's NTT (canonical) */
    uint64_t psi_qout[RS_JINDO_MAX_LIMBS];
    double delta_inv[RS_JINDO_MAX_EXP];        /* Encoder.deltaInv (encoder.go:50-67) */
    /* TwinCDT sampler of the encoder (sigma = ecdStdDev): 128 tables of cdt_len entries */
    int32_t cdt_len;
    int64_t cdt_tail_lo;
    double cdt_sigma;
} rs_jindo_params;

/* Standard deviations of jindo.Parameters (params.go:303-309) the COSAC / rounded samplers need */
typedef struct rs_jindo_noise {
    double ecd_blind_std;      /* row 0 of a data column             (COSAC)   */
    double mask_blind_std;     /* row 0 of the mask column           (COSAC)   */
    double mask_std;           /* other rows of the mask column      (COSAC)   */
    double mask_mlwe_std;      /* MLWE noise of the mask column      (rounded) */
} rs_jindo_noise;

/* jindo.NewParameters(q, targetN, batch) (jindo/params.go:126-321) for callers that do not have the Go library at hand: the
 * parameter search in the reference's float64 order, the RNS primes and psi by lattigo's conventions (restated, unpinned),
 * Encoder.deltaInv and the TwinCDT geometry -- everything rs_jindo_create[_crs] and the samplers take.  q_le: the field modulus
 * (must be b^k + 1, params.go:18-40).  rs_jindo_cdt_tables fills the 128 x p.cdt_len TwinCDT tables (computeCDT,
 * math/csprng/gaussian_twin_cdt.go:13-35, centres i/128, sigma = ecd_std). */
typedef struct rs_jindo_derived {
    rs_jindo_params p;
    rs_jindo_noise noise;
    double ecd_std, mlwe_std;              /* params.ecdStdDev / mlweStdDev (sigma of the two TwinCDT samplers) */
    double res_two_nm, in_com_dcmp_two_nm; /* the verifier's norm bounds (params.go:310-311) */
    double com_size, pf_size;              /* the size estimate the search minimises, bits */
    uint64_t rank;                         /* params.rank = cols * (rows - 1) * slots */
    int32_t batch, in_com_dcmp_len, q_bits, qout_bits;
} rs_jindo_derived;
int rs_jindo_new_parameters(const uint64_t *q_le, int L64, uint64_t target_n, int batch, rs_jindo_derived *out);
int rs_jindo_cdt_tables(const rs_jindo_derived *d, uint64_t *tables);

/* NewProver (jindo/prover.go:28-40): uploads the commit key (ck.In [in_msis][rows], ck.MLWE
 * [in_msis][mlwe] over ringQ; ck.Out [out_msis][(cols+1)*in_msis] over ringQOut) and the
 * 128 x cdt_len TwinCDT tables. */
int rs_jindo_create(rs_field *f, const rs_jindo_params *p, const uint64_t *ck_in, const uint64_t *ck_mlwe,
                    const uint64_t *ck_out, const uint64_t *cdt_tables, rs_jindo **out);
/* NewProver with NewCommitKey(params, crs) evaluated in HBM (jindo/entities.go:21-61): the key is
 * the SampleN stream (math/csprng/uniform.go:87-95, rejection included) of a UniformSampler seeded with the CRS, in
 * the reference's draw order (In, MLWE, Out; polynomial -> coefficient -> limb).  key32 || iv16 = SHA-384(crs)
 * (uniform.go:38-46), hashed by the caller like rs_jindo_set_twin_stream's.  No key bytes cross PCIe.
 * rs_jindo_commit_key_download copies ck.In (which = 0), ck.MLWE (1) or ck.Out (2) back (tests; the shim's
 * CommitKey view), layout as rs_jindo_create's inputs. */
int rs_jindo_create_crs(rs_field *f, const rs_jindo_params *p, const uint8_t *key32, const uint8_t *iv16,
                        const uint64_t *cdt_tables, rs_jindo **out);
int rs_jindo_commit_key_download(const rs_jindo *j, int which, uint64_t *out);
/* the parameter block the prover was created with (the compiled Buckler prover sizes its proof buffers from it) */
int rs_jindo_get_params(const rs_jindo *j, rs_jindo_params *out);
void rs_jindo_destroy(rs_jindo *j);

/* GPU-resident uniform source of the encoder's TwinCDT sampler.
 * Reproduces csprng.UniformSampler (math/csprng/uniform.go:26-70) in HBM: AES-256-CTR keyed by
 * key32 || iv16 = SHA-384(seed) (uniform.go:38-46, hashed by the caller), 8-byte little-endian draws out
 * of the XOR-accumulating 8 KiB buffer.  The stream is stateful like the Go sampler: every Commit whose
 * tape has twin_u == NULL consumes the next n_twin*d draws.  rs_jindo_twin_stream_draw hands the next n
 * draws to the host instead (tests; sample_n / sample_float callers). */
int rs_jindo_set_twin_stream(rs_jindo *j, const uint8_t *key32, const uint8_t *iv16);
int rs_jindo_twin_stream_draw(rs_jindo *j, size_t n, uint64_t *out);

/* GPU-resident Prover.mlweSampler (jindo/prover.go:38, :136-138): a TwinCDTGaussianSampler that is only ever asked for
 * centre 0, i.e. exactly one uniform draw and one binary search in table 0 per sample (gaussian_twin_cdt.go:77-97: c0 == c1,
 * so the slow path is unreachable).  key32 || iv16 = SHA-384(seed) of that sampler's UniformSampler; table0 = computeCDT(0,
 * mlweStdDev) (len entries, gaussian_twin_cdt.go:13-35; rs_compute_cdt), tail_lo = -ceil(9 mlweStdDev).  Stateful like the Go
 * object: every Commit whose tape has mlwe_mask_only set consumes the next cols*(in_msis+mlwe)*d draws, column-major in the
 * reference's call order. */
int rs_jindo_set_mlwe_stream(rs_jindo *j, const uint8_t *key32, const uint8_t *iv16, const uint64_t *table0, int len, int64_t tail_lo);
/* the next n samples of that sampler (tests) */
int rs_jindo_mlwe_stream_sample(rs_jindo *j, size_t n, int64_t *out);

/* GPU-resident COSAC sampler of the encoder (Encoder.cosac, jindo/encoder.go:72-74, :166-171; math/csprng/gaussian_cosac.go): the
 * rows the reference samples with data-dependent rejection loops -- row 0 of every data column (ecdBlindStdDev) and every row of the
 * mask column (maskBlindStdDev / maskStdDev).  One sampler object is sequential, so the device uses the reference's own model of
 * parallelism, one fresh sampler per worker (Prover.SafeCopy, prover.go:327-338): worker w = (running count of COSAC coefficients
 * sampled so far) owns COSACSampler{baseSampler: AES-256-CTR(key_b, IV_b + (w << 16)), roundedSampler: {AES-256-CTR(key_r, IV_r +
 * (w << 16))}}, key || iv = the 48 bytes given here (SHA-384 of a seed, hashed by the caller like rs_jindo_set_twin_stream's).
 * Every Commit whose tape has cosac_g == NULL draws its COSAC polynomials from it, digits and float64 centres computed on the
 * device in the reference's order.  rs_jindo_cosac_sample (tests) returns what a Commit with this first row / these mask rows
 * draws (n_polys = cols + n_mask_rows polynomials of 256 samples), the index of its first worker, and whether a worker ran out
 * of its first 8 KiB of keystream (1024 draws; a sample takes about five). */
int rs_jindo_set_cosac_stream(rs_jindo *j, const uint8_t *key48_base, const uint8_t *key48_rounded, double ecd_blind_std,
                              double mask_blind_std, double mask_std);
/* workers handed out so far and the sticky overflow flag (synchronises the stream) */
int rs_jindo_cosac_status(rs_jindo *j, uint64_t *workers, int *overflowed);
int rs_jindo_cosac_sample(rs_jindo *j, const uint64_t *first_row, const uint64_t *mask_rows, size_t n_mask_rows, int64_t *out,
                          uint64_t *worker0, int *overflowed);

/* Randomness / noise inputs of one Commit, in the reference's consumption order. */
typedef struct rs_jindo_tape {
    const uint64_t *first_row;   /* cols*slots field elements (Montgomery)       prover.go:65-86 */
    const uint64_t *last_row;    /* cols*slots field elements */
    const uint64_t *mask_rows;   /* n_mask_rows*slots field elements, encode order prover.go:93-115 */
    size_t n_mask_rows;
    const uint64_t *twin_u;      /* one uniform draw per TwinCDT sample, call order  encoder.go:168-169;
                                    NULL: drawn on the device from the stream set by rs_jindo_set_twin_stream */
    size_t n_twin_u;
    const int64_t *cosac_g;      /* sampled g for the COSAC rows: data columns' row 0 (col order), then
                                    the mask column's rows in encode order; d values each;
                                    NULL: sampled on the device from the stream set by rs_jindo_set_cosac_stream */
    size_t n_cosac_polys;
    const int64_t *mlwe;         /* (cols+1)*(in_msis+mlwe)*d noise samples          prover.go:130-139;
                                    with mlwe_mask_only: just the mask column's (in_msis+mlwe)*d samples */
    int32_t on_device;           /* != 0: every pointer above is a device pointer (tape already resident in HBM) */
    int32_t mlwe_mask_only;      /* != 0: the data columns' MLWE noise (mlweSampler.Sample(0), prover.go:136-138) is drawn on the
                                    device from the stream set by rs_jindo_set_mlwe_stream; `mlwe` holds the mask column only
                                    (roundedSampler, host) */
} rs_jindo_tape;

/* Prover.Commit (jindo/prover.go:45-62).  v: device batch holding the committed vector
 * (poly v_poly, first len coefficients, len <= rank).  commitment_out: out_msis polys over
 * ringQOut (host, n_qout*d words each) -- Commitment.Value in NTT+Montgomery form.
 * *open: NULL on entry -> a new Opening is allocated; otherwise the given Opening is overwritten. */
int rs_jindo_commit(rs_jindo *j, const rs_batch *v, size_t v_poly, size_t len, const rs_jindo_tape *tape,
                    rs_opening **open, uint64_t *commitment_out);
/* Same Commit without the trailing synchronisation: every copy and kernel is enqueued on the library's stream and the call
 * returns.  commitment_out_pinned and every host buffer of the tape must be pinned memory (rs_pinned_alloc) and stay untouched
 * until the caller has synchronised (rs_sync).  Lets a proof's 12-43 small Commits run back to back while the host hashes the
 * commitments that are already out (the compiled Buckler prover, ringo_buckler.h). */
int rs_jindo_commit_async(rs_jindo *j, const rs_batch *v, size_t v_poly, size_t len, const rs_jindo_tape *tape,
                          rs_opening **open, uint64_t *commitment_out_pinned);
/* Same Commit with v in host memory (len elements, L64 limbs each, Montgomery form; pinned memory for
 * overlap): v is uploaded in pieces on a copy stream while earlier pieces are already being encoded, so
 * the upload hides behind the encode kernels.  The tape must be host resident too. */
int rs_jindo_commit_host(rs_jindo *j, const uint64_t *v_host, size_t len, const rs_jindo_tape *tape,
                         rs_opening **open, uint64_t *commitment_out);
/* Column-sharded Commit (SURVEY 8e iii): the cols+1 columns of the committed matrix are independent until
 * the outer commit (prover.go:89-177 vs :180-202).  A rank runs commitColTo for columns [col_begin, col_end)
 * (the mask column is index cols) with the WHOLE tape -- offsets into it are global, so every shard consumes
 * exactly the draws the single-GPU Commit gives those columns -- which fills Opening.Encode / MLWE for these
 * columns and InCommit[col*in_msis, (col+1)*in_msis).  After the shards' InCommit slices have been gathered
 * into rs_opening_incommit_device() (n_qout*d words per polynomial, (cols+1)*in_msis polynomials; e.g. an NCCL
 * all-gather on that device pointer), rs_jindo_outer_commit produces Commitment.Value. */
int rs_jindo_commit_columns(rs_jindo *j, const rs_batch *v, size_t v_poly, size_t len, const rs_jindo_tape *tape,
                            int col_begin, int col_end, rs_opening **open);
int rs_jindo_outer_commit(rs_jindo *j, const rs_opening *open, uint64_t *commitment_out);
void *rs_opening_incommit_device(rs_opening *open);
/* The same split with the exchange inside the library (NCCL on the library's stream, no host synchronisation; nccl_comm is an
 * ncclComm_t -- the caller's, or one made by rs_nccl_comm_create):
 *   rs_jindo_shard_columns         rank r of world owns columns [r*per, min((r+1)*per, cols+1)), per = ceil((cols+1)/world)
 *   rs_jindo_outer_commit_sharded  Commitment.Value from the shards WITHOUT gathering InCommit: every rank multiplies ck.Out
 *                                  by its own columns' InCommit, the out_msis partial sums (KiB) are all-reduced, the cut-off
 *                                  pipeline runs replicated; bit-identical to the unsharded Commit on every rank
 *   rs_jindo_allgather_incommit    all InCommit polynomials into every rank's Opening (Proof.InCommit, Evaluate), one in-place
 *                                  ncclAllGather; may be issued after the commitment is out, it is not on Commit's critical path */
int rs_jindo_shard_columns(const rs_jindo *j, int rank, int world, int *col_begin, int *col_end);
int rs_jindo_outer_commit_sharded(rs_jindo *j, const rs_opening *open, int rank, int world, void *nccl_comm, uint64_t *commitment_out);
int rs_jindo_allgather_incommit(rs_jindo *j, rs_opening *open, int rank, int world, void *nccl_comm);
void rs_opening_free(rs_opening *o);
/* accessors (tests / the Go shim's Opening view): kind 0 = Encode[col][row] (ringQ),
 * 1 = MLWE[col][idx] (ringQ), 2 = InCommit[idx] (ringQOut, col ignored) */
int rs_opening_download(const rs_opening *o, int kind, size_t col, size_t idx, uint64_t *poly_out);
/* all polynomials of one kind in one copy; kind 2 (InCommit, (cols+1)*in_msis polys over ringQOut) is what Proof.InCommit carries */
int rs_opening_download_all(const rs_opening *o, int kind, uint64_t *out);

/* Prover.Evaluate (jindo/prover.go:205-324), split around the host transcript:
 *   phase 1: Partial[i] = sum_j left[j] (.) Encode[i][j]  for i <= cols  (prover.go:281-289)
 *            left: rows polys over ringQ (encoded leftVec, NTT+Montgomery form, host)
 *            partial_out: (cols+1) polys (host); index cols is PartialMask
 *   phase 2: responses  Encode[i] = Enc[cols][i] + sum_j chal[j] (.) Enc[j][i]   (prover.go:304-316)
 *            chals: cols polys over ringQ (host); enc_out: rows polys, mlwe_out: mlwe+in_msis polys */
int rs_jindo_partial(rs_jindo *j, const rs_opening *o, const uint64_t *left, uint64_t *partial_out);
int rs_jindo_respond(rs_jindo *j, const rs_opening *o, const uint64_t *chals, uint64_t *enc_out, uint64_t *mlwe_out);
/* The same two phases with Evaluate's host-side vectors produced on the device, so that only the evaluation point
 * and the transcript's challenge bytes cross the boundary:
 *   rs_jindo_partial_at    left = Encoder.EncodeTo of leftVec(x) (jindo/utils.go:63-73, encoder.go:113-117); x_mont: the
 *                          evaluation point, L64 Montgomery limbs (host)
 *   rs_jindo_respond_bytes chals = encodeChallengeTo (utils.go:21-46) of cols 16-byte SHAKE outputs (host)
 *   rs_jindo_left_vec / rs_jindo_encode_challenges  hand the encoded vectors back (tests; the verifier-side shim):
 *                          rows polys over ringQ / n polys over ringQ (ring_out = 0) or ringQOut (1) */
int rs_jindo_partial_at(rs_jindo *j, const rs_opening *o, const uint64_t *x_mont, uint64_t *partial_out);
int rs_jindo_respond_bytes(rs_jindo *j, const rs_opening *o, const uint8_t *chal_bytes, uint64_t *enc_out, uint64_t *mlwe_out);
int rs_jindo_left_vec(rs_jindo *j, const uint64_t *x_mont, uint64_t *left_out);
int rs_jindo_encode_challenges(rs_jindo *j, const uint8_t *chal_bytes, size_t n, int ring_out, uint64_t *polys_out);
/* batch > 1 (prover.go:251-265): acc += open (.) challenge, over every polynomial of the opening;
 * _bytes: the challenge polynomial of both rings is encoded on the device from its 16 bytes (prover.go:230-249) */
/* Verifier.Verify (jindo/verifier.go:50-282) with the proof's arithmetic on the device: IMForm / INTT of every proof
 * polynomial, the outer and inner commitment equations (MulCoeffsMontgomeryThenSub against ck.Out / ck.In / ck.MLWE, the
 * cut-off scalars, BasisExtender.ModUpQtoP as a centred lift), CRT reconstruction and the two-norms of verifyNorm, the
 * consistency equation, and the reconstruction of the Partial polynomials that verifyEval decodes.  The transcript stays with
 * the caller, as in Evaluate: batch_bytes (16 per commitment, NULL when batch == 1) and chal_bytes (16 per column) are the
 * SHAKE128 outputs of verifier.go:61-94.  The proof polynomials are host arrays in the form Evaluate returns them.
 * *checks: bit 0 verifyOuterCommitment, bit 1 verifyInnerCommitment, bit 2 verifyConsistency, bit 3 verifyEval; the proof is
 * accepted iff *checks == 15.  in_com_dcmp_two_nm / res_two_nm: Parameters.inComDcmpTwoNm / resTwoNm (params.go). */
typedef struct rs_jindo_proof {
    const uint64_t *partial;       /* cols polys over ringQ          Proof.Partial */
    const uint64_t *partial_mask;  /* 1 poly over ringQ              Proof.PartialMask */
    const uint64_t *encode;        /* rows polys over ringQ          Proof.Encode */
    const uint64_t *mlwe;          /* mlwe + in_msis polys, ringQ    Proof.MLWE */
    const uint64_t *incommit;      /* (cols+1)*in_msis polys, ringQOut  Proof.InCommit */
} rs_jindo_proof;
int rs_jindo_verify(rs_jindo *j, const uint64_t *x_mont, const uint64_t *commitments, int batch, const uint8_t *batch_bytes,
                    const uint8_t *chal_bytes, const uint64_t *ys_mont, const rs_jindo_proof *pf, double in_com_dcmp_two_nm,
                    double res_two_nm, int *checks);

/* Wire format of the transcript inputs, for hosts that do not have lattigo at hand:
 *   rs_jindo_poly_write_to          [lattigo] ring.Poly.WriteTo framing of one polynomial given as n_limbs x d words: u64 LE
 *                                   limb count, then per limb u64 LE length d and d coefficients as u64 LE.
 *   rs_jindo_commitment_write_raw   Commitment.WriteRawTo (jindo/entities.go:96-100): the out_msis polynomials of
 *                                   Commitment.Value one after the other.  NewCommitment allocates them in ringQ
 *                                   (entities.go:85-94) although their data live in ringQOut, so every polynomial is
 *                                   framed with max(n_q, n_qout) limbs, the limbs beyond n_qout all zero.
 * Both return the number of bytes written, or a negative error code (RS_ERR_ARG when cap is too small); with out == NULL
 * they only return the size. */
long long rs_jindo_poly_write_to(const uint64_t *poly, int n_limbs, int d, uint8_t *out, size_t cap);
long long rs_jindo_commitment_write_raw(const rs_jindo *j, const uint64_t *commitment, uint8_t *out, size_t cap);

int rs_opening_alloc_zero(rs_jindo *j, rs_opening **out);
int rs_opening_mul_add(rs_jindo *j, rs_opening *acc, const rs_opening *o, const uint64_t *chal_q, const uint64_t *chal_qout);
int rs_opening_mul_add_bytes(rs_jindo *j, rs_opening *acc, const rs_opening *o, const uint8_t *chal16);
/* the whole loop of prover.go:230-265 in one pass: *out = sum_i opens[i] (.) encodeChallenge(chal_bytes[16 i .. 16 i + 16)); n
 * challenge polynomials per ring are encoded in one launch each, every opening is read once and the new Opening written once */
int rs_opening_combine_bytes(rs_jindo *j, const rs_opening *const *opens, size_t n, const uint8_t *chal_bytes, rs_opening **out);

#ifdef __cplusplus
}
#endif
#endif /* RINGO_JINDO_H */
