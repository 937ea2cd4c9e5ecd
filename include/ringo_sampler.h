/*
 * ringo_sampler.h -- host-side randomness of jindo.Prover.Commit (libringo_b200.so, plain C++ on the host cores).
 *
 * The reference draws every random quantity of a Commit on the CPU (math/csprng + crypto/rand).  The device path
 * takes over the two bulk samplers whose stream positions are closed-form (the encoder's TwinCDT rows and the
 * mlweSampler, one uniform draw per sample: rs_jindo_set_twin_stream / rs_jindo_set_mlwe_stream).  What stays on
 * the host is what the reference samples with data-dependent rejection loops or straight from crypto/rand:
 *   - COSACSampler (math/csprng/gaussian_cosac.go:23-57)          row 0 of every data column, every mask-column row
 *   - RoundedGaussianSampler (gaussian_rounded.go:22-125)          the mask column's MLWE noise, COSAC's proposals
 *   - MustSetRandom elements (zp element.go:303-346)               last row, mask rows
 * This header is the product-side implementation of that part -- what a Go shim would otherwise do with the
 * reference's own csprng package -- so that an end-to-end Commit can be measured from "vector + seeds" with every
 * tape entry produced inside the timed region, on worker threads that overlap the previous Commit's kernels.
 *
 * Samplers are bit-exact restatements: same AES-256-CTR stream (key48 = SHA-384(seed) = key32 || iv16,
 * uniform.go:38-46, incl. the XOR-accumulating 8 KiB refill :64-70), same ziggurat tables, same float64 expressions
 * (compiled without FMA contraction, like Go on amd64).  key48 == NULL seeds from the OS (getrandom), which is what
 * the reference's constructors do with crypto/rand.
 */
#ifndef RINGO_SAMPLER_H
#define RINGO_SAMPLER_H

#include "ringo_jindo.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- csprng.UniformSampler (uniform.go:16-100) ------------------------------------------------------------- */
typedef struct rs_uniform rs_uniform;
rs_uniform *rs_uniform_create(const uint8_t *key48);
void rs_uniform_destroy(rs_uniform *u);
uint64_t rs_uniform_sample(rs_uniform *u);                          /* Sample()      :64-84 */
void rs_uniform_sample_many(rs_uniform *u, uint64_t *out, size_t n);
uint64_t rs_uniform_sample_n(rs_uniform *u, uint64_t n);            /* SampleN()     :87-95 */
double rs_uniform_sample_float(rs_uniform *u);                      /* SampleFloat() :96-100 */
void rs_uniform_read(rs_uniform *u, uint8_t *p, size_t n);          /* Read(): XORKeyStream(p, p) :58-61 */

/* ---- csprng.RoundedGaussianSampler / COSACSampler ------------------------------------------------------------ */
typedef struct rs_rounded rs_rounded;
rs_rounded *rs_rounded_create(const uint8_t *key48);
void rs_rounded_destroy(rs_rounded *r);
double rs_rounded_norm_float(rs_rounded *r);                        /* normFloat()  gaussian_rounded.go:77-117 */
int64_t rs_rounded_sample(rs_rounded *r, double center, double std_dev);   /* Sample() :119-125 */

typedef struct rs_cosac rs_cosac;
rs_cosac *rs_cosac_create(const uint8_t *key48_base, const uint8_t *key48_rounded);   /* NewCOSACSampler :15-20 */
void rs_cosac_destroy(rs_cosac *c);
int64_t rs_cosac_sample(rs_cosac *c, double center, double std_dev);                 /* Sample() :48-57 */

/* computeCDT (gaussian_twin_cdt.go:13-35): table of tailHi - tailLo + 1 entries, tailHi = ceil(9 sigma).
 * Returns the number of entries (<= cap) or a negative error code; *tail_lo = -tailHi. */
int rs_compute_cdt(double center, double sigma, uint64_t *table, int cap, int64_t *tail_lo);

/* ---- the host half of a Commit's randomness tape ------------------------------------------------------------- */
/* rs_jindo_noise (the standard deviations the host samplers need) is declared in ringo_jindo.h */

typedef struct rs_tape_builder rs_tape_builder;

/* q: the committed field's modulus (L64 little-endian limbs).  n_threads >= 1 worker threads.
 * seeds: NULL -> every sampler seeds itself from the OS, like NewProver does with crypto/rand.  Otherwise
 * 3 x 48 bytes key48 for, in NewProver's construction order (encoder.go:72-74, prover.go:36-38), the encoder's
 * COSAC base sampler, the COSAC sampler's rounded sampler, and the prover's roundedSampler; this deterministic
 * form requires n_threads == 1 (one sampler object = one sequential stream, as in the reference) and is what the
 * parity tests use.  With n_threads > 1 every worker owns independently seeded samplers -- the reference's own
 * concurrency model (Prover.SafeCopy hands every goroutine fresh samplers, prover.go:327-338). */
int rs_tape_builder_create(const rs_jindo_params *p, const rs_jindo_noise *noise, const uint64_t *q, int L64,
                           int n_threads, const uint8_t *seeds, rs_tape_builder **out);
/* same with n_slots >= 1 buffer sets instead of two (slot index 0 .. n_slots-1 in start / wait): the compiled Buckler prover
 * keeps one tape per Commit of a proof alive until the proof's last kernel has run */
int rs_tape_builder_create_slots(const rs_jindo_params *p, const rs_jindo_noise *noise, const uint64_t *q, int L64,
                                 int n_threads, int n_slots, const uint8_t *seeds, rs_tape_builder **out);
void rs_tape_builder_destroy(rs_tape_builder *b);
/* on != 0: the tapes carry cosac_g = NULL -- the COSAC rows are sampled on the device from the Jindo prover's own stream
 * (rs_jindo_set_cosac_stream); the builder still draws the first / last / mask row elements and the mask column's MLWE noise */
int rs_tape_builder_set_device_cosac(rs_tape_builder *b, int on);
/* n MustSetRandom field elements (rejection sampling over the builder's crypto/rand stand-in), Montgomery limbs */
int rs_tape_builder_random_elements(rs_tape_builder *b, uint64_t *out_mont, size_t n);
/* Prover.sumCheckMask (buckler/prover.go:381-397): mask_rank coefficients and maskSum */
int rs_tape_builder_sum_check_mask(rs_tape_builder *b, size_t rank, size_t mask_rank, uint64_t *mask_out, uint64_t *mask_sum_out);
/* crypto/rand stand-in for the MustSetRandom draws of the NEXT rs_tape_builder_start (tests): the bytes a
 * crypto/rand.Reader would return, consumed in the reference's order (last row, then the mask rows).  Without it
 * the elements come from an AES-CTR generator keyed from the OS. */
int rs_tape_builder_set_random_bytes(rs_tape_builder *b, const uint8_t *bytes, size_t n);
/* Starts producing the host entries of the tape of Commit(v[0:len]) into pinned buffer set `slot` (0 or 1) on the
 * worker threads and returns at once; v_host must stay valid until the matching wait.  Entries: last_row, first_row
 * (= v[i] - last[i-1], prover.go:76-83), mask_rows, cosac_g (base-b digits + float64 centres of encoder.go:149-164,
 * then COSACSampler.Sample per coefficient), and the mask column's MLWE samples. */
int rs_tape_builder_start(rs_tape_builder *b, const uint64_t *v_host, size_t len, int slot);
/* Blocks until slot is complete and fills *tape: host pointers into the builder's pinned buffers, twin_u = NULL
 * (device stream), mlwe = the mask column's (in_msis + mlwe) * d samples with mlwe_mask_only = 1.
 * *host_ms (optional): wall time the workers spent on this tape. */
int rs_tape_builder_wait(rs_tape_builder *b, int slot, rs_jindo_tape *tape, double *host_ms);

#ifdef __cplusplus
}
#endif
#endif /* RINGO_SAMPLER_H */
