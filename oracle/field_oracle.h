/*
 * oracle/field_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C) of the reference's big-field arithmetic
 * (gnark-crypto generated `zp` packages) and of math/bigpoly.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
 * may link or call this.  The CUDA product under ringo-snark_b200/ never does.
 *
 * Field element: little-endian uint64 limbs, Montgomery form, R = 2^(64 L),
 * always fully reduced (reference: jindo/internal/zp/element.go:24-37).
 *
 * Parity pinning: the per-modulus constants computed here (qInvNeg, rSquare)
 * are checked against the constants vendored in the reference's seven zp
 * packages (tests/golden/field_constants.json), and every op is checked
 * against Python integers mod q -- the same oracle the reference's own gopter
 * tests use (math/big), e.g. examples/mult/zp/element_test.go:946-986.
 */
#ifndef RINGO_FIELD_ORACLE_H
#define RINGO_FIELD_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_L 16

typedef struct orc_field {
    int L;                       /* 64-bit limbs */
    uint64_t q[ORC_MAX_L];       /* modulus */
    uint64_t qinvneg;            /* -q^-1 mod 2^64 (element.go:72) */
    uint64_t rsquare[ORC_MAX_L]; /* R^2 mod q (element.go:784) */
    uint64_t one[ORC_MAX_L];     /* R mod q  */
} orc_field;

int  orc_field_init(orc_field *f, const uint64_t *q_le, int L);

/* element ops (element.go:397-466 Add/Sub/Neg, :481-602 Mul, :604-653 fromMont) */
void orc_add(const orc_field *f, uint64_t *z, const uint64_t *x, const uint64_t *y);
void orc_sub(const orc_field *f, uint64_t *z, const uint64_t *x, const uint64_t *y);
void orc_neg(const orc_field *f, uint64_t *z, const uint64_t *x);
void orc_mul(const orc_field *f, uint64_t *z, const uint64_t *x, const uint64_t *y);
void orc_to_mont(const orc_field *f, uint64_t *z, const uint64_t *x);
void orc_from_mont(const orc_field *f, uint64_t *z, const uint64_t *x);
void orc_set_u64(const orc_field *f, uint64_t *z, uint64_t v);
void orc_exp_limbs(const orc_field *f, uint64_t *z, const uint64_t *x, const uint64_t *e, int elimbs);
void orc_inverse(const orc_field *f, uint64_t *z, const uint64_t *x);

/* math/bigpoly/ntt.go */
typedef struct orc_ntt {
    const orc_field *f;
    int logN;
    int negacyclic;
    size_t N;
    uint64_t *tw;      /* N elements (tw[0] unused for cyclic) */
    uint64_t *twinv;   /* N elements */
    uint64_t rankinv[ORC_MAX_L];
    uint64_t gen_x;    /* the small integer x the generator search stopped at */
} orc_ntt;

int  orc_ntt_init(orc_ntt *t, const orc_field *f, int logN, int negacyclic);
void orc_ntt_free(orc_ntt *t);
void orc_ntt_fwd(const orc_ntt *t, uint64_t *v_out, const uint64_t *v);   /* FwdNTTTo */
void orc_ntt_inv(const orc_ntt *t, uint64_t *v_out, const uint64_t *v);   /* InvNTTTo */
/* batch helpers used by the CPU baseline (OpenMP over polys when threads > 1) */
void orc_ntt_fwd_batch(const orc_ntt *t, uint64_t *v, size_t n_polys, int threads);
void orc_ntt_inv_batch(const orc_ntt *t, uint64_t *v, size_t n_polys, int threads);

/* math/bigpoly/vec.go */
void orc_vec_add(const orc_field *f, uint64_t *o, const uint64_t *a, const uint64_t *b, size_t n);
void orc_vec_sub(const orc_field *f, uint64_t *o, const uint64_t *a, const uint64_t *b, size_t n);
void orc_vec_neg(const orc_field *f, uint64_t *o, const uint64_t *a, size_t n);
void orc_vec_mul(const orc_field *f, uint64_t *o, const uint64_t *a, const uint64_t *b, size_t n);
void orc_vec_scalar_mul(const orc_field *f, uint64_t *o, const uint64_t *a, const uint64_t *c, size_t n);
void orc_bit_reverse(const orc_field *f, uint64_t *v, size_t n);

/* math/bigpoly/cyclic.go:18-37 */
void orc_quorem_vanishing(const orc_field *f, uint64_t *quo, uint64_t *rem, const uint64_t *p, size_t rank, size_t N);
/* math/bigpoly/cyclotomic.go:49-82 */
void orc_aut(const orc_field *f, uint64_t *o, const uint64_t *p, size_t rank, long idx, int is_ntt);
/* math/bigpoly/poly.go:64-76 */
void orc_poly_evaluate(const orc_field *f, uint64_t *z, const uint64_t *coeffs, size_t n, const uint64_t *x);

/* buckler/prover.go:355-379 evalCircuit restated over flat term lists (see include/ringo_cuda.h rs_term) */
typedef struct orc_term {
    int32_t constraint;   /* index of the constraint this term belongs to */
    int32_t coeff;        /* index into coeffs[] (Montgomery scalars) */
    int32_t pw;           /* public witness poly index or -1 */
    int32_t n_w;          /* number of witness factors */
    int32_t w[8];         /* witness poly indices */
} orc_term;
void orc_eval_circuit(const orc_field *f, uint64_t *out, size_t rank,
                      const uint64_t *const *pw_ntt, const uint64_t *const *w_ntt,
                      const orc_term *terms, size_t n_terms, const uint64_t *coeffs,
                      const uint64_t *batch_const);

#ifdef __cplusplus
}
#endif
#endif
