/* jindo_oracle.c -- plain-C restatement of jindo.Prover.Commit for the CPU baseline and mid-size parity.
 *
 * TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's CPU legs may use it.
 * It follows the reference's algorithm (jindo/prover.go:45-202, jindo/encoder.go:113-201,
 * jindo/rns.go:76-114, math/csprng/gaussian_twin_cdt.go:77-111) with the randomness passed in as the
 * same tape the CUDA path takes (first/last/mask row elements, TwinCDT uniform draws, sampled COSAC /
 * rounded-Gaussian integers), so both sides consume identical draws.  Ring arithmetic is the value-level
 * contract of lattigo's ring.Ring (tuneinsight/lattigo/v6 v6.1.0, not under /root/reference: parity
 * unpinned for bytes that depend on it, see oracle/__init__.py): NTT out[i] = a(psi^(2 brv(i)+1)),
 * Montgomery form x * 2^64 mod q.  Pinned against oracle/jindo_ref.py (tests/test_oracle_jindo_c.py),
 * which is itself checked against the reference's parameter tables and verifier equations.
 *
 * Limits: at most four RNS primes per ring with a product below 2^250, d = 256. */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "field_oracle.h"

typedef unsigned __int128 u128;
typedef __int128 i128;

#define ORJ_MAX_LIMBS 4
#define ORJ_MAX_EXP 64
#define ORJ_D 256

/* same layout as rs_jindo_params (include/ringo_jindo.h) so the tests fill one structure for both sides */
typedef struct orj_params {
    int32_t d, slots, exp;
    uint64_t base;
    int32_t rows, cols, in_msis, mlwe, out_msis, log_in_cut, log_out_cut, n_q, n_qout;
    uint64_t q[ORJ_MAX_LIMBS], qout[ORJ_MAX_LIMBS], psi_q[ORJ_MAX_LIMBS], psi_qout[ORJ_MAX_LIMBS];
    double delta_inv[ORJ_MAX_EXP];
    int32_t cdt_len;
    int64_t cdt_tail_lo;
    double cdt_sigma;
} orj_params;

/* ---- one RNS limb ----------------------------------------------------------------------- */
typedef struct {
    uint64_t q, r, r_inv, n_inv;          /* r = 2^64 mod q */
    uint64_t roots[ORJ_D], roots_inv[ORJ_D];
} subring;

static inline uint64_t mulmod(uint64_t a, uint64_t b, uint64_t q) { return (uint64_t)((u128)a * b % q); }
static uint64_t powmod(uint64_t a, uint64_t e, uint64_t q) {
    uint64_t r = 1;
    while (e) { if (e & 1) r = mulmod(r, a, q); a = mulmod(a, a, q); e >>= 1; }
    return r;
}
static unsigned brv8(unsigned i) { unsigned r = 0; for (int k = 0; k < 8; k++) { r = (r << 1) | (i & 1); i >>= 1; } return r; }

static void subring_init(subring *s, uint64_t q, uint64_t psi) {
    s->q = q;
    s->r = (uint64_t)((((u128)1) << 64) % q);
    s->r_inv = powmod(s->r, q - 2, q);
    s->n_inv = powmod(ORJ_D, q - 2, q);
    const uint64_t psi_inv = powmod(psi, q - 2, q);
    uint64_t p = 1, pi = 1;
    for (unsigned j = 0; j < ORJ_D; j++) {                       /* RootsForward[brv(j)] = psi^j */
        s->roots[brv8(j)] = p; s->roots_inv[brv8(j)] = pi;
        p = mulmod(p, psi, q); pi = mulmod(pi, psi_inv, q);
    }
}
/* [lattigo] NTT: Cooley-Tukey, canonical output */
static void sr_ntt(const subring *s, uint64_t *a) {
    const uint64_t q = s->q;
    int t = ORJ_D;
    for (int m = 1; m < ORJ_D; m <<= 1) {
        t >>= 1;
        for (int i = 0; i < m; i++) {
            const uint64_t w = s->roots[m + i];
            const int j1 = 2 * i * t;
            for (int j = j1; j < j1 + t; j++) {
                const uint64_t v = mulmod(a[j + t], w, q), u = a[j];
                a[j] = u + v >= q ? u + v - q : u + v;
                a[j + t] = u >= v ? u - v : u + q - v;
            }
        }
    }
}
/* [lattigo] INTT: Gentleman-Sande, includes N^-1 */
static void sr_intt(const subring *s, uint64_t *a) {
    const uint64_t q = s->q;
    int t = 1;
    for (int m = ORJ_D; m > 1; m >>= 1) {
        const int h = m >> 1;
        int j1 = 0;
        for (int i = 0; i < h; i++) {
            const uint64_t w = s->roots_inv[h + i];
            for (int j = j1; j < j1 + t; j++) {
                const uint64_t u = a[j], v = a[j + t];
                a[j] = u + v >= q ? u + v - q : u + v;
                a[j + t] = mulmod(u >= v ? u - v : u + q - v, w, q);
            }
            j1 += 2 * t;
        }
        t <<= 1;
    }
    for (int j = 0; j < ORJ_D; j++) a[j] = mulmod(a[j], s->n_inv, q);
}


/* 256-bit two's-complement integers, little-endian words: enough for the CRT of up to four 62-bit primes */
typedef struct { uint64_t w[4]; } i256;

static i256 i256_from_i64(int64_t v) { i256 r; r.w[0] = (uint64_t)v; r.w[1] = r.w[2] = r.w[3] = v < 0 ? ~0ull : 0; return r; }
static void i256_mul_add_small(i256 *x, uint64_t m, uint64_t a) {         /* x = x*m + a (unsigned) */
    u128 c = a;
    for (int k = 0; k < 4; k++) { c += (u128)x->w[k] * m; x->w[k] = (uint64_t)c; c >>= 64; }
}
static int i256_ucmp(const i256 *a, const i256 *b) {
    for (int k = 3; k >= 0; k--) if (a->w[k] != b->w[k]) return a->w[k] < b->w[k] ? -1 : 1;
    return 0;
}
static void i256_sub(i256 *a, const i256 *b) {
    uint64_t br = 0;
    for (int k = 0; k < 4; k++) { const u128 d = (u128)a->w[k] - b->w[k] - br; a->w[k] = (uint64_t)d; br = (uint64_t)(d >> 64) & 1; }
}
static void i256_neg(i256 *a) {
    uint64_t c = 1;
    for (int k = 0; k < 4; k++) { const u128 s = (u128)(~a->w[k]) + c; a->w[k] = (uint64_t)s; c = (uint64_t)(s >> 64); }
}
static void i256_sar(i256 *a, int sh) {                                   /* arithmetic shift right == big.Int.Rsh (floor) */
    const uint64_t fill = (a->w[3] >> 63) ? ~0ull : 0;
    const int ws = sh / 64, bs = sh % 64;
    uint64_t t[4];
    for (int k = 0; k < 4; k++) {
        const uint64_t lo = k + ws < 4 ? a->w[k + ws] : fill, hi = k + ws + 1 < 4 ? a->w[k + ws + 1] : fill;
        t[k] = bs ? (lo >> bs) | (hi << (64 - bs)) : lo;
    }
    memcpy(a->w, t, sizeof(t));
}
static uint64_t i256_umod_small(const i256 *a, uint64_t q) {
    u128 r = 0;
    for (int k = 3; k >= 0; k--) r = ((r << 64) | a->w[k]) % q;
    return (uint64_t)r;
}
/* Euclidean residue of a signed value (rns.go:108-114) */
static uint64_t i256_mod_small(const i256 *a, uint64_t q) {
    if (!(a->w[3] >> 63)) return i256_umod_small(a, q);
    i256 n = *a; i256_neg(&n);
    const uint64_t m = i256_umod_small(&n, q);
    return m ? q - m : 0;
}

typedef struct {
    int n;
    uint64_t q[ORJ_MAX_LIMBS];
    uint64_t inv[ORJ_MAX_LIMBS][ORJ_MAX_LIMBS];          /* inv[i][j] = q_j^-1 mod q_i, j < i (Garner) */
    i256 Q, Qhalf;
} crt_t;

typedef struct {
    orj_params p;
    const orc_field *f;
    subring rq[ORJ_MAX_LIMBS], rqo[ORJ_MAX_LIMBS];
    crt_t crt_q, crt_qo;                                /* CRT of ringQ / ringQOut (rns.go:30-58) */
    const uint64_t *ck_in, *ck_mlwe, *ck_out, *cdt;
    const uint64_t *v; size_t len;
    const uint64_t *first, *last, *mask; const uint64_t *twin_u; const int64_t *cosac_g, *mlwe_s;
    uint64_t *encode, *mlwe, *incommit, *commitment;
    /* per-(col,row) tape indices, precomputed in commitColTo order */
    int32_t *twin_idx, *cosac_idx, *mask_idx;
} ctx_t;

static int crt_setup(crt_t *c, const subring *r, int n) {
    c->n = n;
    memset(&c->Q, 0, sizeof(c->Q)); c->Q.w[0] = 1;
    int bits = 0;
    for (int i = 0; i < n; i++) {
        c->q[i] = r[i].q;
        bits += 64 - __builtin_clzll(r[i].q);
        i256_mul_add_small(&c->Q, r[i].q, 0);
        for (int j = 0; j < i; j++) c->inv[i][j] = powmod(r[j].q % r[i].q, r[i].q - 2, r[i].q);
    }
    if (bits > 250) return -1;
    c->Qhalf = c->Q;
    for (int k = 0; k < 4; k++) c->Qhalf.w[k] = (c->Q.w[k] >> 1) | (k < 3 ? c->Q.w[k + 1] << 63 : 0);
    return 0;
}

/* rns.go:94-105: the residue vector as a centred integer.  The reference sums res_j * gad_j mod Q; the value in
 * [0, Q) is unique, so it is computed here by Garner's mixed-radix form, then centred with the same
 * "acc >= Q >> 1" rule.  (The caller handles the fast path of rns.go:80-92.) */
static i256 reconstruct(const crt_t *c, const uint64_t *res) {
    uint64_t x[ORJ_MAX_LIMBS];
    for (int i = 0; i < c->n; i++) {
        const uint64_t qi = c->q[i];
        uint64_t t = res[i] % qi;
        for (int j = 0; j < i; j++) {
            const uint64_t xj = x[j] % qi;
            t = mulmod(t >= xj ? t - xj : t + qi - xj, c->inv[i][j], qi);
        }
        x[i] = t;
    }
    i256 acc; memset(&acc, 0, sizeof(acc));
    acc.w[0] = x[c->n - 1];
    for (int i = c->n - 2; i >= 0; i--) i256_mul_add_small(&acc, c->q[i], x[i]);
    if (i256_ucmp(&acc, &c->Qhalf) >= 0) i256_sub(&acc, &c->Q);
    return acc;
}

/* rns.go:68-73 */
static inline int64_t balanced(uint64_t x, uint64_t q) { return x > (q >> 1) ? (int64_t)x - (int64_t)q : (int64_t)x; }

/* prover.go:164-176 / :190-200: IMForm, INTT, CRT, >> cut (floor), re-RNS, MForm, NTT */
static void cut_pipeline(const ctx_t *c, int outer, const uint64_t *poly_in, uint64_t *poly_out) {
    const subring *ri = outer ? c->rqo : c->rq;
    const crt_t *crt = outer ? &c->crt_qo : &c->crt_q;
    const int ni = crt->n;
    const int cut = outer ? c->p.log_out_cut : c->p.log_in_cut;
    uint64_t tmp[ORJ_MAX_LIMBS][ORJ_D];
    for (int l = 0; l < ni; l++) {
        for (int j = 0; j < ORJ_D; j++) tmp[l][j] = mulmod(poly_in[l * ORJ_D + j], ri[l].r_inv, ri[l].q);
        sr_intt(&ri[l], tmp[l]);
    }
    static _Thread_local i256 big[ORJ_D];
    for (int j = 0; j < ORJ_D; j++) {
        uint64_t res[ORJ_MAX_LIMBS] = {0, 0, 0, 0};
        for (int l = 0; l < ni; l++) res[l] = tmp[l][j];
        /* fast path of rns.go:80-92 */
        const int64_t c0 = balanced(res[0], ri[0].q);
        int agree = 1;
        for (int l = 1; l < ni; l++) if (balanced(res[l], ri[l].q) != c0) { agree = 0; break; }
        big[j] = agree ? i256_from_i64(c0) : reconstruct(crt, res);
        i256_sar(&big[j], cut);
    }
    const int no = c->p.n_qout;
    for (int l = 0; l < no; l++) {
        const uint64_t q = c->rqo[l].q;
        uint64_t *o = poly_out + l * ORJ_D;
        for (int j = 0; j < ORJ_D; j++) o[j] = mulmod(i256_mod_small(&big[j], q), c->rqo[l].r, q);
        sr_ntt(&c->rqo[l], o);
    }
}

/* gaussian_twin_cdt.go:77-111 with the uniform draw passed in */
static int64_t twin_sample(const ctx_t *c, double center, uint64_t u) {
    const int len = c->p.cdt_len;
    const double sigma = c->p.cdt_sigma;
    const double cfloor = floor(center), cfrac = center - cfloor;
    const int c0 = (int)(((int64_t)floor(128.0 * cfrac)) % 128), c1 = (int)(((int64_t)ceil(128.0 * cfrac)) % 128);
    int v[2];
    const int cc[2] = {c0, c1};
    for (int t = 0; t < 2; t++) {
        const uint64_t *tab = c->cdt + (size_t)cc[t] * len;
        int lo = 0, hi = len;                             /* slices.BinarySearch: first index with tab[i] >= u */
        while (lo < hi) { const int m = (lo + hi) >> 1; if (tab[m] < u) lo = m + 1; else hi = m; }
        if (lo < len && tab[lo] == u) lo -= 1;
        v[t] = lo;
    }
    if (v[0] == v[1]) return (int64_t)v[0] + (int64_t)cfloor + c->p.cdt_tail_lo;
    double cdf = 0.0;
    const double norm = sqrt(2 * M_PI) * sigma;
    for (int64_t x = c->p.cdt_tail_lo; x <= (int64_t)v[0]; x++) {
        const double xf = (double)x;
        cdf += exp(-(xf - cfrac) * (xf - cfrac) / (2 * sigma * sigma)) / norm;
    }
    const double p = (double)u / 18446744073709551616.0;
    return (int64_t)(p < cdf ? v[0] : v[1]) + c->p.cdt_tail_lo + (int64_t)cfloor;
}

/* encoder.go:149-201 randEncodeTo for `cnt` elements (Montgomery limbs at src); noise: TwinCDT draws (twin != NULL)
 * or given integers g */
static void rand_encode(const ctx_t *c, const uint64_t *src, int cnt, const uint64_t *twin, const int64_t *g_in, uint64_t *out) {
    const orj_params *p = &c->p;
    const int L = c->f->L, slots = p->slots, ex = p->exp;
    uint64_t digits[ORJ_D];
    memset(digits, 0, sizeof(digits));
    for (int i = 0; i < cnt; i++) {                       /* baseEncodeTo encoder.go:120-146 */
        uint64_t limbs[ORC_MAX_L];
        orc_from_mont(c->f, limbs, src + (size_t)i * L);
        for (int j = 0; j < ex - 1; j++) {                /* divMod64 by base, most significant limb first */
            u128 rem = 0;
            for (int k = L - 1; k >= 0; k--) {
                const u128 cur = (rem << 64) | limbs[k];
                limbs[k] = (uint64_t)(cur / p->base);
                rem = cur % p->base;
            }
            digits[j * slots + i] = (uint64_t)rem;
        }
        digits[(ex - 1) * slots + i] = limbs[0];
    }
    int64_t g[ORJ_D];
    if (twin) {
        double fp[ORJ_D];
        for (int j = 0; j < ORJ_D; j++) fp[j] = 0.0;
        for (int i = 0; i < ex; i++) {                    /* fpSample encoder.go:152-164 (unfused, this order) */
            const double di = p->delta_inv[i];
            if (di == 0.0) continue;
            const int dd = ORJ_D - (i + 1) * slots;
            for (int j = 0; j < ORJ_D - dd; j++) fp[j + dd] += di * (double)digits[j];
            for (int j = ORJ_D - dd, jj = 0; j < ORJ_D; j++, jj++) fp[jj] -= di * (double)digits[j];
        }
        for (int j = 0; j < ORJ_D; j++) g[j] = twin_sample(c, -fp[j], twin[j]);
    } else {
        for (int j = 0; j < ORJ_D; j++) g[j] = g_in[j];
    }
    for (int l = 0; l < p->n_q; l++) {
        const subring *s = &c->rq[l];
        const uint64_t q = s->q;
        uint64_t ps[ORJ_D], *o = out + (size_t)l * ORJ_D;
        for (int j = 0; j < ORJ_D; j++) {
            const int64_t gv = g[j];
            const uint64_t res = gv >= 0 ? (uint64_t)gv % q : (uint64_t)((gv % (int64_t)q) + (int64_t)q) % q;
            ps[j] = mulmod(res, s->r, q);
        }
        const uint64_t bq = p->base % q;
        for (int j = 0; j < ORJ_D; j++) {
            const uint64_t sh = j >= slots ? ps[j - slots] : (q - ps[j + ORJ_D - slots]) % q;   /* X^slots * g */
            const uint64_t bg = mulmod(ps[j], bq, q);
            const uint64_t t = sh >= bg ? sh - bg : sh + q - bg;
            const uint64_t dm = mulmod(digits[j] % q, s->r, q);
            o[j] = dm + t >= q ? dm + t - q : dm + t;
        }
        sr_ntt(s, o);
    }
}

static void noise_poly(const ctx_t *c, const int64_t *smp, uint64_t *out) {   /* prover.go:130-142 */
    for (int l = 0; l < c->p.n_q; l++) {
        const subring *s = &c->rq[l];
        uint64_t *o = out + (size_t)l * ORJ_D;
        for (int j = 0; j < ORJ_D; j++) {
            const int64_t gv = smp[j];
            const uint64_t res = gv >= 0 ? (uint64_t)gv % s->q : (uint64_t)((gv % (int64_t)s->q) + (int64_t)s->q) % s->q;
            o[j] = mulmod(res, s->r, s->q);
        }
        sr_ntt(s, o);
    }
}

/* commitColTo (prover.go:89-177) for column i */
static void commit_col(const ctx_t *c, int i) {
    const orj_params *p = &c->p;
    const int L = c->f->L, nq = p->n_q, rows = p->rows, cols = p->cols, slots = p->slots;
    const size_t pq = (size_t)nq * ORJ_D, cs = (size_t)cols * slots;
    uint64_t *enc = c->encode + (size_t)i * rows * pq;
    for (int r = 0; r < rows; r++) {
        uint64_t *o = enc + (size_t)r * pq;
        const int ti = c->twin_idx[i * rows + r], ci = c->cosac_idx[i * rows + r], mi = c->mask_idx[i * rows + r];
        if (ti < 0 && ci < 0) { memset(o, 0, pq * 8); continue; }           /* row past len(v) */
        const uint64_t *src; int cnt = slots;
        if (i == cols) src = c->mask + (size_t)mi * slots * L;
        else if (r == 0) src = c->first + (size_t)i * slots * L;
        else if (r == rows - 1) src = c->last + (size_t)i * slots * L;
        else {
            const size_t s0 = (size_t)r * cs + (size_t)i * slots, e0 = s0 + slots;
            src = c->v + s0 * L; cnt = (int)((e0 < c->len ? e0 : c->len) - s0);
        }
        if (ti >= 0) rand_encode(c, src, cnt, c->twin_u + (size_t)ti * ORJ_D, NULL, o);
        else rand_encode(c, src, cnt, NULL, c->cosac_g + (size_t)ci * ORJ_D, o);
    }
    const int nm = p->mlwe + p->in_msis;
    uint64_t *ml = c->mlwe + (size_t)i * nm * pq;
    for (int k = 0; k < nm; k++) noise_poly(c, c->mlwe_s + ((size_t)i * nm + k) * ORJ_D, ml + (size_t)k * pq);
    uint64_t com[ORJ_MAX_LIMBS * ORJ_D];
    for (int o = 0; o < p->in_msis; o++) {                                  /* prover.go:144-157 */
        for (int l = 0; l < nq; l++) {
            const uint64_t q = c->rq[l].q, ri = c->rq[l].r_inv;
            for (int j = 0; j < ORJ_D; j++) {
                uint64_t acc = 0;
                for (int k = 0; k < rows; k++) {
                    const uint64_t a = c->ck_in[((size_t)o * rows + k) * pq + l * ORJ_D + j], b = enc[(size_t)k * pq + l * ORJ_D + j];
                    const uint64_t t = mulmod(mulmod(a, b, q), ri, q);
                    acc = acc + t >= q ? acc + t - q : acc + t;
                }
                for (int k = 0; k < p->mlwe; k++) {
                    const uint64_t a = c->ck_mlwe[((size_t)o * p->mlwe + k) * pq + l * ORJ_D + j], b = ml[(size_t)k * pq + l * ORJ_D + j];
                    const uint64_t t = mulmod(mulmod(a, b, q), ri, q);
                    acc = acc + t >= q ? acc + t - q : acc + t;
                }
                const uint64_t e = ml[(size_t)(p->mlwe + o) * pq + l * ORJ_D + j];
                com[l * ORJ_D + j] = acc + e >= q ? acc + e - q : acc + e;
            }
        }
        cut_pipeline(c, 0, com, c->incommit + ((size_t)i * p->in_msis + o) * p->n_qout * ORJ_D);
    }
}

typedef struct { const ctx_t *c; int t, nt; } job_t;
static void *col_worker(void *arg) {
    const job_t *j = (const job_t *)arg;
    for (int i = j->t; i <= j->c->p.cols; i += j->nt) commit_col(j->c, i);
    return NULL;
}

/* Prover.Commit (prover.go:45-62, :180-202).  Outputs (caller allocated, lattigo layout [poly][limb][256]):
 *   encode  (cols+1)*rows polys over ringQ, index col*rows+row     mlwe  (cols+1)*(mlwe+in_msis) polys
 *   incommit (cols+1)*in_msis polys over ringQOut                  commitment  out_msis polys over ringQOut
 * Columns are independent and are split over `threads` pthreads (the reference is single-threaded). */
int orj_commit(const orc_field *f, const orj_params *p, const uint64_t *ck_in, const uint64_t *ck_mlwe, const uint64_t *ck_out,
               const uint64_t *cdt, const uint64_t *v_mont, size_t len,
               const uint64_t *first_row, const uint64_t *last_row, const uint64_t *mask_rows,
               const uint64_t *twin_u, const int64_t *cosac_g, const int64_t *mlwe_s,
               uint64_t *encode, uint64_t *mlwe, uint64_t *incommit, uint64_t *commitment, int threads) {
    if (p->d != ORJ_D || p->n_q > ORJ_MAX_LIMBS || p->n_qout > ORJ_MAX_LIMBS) return -1;
    ctx_t *c = (ctx_t *)calloc(1, sizeof(ctx_t));
    c->p = *p; c->f = f;
    for (int l = 0; l < p->n_q; l++) subring_init(&c->rq[l], p->q[l], p->psi_q[l]);
    for (int l = 0; l < p->n_qout; l++) subring_init(&c->rqo[l], p->qout[l], p->psi_qout[l]);
    if (crt_setup(&c->crt_q, c->rq, p->n_q) || crt_setup(&c->crt_qo, c->rqo, p->n_qout)) { free(c); return -2; }
    c->ck_in = ck_in; c->ck_mlwe = ck_mlwe; c->ck_out = ck_out; c->cdt = cdt; c->v = v_mont; c->len = len;
    c->first = first_row; c->last = last_row; c->mask = mask_rows; c->twin_u = twin_u; c->cosac_g = cosac_g; c->mlwe_s = mlwe_s;
    c->encode = encode; c->mlwe = mlwe; c->incommit = incommit; c->commitment = commitment;
    const int rows = p->rows, cols = p->cols, slots = p->slots;
    const size_t cs = (size_t)cols * slots, n = (size_t)(cols + 1) * rows;
    c->twin_idx = (int32_t *)malloc(n * 4); c->cosac_idx = (int32_t *)malloc(n * 4); c->mask_idx = (int32_t *)malloc(n * 4);
    for (size_t k = 0; k < n; k++) c->twin_idx[k] = c->cosac_idx[k] = c->mask_idx[k] = -1;
    int twin = 0, cosac = 0, mask = 0;
    /* consumption order of the tape == the reference's call order (prover.go:52-58, :89-128): the first rows
     * of all data columns are encoded first (gen + COSAC), then column by column */
    for (int i = 0; i < cols; i++) c->cosac_idx[i * rows + 0] = cosac++;
    for (int i = 0; i < cols; i++) {
        for (int r = 1; r < rows - 1; r++) {
            if ((size_t)r * cs + (size_t)i * slots > len) break;             /* note '>' (prover.go:121) */
            c->twin_idx[i * rows + r] = twin++;
        }
        c->twin_idx[i * rows + rows - 1] = twin++;
    }
    c->cosac_idx[cols * rows + 0] = cosac++; c->mask_idx[cols * rows + 0] = mask++;
    for (int r = 1; r < rows - 1; r++) {
        if ((size_t)r * cs > len) break;
        c->cosac_idx[cols * rows + r] = cosac++; c->mask_idx[cols * rows + r] = mask++;
    }
    c->cosac_idx[cols * rows + rows - 1] = cosac++; c->mask_idx[cols * rows + rows - 1] = mask++;

    if (threads < 1) threads = 1;
    if (threads > cols + 1) threads = cols + 1;
    pthread_t th[256]; job_t jobs[256];
    if (threads > 256) threads = 256;
    for (int t = 0; t < threads; t++) { jobs[t].c = c; jobs[t].t = t; jobs[t].nt = threads; pthread_create(&th[t], NULL, col_worker, &jobs[t]); }
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);

    /* outer commit (prover.go:180-202) */
    const int no = p->n_qout, nin = (cols + 1) * p->in_msis;
    const size_t pqo = (size_t)no * ORJ_D;
    uint64_t com[ORJ_MAX_LIMBS * ORJ_D];
    for (int o = 0; o < p->out_msis; o++) {
        for (int l = 0; l < no; l++) {
            const uint64_t q = c->rqo[l].q, ri = c->rqo[l].r_inv;
            for (int j = 0; j < ORJ_D; j++) {
                uint64_t acc = 0;
                for (int k = 0; k < nin; k++) {
                    const uint64_t a = ck_out[((size_t)o * nin + k) * pqo + l * ORJ_D + j], b = incommit[(size_t)k * pqo + l * ORJ_D + j];
                    const uint64_t t = mulmod(mulmod(a, b, q), ri, q);
                    acc = acc + t >= q ? acc + t - q : acc + t;
                }
                com[l * ORJ_D + j] = acc;
            }
        }
        cut_pipeline(c, 1, com, commitment + (size_t)o * pqo);
    }
    free(c->twin_idx); free(c->cosac_idx); free(c->mask_idx); free(c);
    return 0;
}
