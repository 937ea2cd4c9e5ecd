/*
 * oracle/field_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see field_oracle.h).
 *
 * Restates, in plain C:
 *   - gnark-crypto generated Montgomery field arithmetic
 *     (reference: jindo/internal/zp/element.go:397-466, :481-602, :604-653;
 *      2-limb variant examples/mult/zp/element_purego.go:40-119)
 *   - math/bigpoly/ntt.go (twiddle generation :26-95, :153-203; radix-2
 *     butterflies :254-275, :365-386), vec.go, cyclic.go, cyclotomic.go, poly.go
 *   - buckler/prover.go:355-379 (evalCircuit)
 *
 * Everything is value-level identical to the reference: elements are always
 * canonical Montgomery residues in [0, q).
 */
#include "field_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;

/* ------------------------------------------------------------------------- */
/* multi-limb helpers                                                        */
/* ------------------------------------------------------------------------- */

static inline int limbs_geq(const uint64_t *a, const uint64_t *b, int L) {
    for (int i = L - 1; i >= 0; i--) {
        if (a[i] > b[i]) return 1;
        if (a[i] < b[i]) return 0;
    }
    return 1;
}

static inline uint64_t limbs_add(uint64_t *z, const uint64_t *a, const uint64_t *b, int L) {
    uint64_t c = 0;
    for (int i = 0; i < L; i++) {
        u128 s = (u128)a[i] + b[i] + c;
        z[i] = (uint64_t)s;
        c = (uint64_t)(s >> 64);
    }
    return c;
}

static inline uint64_t limbs_sub(uint64_t *z, const uint64_t *a, const uint64_t *b, int L) {
    uint64_t br = 0;
    for (int i = 0; i < L; i++) {
        u128 d = (u128)a[i] - b[i] - br;
        z[i] = (uint64_t)d;
        br = (uint64_t)(d >> 64) & 1;
    }
    return br;
}

/* z = x + y mod q   (element.go:397-412; carry-out handled as in
 * examples/mult/zp/element.go:372-395 for moduli with a full top limb) */
void orc_add(const orc_field *f, uint64_t *z, const uint64_t *x, const uint64_t *y) {
    uint64_t t[ORC_MAX_L];
    uint64_t c = limbs_add(t, x, y, f->L);
    if (c || limbs_geq(t, f->q, f->L)) limbs_sub(t, t, f->q, f->L);
    memcpy(z, t, sizeof(uint64_t) * f->L);
}

/* z = x - y mod q   (element.go:427-446) */
void orc_sub(const orc_field *f, uint64_t *z, const uint64_t *x, const uint64_t *y) {
    uint64_t t[ORC_MAX_L];
    uint64_t b = limbs_sub(t, x, y, f->L);
    if (b) limbs_add(t, t, f->q, f->L);
    memcpy(z, t, sizeof(uint64_t) * f->L);
}

/* z = -x mod q, with -0 = 0   (element.go:448-466) */
void orc_neg(const orc_field *f, uint64_t *z, const uint64_t *x) {
    int zero = 1;
    for (int i = 0; i < f->L; i++) zero &= (x[i] == 0);
    if (zero) { memset(z, 0, sizeof(uint64_t) * f->L); return; }
    uint64_t t[ORC_MAX_L];
    limbs_sub(t, f->q, x, f->L);
    memcpy(z, t, sizeof(uint64_t) * f->L);
}

/* Montgomery CIOS product z = x*y*R^-1 mod q  (element.go:481-602 _mulGeneric).
 * One extra word keeps the running sum for moduli without a spare bit. */
#define DEFINE_MUL(NAME, LL)                                                          \
    static inline void NAME(const orc_field *f, uint64_t *z, const uint64_t *x, const uint64_t *y) { \
        uint64_t t[(LL) + 2];                                                         \
        for (int i = 0; i < (LL) + 2; i++) t[i] = 0;                                  \
        const uint64_t *q = f->q;                                                     \
        const uint64_t qinv = f->qinvneg;                                             \
        for (int i = 0; i < (LL); i++) {                                              \
            uint64_t c = 0;                                                           \
            const uint64_t yi = y[i];                                                 \
            for (int j = 0; j < (LL); j++) {                                          \
                u128 s = (u128)x[j] * yi + t[j] + c;                                  \
                t[j] = (uint64_t)s;                                                   \
                c = (uint64_t)(s >> 64);                                              \
            }                                                                         \
            u128 s2 = (u128)t[LL] + c;                                                \
            t[LL] = (uint64_t)s2;                                                     \
            t[(LL) + 1] = (uint64_t)(s2 >> 64);                                       \
            const uint64_t m = t[0] * qinv;                                           \
            u128 s = (u128)m * q[0] + t[0];                                           \
            c = (uint64_t)(s >> 64);                                                  \
            for (int j = 1; j < (LL); j++) {                                          \
                s = (u128)m * q[j] + t[j] + c;                                        \
                t[j - 1] = (uint64_t)s;                                               \
                c = (uint64_t)(s >> 64);                                              \
            }                                                                         \
            s2 = (u128)t[LL] + c;                                                     \
            t[(LL) - 1] = (uint64_t)s2;                                               \
            t[LL] = t[(LL) + 1] + (uint64_t)(s2 >> 64);                               \
        }                                                                             \
        if (t[LL] || limbs_geq(t, q, (LL))) limbs_sub(t, t, q, (LL));                 \
        for (int i = 0; i < (LL); i++) z[i] = t[i];                                   \
    }

DEFINE_MUL(mul_2, 2)
DEFINE_MUL(mul_4, 4)
DEFINE_MUL(mul_7, 7)
DEFINE_MUL(mul_14, 14)

static void mul_generic(const orc_field *f, uint64_t *z, const uint64_t *x, const uint64_t *y) {
    const int L = f->L;
    uint64_t t[ORC_MAX_L + 2];
    for (int i = 0; i < L + 2; i++) t[i] = 0;
    for (int i = 0; i < L; i++) {
        uint64_t c = 0;
        for (int j = 0; j < L; j++) {
            u128 s = (u128)x[j] * y[i] + t[j] + c;
            t[j] = (uint64_t)s;
            c = (uint64_t)(s >> 64);
        }
        u128 s2 = (u128)t[L] + c;
        t[L] = (uint64_t)s2;
        t[L + 1] = (uint64_t)(s2 >> 64);
        const uint64_t m = t[0] * f->qinvneg;
        u128 s = (u128)m * f->q[0] + t[0];
        c = (uint64_t)(s >> 64);
        for (int j = 1; j < L; j++) {
            s = (u128)m * f->q[j] + t[j] + c;
            t[j - 1] = (uint64_t)s;
            c = (uint64_t)(s >> 64);
        }
        s2 = (u128)t[L] + c;
        t[L - 1] = (uint64_t)s2;
        t[L] = t[L + 1] + (uint64_t)(s2 >> 64);
    }
    if (t[L] || limbs_geq(t, f->q, L)) limbs_sub(t, t, f->q, L);
    for (int i = 0; i < L; i++) z[i] = t[i];
}

void orc_mul(const orc_field *f, uint64_t *z, const uint64_t *x, const uint64_t *y) {
    switch (f->L) {
        case 2: mul_2(f, z, x, y); return;
        case 4: mul_4(f, z, x, y); return;
        case 7: mul_7(f, z, x, y); return;
        case 14: mul_14(f, z, x, y); return;
        default: mul_generic(f, z, x, y); return;
    }
}

void orc_to_mont(const orc_field *f, uint64_t *z, const uint64_t *x) { orc_mul(f, z, x, f->rsquare); }

/* fromMont = multiply by 1 (element.go:604-653) */
void orc_from_mont(const orc_field *f, uint64_t *z, const uint64_t *x) {
    uint64_t one[ORC_MAX_L] = {1};
    orc_mul(f, z, x, one);
}

void orc_set_u64(const orc_field *f, uint64_t *z, uint64_t v) {
    uint64_t t[ORC_MAX_L] = {0};
    t[0] = v;
    if (f->L == 1 && v >= f->q[0]) t[0] = v % f->q[0];
    orc_to_mont(f, z, t);
}

void orc_exp_limbs(const orc_field *f, uint64_t *z, const uint64_t *x, const uint64_t *e, int elimbs) {
    uint64_t r[ORC_MAX_L], b[ORC_MAX_L];
    memcpy(r, f->one, sizeof(uint64_t) * f->L);
    memcpy(b, x, sizeof(uint64_t) * f->L);
    for (int i = 0; i < elimbs; i++) {
        uint64_t w = e[i];
        for (int k = 0; k < 64; k++) {
            if (w & 1) orc_mul(f, r, r, b);
            orc_mul(f, b, b, b);
            w >>= 1;
        }
    }
    memcpy(z, r, sizeof(uint64_t) * f->L);
}

/* value-level identical to element.go Inverse (x^(q-2); 0 -> 0) */
void orc_inverse(const orc_field *f, uint64_t *z, const uint64_t *x) {
    uint64_t e[ORC_MAX_L], two[ORC_MAX_L] = {2};
    limbs_sub(e, f->q, two, f->L);
    orc_exp_limbs(f, z, x, e, f->L);
}

int orc_field_init(orc_field *f, const uint64_t *q_le, int L) {
    if (L < 1 || L > ORC_MAX_L - 2 || !(q_le[0] & 1)) return -1;
    memset(f, 0, sizeof(*f));
    f->L = L;
    memcpy(f->q, q_le, sizeof(uint64_t) * L);
    /* qInvNeg by Newton iteration on 2-adic inverse */
    uint64_t inv = 1;
    for (int i = 0; i < 6; i++) inv *= 2 - q_le[0] * inv;
    f->qinvneg = (uint64_t)0 - inv;
    /* R mod q and R^2 mod q by repeated doubling of 1 */
    uint64_t t[ORC_MAX_L] = {0};
    t[0] = 1;
    for (int i = 0; i < 128 * L; i++) {
        uint64_t c = limbs_add(t, t, t, L);
        if (c || limbs_geq(t, f->q, L)) limbs_sub(t, t, f->q, L);
        if (i == 64 * L - 1) memcpy(f->one, t, sizeof(uint64_t) * L);
    }
    memcpy(f->rsquare, t, sizeof(uint64_t) * L);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* math/bigpoly/vec.go                                                       */
/* ------------------------------------------------------------------------- */

void orc_vec_add(const orc_field *f, uint64_t *o, const uint64_t *a, const uint64_t *b, size_t n) {
    const int L = f->L;
    for (size_t i = 0; i < n; i++) orc_add(f, o + i * L, a + i * L, b + i * L);
}
void orc_vec_sub(const orc_field *f, uint64_t *o, const uint64_t *a, const uint64_t *b, size_t n) {
    const int L = f->L;
    for (size_t i = 0; i < n; i++) orc_sub(f, o + i * L, a + i * L, b + i * L);
}
void orc_vec_neg(const orc_field *f, uint64_t *o, const uint64_t *a, size_t n) {
    const int L = f->L;
    for (size_t i = 0; i < n; i++) orc_neg(f, o + i * L, a + i * L);
}
void orc_vec_mul(const orc_field *f, uint64_t *o, const uint64_t *a, const uint64_t *b, size_t n) {
    const int L = f->L;
    for (size_t i = 0; i < n; i++) orc_mul(f, o + i * L, a + i * L, b + i * L);
}
void orc_vec_scalar_mul(const orc_field *f, uint64_t *o, const uint64_t *a, const uint64_t *c, size_t n) {
    const int L = f->L;
    for (size_t i = 0; i < n; i++) orc_mul(f, o + i * L, a + i * L, c);
}

/* vec.go:123-136 bitReverseInPlace */
void orc_bit_reverse(const orc_field *f, uint64_t *v, size_t n) {
    const int L = f->L;
    uint64_t tmp[ORC_MAX_L];
    size_t j = 0;
    for (size_t i = 1; i < n; i++) {
        size_t bit = n >> 1;
        while (j >= bit) { j -= bit; bit >>= 1; }
        j += bit;
        if (i < j) {
            memcpy(tmp, v + i * L, 8 * L);
            memcpy(v + i * L, v + j * L, 8 * L);
            memcpy(v + j * L, tmp, 8 * L);
        }
    }
}

/* ------------------------------------------------------------------------- */
/* math/bigpoly/ntt.go                                                       */
/* ------------------------------------------------------------------------- */

/* (q-1)/d for a small power-of-two d, and the generator search of
 * ntt.go:40-53 / :167-180: smallest x >= 2 with (x^((q-1)/order))^(order/2) != 1 */
static void find_generator(const orc_field *f, uint64_t order, uint64_t *g_out, uint64_t *x_out) {
    const int L = f->L;
    uint64_t e[ORC_MAX_L], qm1[ORC_MAX_L], one_[ORC_MAX_L] = {1};
    limbs_sub(qm1, f->q, one_, L);
    /* e = (q-1) / order */
    int sh = __builtin_ctzll(order);
    for (int i = 0; i < L; i++) {
        uint64_t lo = qm1[i] >> sh;
        uint64_t hi = (i + 1 < L && sh) ? (qm1[i + 1] << (64 - sh)) : 0;
        e[i] = lo | hi;
    }
    uint64_t half[1] = {order >> 1};
    for (uint64_t x = 2;; x++) {
        uint64_t xm[ORC_MAX_L], g[ORC_MAX_L], gp[ORC_MAX_L];
        orc_set_u64(f, xm, x);
        orc_exp_limbs(f, g, xm, e, L);
        orc_exp_limbs(f, gp, g, half, 1);
        if (memcmp(gp, f->one, 8 * L) != 0) {
            memcpy(g_out, g, 8 * L);
            *x_out = x;
            return;
        }
    }
}

int orc_ntt_init(orc_ntt *t, const orc_field *f, int logN, int negacyclic) {
    memset(t, 0, sizeof(*t));
    const int L = f->L;
    const size_t N = (size_t)1 << logN;
    t->f = f; t->logN = logN; t->negacyclic = negacyclic; t->N = N;
    /* (q-1) mod 2N == 0  (ntt.go:35-37, :162-164) */
    if (((f->q[0] - 1) & (2 * N - 1)) != 0) return -1;
    t->tw = (uint64_t *)calloc(N * L, 8);
    t->twinv = (uint64_t *)calloc(N * L, 8);
    if (!t->tw || !t->twinv) return -2;
    uint64_t g[ORC_MAX_L], ginv[ORC_MAX_L];
    if (negacyclic) {
        /* ntt.go:166-192 */
        find_generator(f, 2 * N, g, &t->gen_x);
        orc_inverse(f, ginv, g);
        memcpy(t->tw, f->one, 8 * L);
        memcpy(t->twinv, f->one, 8 * L);
        for (size_t i = 1; i < N; i++) {
            orc_mul(f, t->tw + i * L, t->tw + (i - 1) * L, g);
            orc_mul(f, t->twinv + i * L, t->twinv + (i - 1) * L, ginv);
        }
        orc_bit_reverse(f, t->tw, N);
        orc_bit_reverse(f, t->twinv, N);
    } else {
        /* ntt.go:39-84 */
        if (N >= 2) {
            find_generator(f, N, g, &t->gen_x);
            orc_inverse(f, ginv, g);
            size_t H = N / 2;
            uint64_t *ref = (uint64_t *)calloc(H * L, 8), *refinv = (uint64_t *)calloc(H * L, 8);
            memcpy(ref, f->one, 8 * L);
            memcpy(refinv, f->one, 8 * L);
            for (size_t i = 1; i < H; i++) {
                orc_mul(f, ref + i * L, ref + (i - 1) * L, g);
                orc_mul(f, refinv + i * L, refinv + (i - 1) * L, ginv);
            }
            orc_bit_reverse(f, ref, H);
            orc_bit_reverse(f, refinv, H);
            for (size_t m = 1; m <= H; m <<= 1)
                for (size_t i = 0; i < m; i++) {
                    memcpy(t->tw + (m + i) * L, ref + i * L, 8 * L);
                    memcpy(t->twinv + (m + i) * L, refinv + i * L, 8 * L);
                }
            free(ref); free(refinv);
        }
    }
    uint64_t n_m[ORC_MAX_L];
    orc_set_u64(f, n_m, (uint64_t)N);
    orc_inverse(f, t->rankinv, n_m);
    return 0;
}

void orc_ntt_free(orc_ntt *t) {
    free(t->tw); free(t->twinv);
    t->tw = t->twinv = NULL;
}

/* ntt.go:254-259 */
static inline void butterfly(const orc_field *f, uint64_t *u, uint64_t *v, const uint64_t *w) {
    orc_mul(f, v, v, w);
    orc_add(f, u, u, v);
    uint64_t d[ORC_MAX_L];
    orc_add(f, d, v, v);
    orc_sub(f, v, u, d);
}
/* ntt.go:365-370 */
static inline void inv_butterfly(const orc_field *f, uint64_t *u, uint64_t *v, const uint64_t *w) {
    orc_add(f, u, u, v);
    uint64_t d[ORC_MAX_L];
    orc_add(f, d, v, v);
    orc_sub(f, v, u, d);
    orc_mul(f, v, v, w);
}

/* Fixed-width copies of the two transforms for the limb counts the reference's moduli use: same
 * butterfly sequence, with the field ops inlined and unrolled so the timed CPU baseline is not handicapped
 * by the generic-L loops (the reference's Element code is generated per modulus with fixed limb counts). */
#define DEFINE_NTT(LL)                                                                                    \
    static inline void add_##LL(const uint64_t *q, uint64_t *z, const uint64_t *x, const uint64_t *y) {   \
        uint64_t t[LL], c = 0;                                                                            \
        for (int i = 0; i < (LL); i++) { u128 s = (u128)x[i] + y[i] + c; t[i] = (uint64_t)s; c = (uint64_t)(s >> 64); } \
        if (c || limbs_geq(t, q, (LL))) limbs_sub(t, t, q, (LL));                                         \
        for (int i = 0; i < (LL); i++) z[i] = t[i];                                                       \
    }                                                                                                     \
    static inline void sub_##LL(const uint64_t *q, uint64_t *z, const uint64_t *x, const uint64_t *y) {   \
        uint64_t t[LL];                                                                                   \
        uint64_t b = limbs_sub(t, x, y, (LL));                                                            \
        if (b) limbs_add(t, t, q, (LL));                                                                  \
        for (int i = 0; i < (LL); i++) z[i] = t[i];                                                       \
    }                                                                                                     \
    static void ntt_in_place_##LL(const orc_ntt *t, uint64_t *p) {                                        \
        const orc_field *f = t->f;                                                                        \
        const uint64_t *q = f->q;                                                                         \
        const size_t N = t->N;                                                                            \
        size_t tt = N;                                                                                    \
        for (size_t m = 1; m <= N / 2; m <<= 1) {                                                         \
            tt >>= 1;                                                                                     \
            for (size_t i = 0; i < m; i++) {                                                              \
                size_t j1 = (i * tt) << 1, j2 = j1 + tt;                                                  \
                const uint64_t *w = t->tw + (m + i) * (LL);                                               \
                for (size_t j = j1; j < j2; j++) {                                                        \
                    uint64_t *u = p + j * (LL), *v = p + (j + tt) * (LL), d[LL];                          \
                    mul_##LL(f, v, v, w);                                                                 \
                    add_##LL(q, u, u, v);                                                                 \
                    add_##LL(q, d, v, v);                                                                 \
                    sub_##LL(q, v, u, d);                                                                 \
                }                                                                                         \
            }                                                                                             \
        }                                                                                                 \
    }                                                                                                     \
    static void intt_in_place_##LL(const orc_ntt *t, uint64_t *p) {                                       \
        const orc_field *f = t->f;                                                                        \
        const uint64_t *q = f->q;                                                                         \
        const size_t N = t->N;                                                                            \
        size_t tt = 1;                                                                                    \
        for (size_t m = N / 2; m >= 1; m >>= 1) {                                                         \
            for (size_t i = 0; i < m; i++) {                                                              \
                size_t j1 = (i * tt) << 1, j2 = j1 + tt;                                                  \
                const uint64_t *w = t->twinv + (m + i) * (LL);                                            \
                for (size_t j = j1; j < j2; j++) {                                                        \
                    uint64_t *u = p + j * (LL), *v = p + (j + tt) * (LL), d[LL];                          \
                    add_##LL(q, u, u, v);                                                                 \
                    add_##LL(q, d, v, v);                                                                 \
                    sub_##LL(q, v, u, d);                                                                 \
                    mul_##LL(f, v, v, w);                                                                 \
                }                                                                                         \
            }                                                                                             \
            tt <<= 1;                                                                                     \
        }                                                                                                 \
    }

DEFINE_NTT(2)
DEFINE_NTT(4)
DEFINE_NTT(7)
DEFINE_NTT(14)

/* ntt.go:261-275 nttInPlaceRef (the unrolled variant computes the same values) */
static void ntt_in_place(const orc_ntt *t, uint64_t *p) {
    const orc_field *f = t->f;
    const int L = f->L;
    switch (L) {
        case 2: ntt_in_place_2(t, p); return;
        case 4: ntt_in_place_4(t, p); return;
        case 7: ntt_in_place_7(t, p); return;
        case 14: ntt_in_place_14(t, p); return;
        default: break;
    }
    const size_t N = t->N;
    size_t tt = N;
    for (size_t m = 1; m <= N / 2; m <<= 1) {
        tt >>= 1;
        for (size_t i = 0; i < m; i++) {
            size_t j1 = (i * tt) << 1, j2 = j1 + tt;
            const uint64_t *w = t->tw + (m + i) * L;
            for (size_t j = j1; j < j2; j++) butterfly(f, p + j * L, p + (j + tt) * L, w);
        }
    }
}

/* ntt.go:372-386 inttInPlaceRef */
static void intt_in_place(const orc_ntt *t, uint64_t *p) {
    const orc_field *f = t->f;
    const int L = f->L;
    switch (L) {
        case 2: intt_in_place_2(t, p); return;
        case 4: intt_in_place_4(t, p); return;
        case 7: intt_in_place_7(t, p); return;
        case 14: intt_in_place_14(t, p); return;
        default: break;
    }
    const size_t N = t->N;
    size_t tt = 1;
    for (size_t m = N / 2; m >= 1; m >>= 1) {
        for (size_t i = 0; i < m; i++) {
            size_t j1 = (i * tt) << 1, j2 = j1 + tt;
            const uint64_t *w = t->twinv + (m + i) * L;
            for (size_t j = j1; j < j2; j++) inv_butterfly(f, p + j * L, p + (j + tt) * L, w);
        }
        tt <<= 1;
    }
}

void orc_ntt_fwd(const orc_ntt *t, uint64_t *v_out, const uint64_t *v) {
    if (v_out != v) memmove(v_out, v, t->N * t->f->L * 8);
    ntt_in_place(t, v_out);
}

void orc_ntt_inv(const orc_ntt *t, uint64_t *v_out, const uint64_t *v) {
    if (v_out != v) memmove(v_out, v, t->N * t->f->L * 8);
    intt_in_place(t, v_out);
    orc_vec_scalar_mul(t->f, v_out, v_out, t->rankinv, t->N);   /* ntt.go:134-135, :242-243 */
}

/* batch-parallel drivers for the CPU baseline: polys are independent, so they
 * are split statically over `threads` pthreads (the reference itself is
 * single-threaded; threads > 1 is "all the host cores" for bench.py). */
typedef struct { const orc_ntt *t; uint64_t *v; size_t p0, p1; int inverse; } ntt_job;

static void *ntt_job_run(void *arg) {
    ntt_job *j = (ntt_job *)arg;
    const size_t stride = j->t->N * j->t->f->L;
    for (size_t p = j->p0; p < j->p1; p++) {
        uint64_t *x = j->v + p * stride;
        if (j->inverse) {
            intt_in_place(j->t, x);
            orc_vec_scalar_mul(j->t->f, x, x, j->t->rankinv, j->t->N);
        } else {
            ntt_in_place(j->t, x);
        }
    }
    return NULL;
}

static void ntt_batch(const orc_ntt *t, uint64_t *v, size_t n_polys, int threads, int inverse) {
    if (threads < 1) threads = 1;
    if ((size_t)threads > n_polys) threads = (int)(n_polys ? n_polys : 1);
    pthread_t th[256];
    ntt_job jobs[256];
    if (threads > 256) threads = 256;
    for (int k = 0; k < threads; k++) {
        jobs[k] = (ntt_job){t, v, n_polys * k / threads, n_polys * (k + 1) / threads, inverse};
        if (k + 1 < threads) pthread_create(&th[k], NULL, ntt_job_run, &jobs[k]);
    }
    ntt_job_run(&jobs[threads - 1]);
    for (int k = 0; k + 1 < threads; k++) pthread_join(th[k], NULL);
}

void orc_ntt_fwd_batch(const orc_ntt *t, uint64_t *v, size_t n_polys, int threads) { ntt_batch(t, v, n_polys, threads, 0); }
void orc_ntt_inv_batch(const orc_ntt *t, uint64_t *v, size_t n_polys, int threads) { ntt_batch(t, v, n_polys, threads, 1); }

/* ------------------------------------------------------------------------- */
/* cyclic.go / cyclotomic.go / poly.go                                       */
/* ------------------------------------------------------------------------- */

/* cyclic.go:18-37 */
void orc_quorem_vanishing(const orc_field *f, uint64_t *quo, uint64_t *rem, const uint64_t *p, size_t rank, size_t N) {
    const int L = f->L;
    memset(quo, 0, rank * L * 8);
    if (rem != p) memmove(rem, p, rank * L * 8);
    for (size_t i = rank; i-- > N;) {
        orc_add(f, quo + (i - N) * L, quo + (i - N) * L, rem + i * L);
        orc_add(f, rem + (i - N) * L, rem + (i - N) * L, rem + i * L);
        memset(rem + i * L, 0, 8 * L);
    }
}

/* cyclotomic.go:29-82 (AutTo dispatch, autTo, autNTTTo) */
void orc_aut(const orc_field *f, uint64_t *o, const uint64_t *p, size_t rank, long idx, int is_ntt) {
    const int L = f->L;
    const long twoN = 2 * (long)rank;
    idx %= twoN;
    if (idx < 0) idx += twoN;
    uint64_t *buf = (uint64_t *)calloc(rank * L, 8);
    if (!is_ntt) {
        for (size_t i = 0; i < rank; i++) {
            size_t j = (size_t)(((unsigned long long)i * (unsigned long long)idx) % (unsigned long long)twoN);
            if (j < rank) memcpy(buf + j * L, p + i * L, 8 * L);
            else orc_neg(f, buf + (j - rank) * L, p + i * L);
        }
        memcpy(o, buf, rank * L * 8);
    } else {
        memcpy(buf, p, rank * L * 8);
        orc_bit_reverse(f, buf, rank);
        for (size_t i = 0; i < rank; i++) {
            size_t j = (size_t)(((unsigned long long)(2 * i + 1) * (unsigned long long)idx) % (unsigned long long)twoN);
            j = (j - 1) >> 1;
            memcpy(o + i * L, buf + j * L, 8 * L);
        }
        orc_bit_reverse(f, o, rank);
    }
    free(buf);
}

/* poly.go:64-76 Horner */
void orc_poly_evaluate(const orc_field *f, uint64_t *z, const uint64_t *coeffs, size_t n, const uint64_t *x) {
    uint64_t acc[ORC_MAX_L] = {0};
    for (size_t i = n; i-- > 0;) {
        orc_mul(f, acc, acc, x);
        orc_add(f, acc, acc, coeffs + i * f->L);
    }
    memcpy(z, acc, 8 * f->L);
}

/* buckler/prover.go:355-379.  Terms are grouped by constraint, in order. */
void orc_eval_circuit(const orc_field *f, uint64_t *out, size_t rank,
                      const uint64_t *const *pw_ntt, const uint64_t *const *w_ntt,
                      const orc_term *terms, size_t n_terms, const uint64_t *coeffs,
                      const uint64_t *batch_const) {
    const int L = f->L;
    for (size_t j = 0; j < rank; j++) {
        uint64_t acc[ORC_MAX_L] = {0}, eval[ORC_MAX_L] = {0}, term[ORC_MAX_L];
        int cur = n_terms ? terms[0].constraint : -1;
        for (size_t k = 0; k <= n_terms; k++) {
            if (k == n_terms || terms[k].constraint != cur) {
                orc_mul(f, eval, eval, batch_const);
                orc_add(f, acc, acc, eval);
                memset(eval, 0, sizeof(eval));
                if (k == n_terms) break;
                cur = terms[k].constraint;
            }
            const orc_term *t = &terms[k];
            memcpy(term, coeffs + (size_t)t->coeff * L, 8 * L);
            if (t->pw >= 0) orc_mul(f, term, term, pw_ntt[t->pw] + j * L);
            for (int w = 0; w < t->n_w; w++) orc_mul(f, term, term, w_ntt[t->w[w]] + j * L);
            orc_add(f, eval, eval, term);
        }
        memcpy(out + j * L, acc, 8 * L);
    }
}
