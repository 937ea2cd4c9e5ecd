// shoup.cuh -- products by a FIXED multiplier with a precomputed quotient (Shoup), and predicated lazy add / sub,
// for the butterflies of the big-field NTT (math/bigpoly/ntt.go:254-259, :365-370).
//
// Every product of a butterfly has a twiddle factor as one operand, and the twiddles are table entries.  Let
// beta = 2^(32 NL), w the twiddle as a plain integer in [0, q) and w' = floor(w beta / q).  For any x < beta
//     h = floor(x w' / beta),     x w - h q  in  [0, (1 + x / beta) q)          (Shoup / Harvey)
// and only the LOW halves of x w and h q are needed (the difference is below beta).  With the data kept in Montgomery
// form (x = a R) the result is (a w) R again, i.e. exactly the residue class the reference's Montgomery product
// a R * w R * R^-1 lands in; the NTT output is canonicalised at the end, so every value the transform stores is the
// reference's value bit for bit.
//
// Cost for 8 limbs against the CIOS Montgomery product of field.cuh (64 + 48 wide multiply-adds for q == 1 mod 2^64):
//     low half of x w           36 limb products (8 of them low-word only)
//     high half of x w'         43 limb products (columns i + j >= NL - 2; see the error bound below)
//     low half of h q           21 limb products (q = 1 + 2^64 qh: h q = h + 2^64 (h qh), 6 low-word only)
// = 86 wide + 14 narrow multiply-adds instead of 112 wide + 8 narrow, with about the same number of additions.
//
// High half from a truncated product: with S = x w' and S_K the sum of the kept limb products (i + j >= NL - 2),
//     S_K <= S < S_K + (NL - 2) 2^(32 (NL - 1)) (1 + 2^-31),
// so floor(S_K / beta) == h unless limb NL-1 of S_K is within NL of 2^32 - 1.  That case (probability ~ NL 2^-32
// per product) takes a slow path: the quotient may be one short, the difference is then below (2 + x / beta) q and one
// conditional subtraction of q brings it back under (1 + x / beta) q.  No probabilistic claim is involved.
//
// The table holds (w, w') pairs: w is the Montgomery table entry e = w R mod q taken out of Montgomery form, and
//     w' = floor(w beta / q) = (w beta - e) / q = e * (-q^-1)  mod beta
// (the division is exact and w' < beta), one low-half product with the constant -q^-1 mod beta.
//
// Lazy ranges.  FORWARD transform: values live in [0, 2q + delta), delta = 2^(32 (NL - 2)): the sum x + t is reduced by
// comparing only its top 64 bits with those of 2q (a full comparison costs a second carry chain and eight selects), so a
// value may stay up to delta above 2q; the subtrahend of x - t + 2q is always a product (t < 1.56 q), so the difference is
// never negative, and the range does not grow from stage to stage.  Validity: 3.6 q < beta (x + t < beta: no carry word) and
// q >= 2^8 delta (rs_field::shoup_ok).  INVERSE transform (u' = u + v, v' = (u - v) w): the subtrahend is an arbitrary
// value, so sums need the exact reduction to [0, 2q), and the product comes last, with the difference live beside the tile
// registers; built and measured (zp255, 2^16 x 512: 4.58 against 4.59 ms), it keeps the Montgomery kernels.
#pragma once
#include "field.cuh"

namespace rs {

// ---------------------------------------------------------------------------
// triangular low product on two carry chains (see field.cuh for the even / odd limb arrangement):
//   (E + O 2^32)  (+)=  sum_{i + j < M} a[j] b[i] 2^(32 (i + j))      mod 2^(32 M)
// E[k] is limb k of the chain whose 64-bit products sit at even limbs, O[k] limb k + 1 of the odd chain.
// ADD0: E holds an addend on entry (row 0 adds into it); otherwise row 0 initialises E.  O is always initialised here.
// ---------------------------------------------------------------------------
template <int M, bool ADD0>
RS_HD void tri_lo(uint32_t *E, uint32_t *O, const uint32_t *a, const uint32_t *b) {
    CC c;
    if (ADD0) {
        bool first = true;
#pragma unroll
        for (int p = 0; p < M; p += 2) {
            if (first) c.mad_lo_cc(E[p], a[p], b[0], E[p]); else c.madc_lo_cc(E[p], a[p], b[0], E[p]);
            if (p < M - 1) c.madc_hi_cc(E[p + 1], a[p], b[0], E[p + 1]);
            first = false;
        }
    } else {
#pragma unroll
        for (int p = 0; p < M; p += 2) {
            E[p] = mul_lo(a[p], b[0]);
            if (p < M - 1) E[p + 1] = mul_hi(a[p], b[0]);
        }
    }
#pragma unroll
    for (int p = 1; p < M; p += 2) {
        O[p - 1] = mul_lo(a[p], b[0]);
        if (p < M - 1) O[p] = mul_hi(a[p], b[0]);
    }
#pragma unroll
    for (int i = 1; i < M; i++) {
        {
            bool first = true;
#pragma unroll
            for (int p = i + (i & 1); p < M; p += 2) {
                const int j = p - i;
                if (first) c.mad_lo_cc(E[p], a[j], b[i], E[p]); else c.madc_lo_cc(E[p], a[j], b[i], E[p]);
                if (p < M - 1) c.madc_hi_cc(E[p + 1], a[j], b[i], E[p + 1]);
                first = false;
            }
        }
        {
            bool first = true;
#pragma unroll
            for (int p = i + 1 - (i & 1); p < M; p += 2) {
                const int j = p - i;
                if (first) c.mad_lo_cc(O[p - 1], a[j], b[i], O[p - 1]); else c.madc_lo_cc(O[p - 1], a[j], b[i], O[p - 1]);
                if (p < M - 1) c.madc_hi_cc(O[p], a[j], b[i], O[p]);
                first = false;
            }
        }
    }
}

// r[0..M) = E + (O << 32)  mod 2^(32 M)
template <int M>
RS_HD void tri_merge(uint32_t *r, const uint32_t *E, const uint32_t *O) {
    CC c;
    r[0] = E[0];
    if (M > 1) {
        if (M == 2) { r[1] = E[1] + O[0]; return; }
        c.add_cc(r[1], E[1], O[0]);
#pragma unroll
        for (int k = 2; k < M - 1; k++) c.addc_cc(r[k], E[k], O[k - 1]);
        c.addc(r[M - 1], E[M - 1], O[M - 2]);
    }
}

// ---------------------------------------------------------------------------
// truncated high product: h = limbs NL .. 2NL-1 and frac = limb NL-1 of  S_K = sum_{i + j >= NL - 2} x[j] wp[i] 2^(32 (i + j))
// ---------------------------------------------------------------------------
template <int NL>
RS_HD void tri_hi(uint32_t *h, uint32_t &frac, const uint32_t *x, const uint32_t *wp) {
    static_assert(NL % 2 == 0 && NL >= 4, "NL must be even");
    uint32_t EH[NL + 2], OH[NL + 1];   // EH[k]: limb NL-2+k of the even chain, OH[k]: limb NL-1+k of the odd chain
#pragma unroll
    for (int k = 0; k < NL + 2; k++) EH[k] = 0;
#pragma unroll
    for (int k = 0; k < NL + 1; k++) OH[k] = 0;
    CC c;
#pragma unroll
    for (int i = 0; i < NL; i++) {
        const int jmin = (NL - 2 - i) > 0 ? (NL - 2 - i) : 0;
        {
            bool first = true;
            int top = 0, jtop = 0;
#pragma unroll
            for (int j = jmin; j < NL; j++) {
                if (((i + j) & 1) == 0) {
                    const int k = i + j - (NL - 2);
                    if (first) c.mad_lo_cc(EH[k], x[j], wp[i], EH[k]); else c.madc_lo_cc(EH[k], x[j], wp[i], EH[k]);
                    c.madc_hi_cc(EH[k + 1], x[j], wp[i], EH[k + 1]);
                    first = false;
                    top = k + 1; jtop = j;
                }
            }
            // A chain that ends with x[NL-1] writes its top limb for the first time (high word + carry cannot overflow): no
            // carry out.  The other chain ends on a limb the previous row filled: its carry lands one limb higher, in a limb
            // that so far holds nothing (the next row's x[NL-1] product is the first to add into it).
            if (i > 0 && jtop != NL - 1 && top + 1 < NL + 2) c.addc(EH[top + 1], EH[top + 1], 0);
        }
        {
            bool first = true;
            int top = 0, jtop = 0;
#pragma unroll
            for (int j = jmin; j < NL; j++) {
                if (((i + j) & 1) == 1) {
                    const int k = i + j - (NL - 1);
                    if (first) c.mad_lo_cc(OH[k], x[j], wp[i], OH[k]); else c.madc_lo_cc(OH[k], x[j], wp[i], OH[k]);
                    c.madc_hi_cc(OH[k + 1], x[j], wp[i], OH[k + 1]);
                    first = false;
                    top = k + 1; jtop = j;
                }
            }
            if (i > 0 && jtop != NL - 1 && top + 1 < NL + 1) c.addc(OH[top + 1], OH[top + 1], 0);
        }
    }
    c.add_cc(frac, EH[1], OH[0]);
#pragma unroll
    for (int k = 0; k < NL - 1; k++) c.addc_cc(h[k], EH[k + 2], OH[k + 1]);
    c.addc(h[NL - 1], EH[NL + 1], OH[NL]);
}

// z = x w - h q  with h = floor(x wp / beta) (or one less on the slow path, then corrected):
// x < beta, (w, wp) a table pair  ->  z == x w (mod q), z in [0, (1 + x / beta) q)
template <int NL, int QS>
RS_HD void shoup_mul_lazy(uint32_t *z, const uint32_t *x, const uint32_t *w, const uint32_t *wp, const FieldParams<NL> &F) {
    uint32_t h[NL], frac;
    tri_hi<NL>(h, frac, x, wp);
    // hq = h q mod beta
    uint32_t hq[NL];
    if (QS == 0) {
        uint32_t E[NL], O[NL];
        tri_lo<NL, false>(E, O, F.q, h);
        tri_merge<NL>(hq, E, O);
    } else {
        // q = 1 + 2^(32 QS) qh:  h q = h + 2^(32 QS) (h qh); the addend h rides in the even chain
        constexpr int M = NL - QS;
        uint32_t E[M], O[M];
#pragma unroll
        for (int k = 0; k < M; k++) E[k] = h[QS + k];
        tri_lo<M, true>(E, O, F.q + QS, h);
#pragma unroll
        for (int k = 0; k < QS; k++) hq[k] = h[k];
        tri_merge<M>(hq + QS, E, O);
    }
    uint32_t E[NL], O[NL], p[NL];
    tri_lo<NL, false>(E, O, x, w);
    tri_merge<NL>(p, E, O);
    CC c;
    c.sub_cc(z[0], p[0], hq[0]);
#pragma unroll
    for (int k = 1; k < NL - 1; k++) c.subc_cc(z[k], p[k], hq[k]);
    c.subc(z[NL - 1], p[NL - 1], hq[NL - 1]);
    if (frac > 0xffffffffu - (uint32_t)NL) cond_sub<NL>(z, F.q);   // slow path, ~NL 2^-32 of the products
}

// ---------------------------------------------------------------------------
// predicated multi-limb add / sub:  if (p) z (+-)= m   (wrap-around mod beta)
// On the device the carry chain is predicated instruction by instruction (no select, no mask).
// ---------------------------------------------------------------------------
template <int NL>
RS_HD void pred_sub(uint32_t *z, const uint32_t *m, bool p) {
#ifdef __CUDA_ARCH__
    static_assert(NL == 4 || NL == 8, "pred_sub: 4 or 8 limbs");
    if constexpr (NL == 8) {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %16, 0;\n\t"
                     "@p sub.cc.u32 %0, %0, %8;\n\t@p subc.cc.u32 %1, %1, %9;\n\t@p subc.cc.u32 %2, %2, %10;\n\t"
                     "@p subc.cc.u32 %3, %3, %11;\n\t@p subc.cc.u32 %4, %4, %12;\n\t@p subc.cc.u32 %5, %5, %13;\n\t"
                     "@p subc.cc.u32 %6, %6, %14;\n\t@p subc.u32 %7, %7, %15;\n\t}"
                     : "+r"(z[0]), "+r"(z[1]), "+r"(z[2]), "+r"(z[3]), "+r"(z[4]), "+r"(z[5]),
                       "+r"(z[6]), "+r"(z[7])
                     : "r"(m[0]), "r"(m[1]), "r"(m[2]), "r"(m[3]), "r"(m[4]), "r"(m[5]),
                       "r"(m[6]), "r"(m[7]), "r"((uint32_t)p));
    } else {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %8, 0;\n\t"
                     "@p sub.cc.u32 %0, %0, %4;\n\t@p subc.cc.u32 %1, %1, %5;\n\t@p subc.cc.u32 %2, %2, %6;\n\t"
                     "@p subc.u32 %3, %3, %7;\n\t}"
                     : "+r"(z[0]), "+r"(z[1]), "+r"(z[2]), "+r"(z[3])
                     : "r"(m[0]), "r"(m[1]), "r"(m[2]), "r"(m[3]), "r"((uint32_t)p));
    }
#else
    if (p) {
        CC c;
        c.sub_cc(z[0], z[0], m[0]);
        for (int i = 1; i < NL - 1; i++) c.subc_cc(z[i], z[i], m[i]);
        c.subc(z[NL - 1], z[NL - 1], m[NL - 1]);
    }
#endif
}

template <int NL>
RS_HD void pred_add(uint32_t *z, const uint32_t *m, bool p) {
#ifdef __CUDA_ARCH__
    static_assert(NL == 4 || NL == 8, "pred_add: 4 or 8 limbs");
    if constexpr (NL == 8) {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %16, 0;\n\t"
                     "@p add.cc.u32 %0, %0, %8;\n\t@p addc.cc.u32 %1, %1, %9;\n\t@p addc.cc.u32 %2, %2, %10;\n\t"
                     "@p addc.cc.u32 %3, %3, %11;\n\t@p addc.cc.u32 %4, %4, %12;\n\t@p addc.cc.u32 %5, %5, %13;\n\t"
                     "@p addc.cc.u32 %6, %6, %14;\n\t@p addc.u32 %7, %7, %15;\n\t}"
                     : "+r"(z[0]), "+r"(z[1]), "+r"(z[2]), "+r"(z[3]), "+r"(z[4]), "+r"(z[5]),
                       "+r"(z[6]), "+r"(z[7])
                     : "r"(m[0]), "r"(m[1]), "r"(m[2]), "r"(m[3]), "r"(m[4]), "r"(m[5]),
                       "r"(m[6]), "r"(m[7]), "r"((uint32_t)p));
    } else {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %8, 0;\n\t"
                     "@p add.cc.u32 %0, %0, %4;\n\t@p addc.cc.u32 %1, %1, %5;\n\t@p addc.cc.u32 %2, %2, %6;\n\t"
                     "@p addc.u32 %3, %3, %7;\n\t}"
                     : "+r"(z[0]), "+r"(z[1]), "+r"(z[2]), "+r"(z[3])
                     : "r"(m[0]), "r"(m[1]), "r"(m[2]), "r"(m[3]), "r"((uint32_t)p));
    }
#else
    if (p) {
        CC c;
        c.add_cc(z[0], z[0], m[0]);
        for (int i = 1; i < NL - 1; i++) c.addc_cc(z[i], z[i], m[i]);
        c.addc(z[NL - 1], z[NL - 1], m[NL - 1]);
    }
#endif
}

template <int NL>
RS_HD bool top64_above(const uint32_t *s, const uint32_t *m) {
    const uint64_t a = ((uint64_t)s[NL - 1] << 32) | s[NL - 2], b = ((uint64_t)m[NL - 1] << 32) | m[NL - 2];
    return a > b;
}

// z = x + y, minus 2q when the top 64 bits of the sum exceed those of 2q.  The caller guarantees x + y < beta
// (x < 2q + delta, y a product below 1.56 q).  The sum is >= 2q whenever it is reduced and below 2q + delta when it is not.
template <int NL>
RS_HD void lazy_add_p(uint32_t *z, const uint32_t *x, const uint32_t *y, const uint32_t *q2) {
    CC c;
    uint32_t s[NL];
    c.add_cc(s[0], x[0], y[0]);
#pragma unroll
    for (int i = 1; i < NL - 1; i++) c.addc_cc(s[i], x[i], y[i]);
    c.addc(s[NL - 1], x[NL - 1], y[NL - 1]);
    pred_sub<NL>(s, q2, top64_above<NL>(s, q2));
#pragma unroll
    for (int i = 0; i < NL; i++) z[i] = s[i];
}

// z = x - y, plus 2q when the difference is negative:  x in [0, 2q + delta), y in [0, 2q]  ->  z in [0, 2q + delta)
template <int NL>
RS_HD void lazy_sub_p(uint32_t *z, const uint32_t *x, const uint32_t *y, const uint32_t *q2) {
    CC c;
    uint32_t d[NL], mask;
    c.sub_cc(d[0], x[0], y[0]);
#pragma unroll
    for (int i = 1; i < NL; i++) c.subc_cc(d[i], x[i], y[i]);
    c.subc(mask, 0, 0);
    pred_add<NL>(d, q2, mask != 0);
#pragma unroll
    for (int i = 0; i < NL; i++) z[i] = d[i];
}

}  // namespace rs
