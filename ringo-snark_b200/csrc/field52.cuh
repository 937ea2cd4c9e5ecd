// field52.cuh -- big-field Montgomery arithmetic on the FP64 pipe (DFMA), for the NTT butterflies of
// math/bigpoly/ntt.go:254-259 / :365-370 over the 8-word (220 / 240 / 255-bit) jindo-modulus fields.
//
// Why: a 32x32+64 multiply-add (IMAD.WIDE) issues at 32 lanes/clk/SM on sm_100a, DFMA at 64 lanes/clk/SM on a
// pipe the integer kernels leave idle.  An element is held as ND limbs of 52 bits; a limb product a*b < 2^104 is
// split exactly by two fused multiply-adds in round-towards-zero mode:
//     ph = fma_rz(a, b, 2^104)                  = 2^104 + floor(ab / 2^52) 2^52       (ulp of that binade is 2^52)
//     pl = fma_rz(a, b, (2^104 + 2^52) - ph)    = 2^52  + (ab mod 2^52)               (exact, ulp 1)
// so the low 52 bits of the two bit patterns ARE the two halves of the product, and column sums are plain 64-bit
// integer additions of the patterns (the exponent fields are cancelled by a constant folded into the column's
// initial value).  3 FP64-pipe instructions + 2 integer adds per 52x52 product, against 1 IMAD.WIDE (4 pipe cycles)
// per 32x32 product.
//
// Montgomery radix here is R' = 2^(52 ND) (2^260 for ND = 5), not the reference's R = 2^(64 L).  Data keeps the
// reference's representation x R mod q (jindo/internal/zp/element.go:24-37); only the twiddle tables are re-scaled
// on the host to w R' mod q, so  mont52(x R, w R') = x w R  -- the same field element the reference's Mul gives.
// All arithmetic is exact modulo q; values are lazy (any representative below 2^(52 ND)) between butterflies and
// are brought back to [0, q) before they leave the kernel, so results are bit-identical to the reference.
//
// The same source compiles for the host (tests/host_field52_test.cpp runs it under FE_TOWARDZERO against
// Python integers).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "field.cuh"

namespace rs {

constexpr uint64_t F52_MASK = (1ull << 52) - 1;
constexpr uint64_t F52_EXP_LO = 0x4330000000000000ull;   // bit pattern of 2^52
constexpr uint64_t F52_EXP_HI = 0x4670000000000000ull;   // bit pattern of 2^104

RS_HD uint64_t f52_bits(double x) {
#ifdef __CUDA_ARCH__
    return (uint64_t)__double_as_longlong(x);
#else
    uint64_t u; memcpy(&u, &x, 8); return u;
#endif
}
RS_HD double f52_dbl(uint64_t u) {
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)u);
#else
    double x; memcpy(&x, &u, 8); return x;
#endif
}
RS_HD double f52_fma_rz(double a, double b, double c) {
#ifdef __CUDA_ARCH__
    return __fma_rz(a, b, c);
#else
    return fma(a, b, c);   // the host test harness runs under fesetround(FE_TOWARDZERO)
#endif
}
// exact integer -> double for 0 <= v < 2^52
RS_HD double f52_to_double(uint64_t v) { return f52_dbl(v | F52_EXP_LO) - 0x1p52; }

// exact split of a limb product: hi/lo are the raw bit patterns (exponent fields included)
RS_HD void f52_split(double a, double b, uint64_t &hi, uint64_t &lo) {
    const double ph = f52_fma_rz(a, b, 0x1p104);
    const double sub = (0x1p104 + 0x1p52) - ph;
    const double pl = f52_fma_rz(a, b, sub);
    hi = f52_bits(ph);
    lo = f52_bits(pl);
}

// multiples of q used as subtraction offsets (u - v + k q stays non-negative), see ntt52.cuh for the bound bookkeeping
constexpr int F52_NOFF = 5;
constexpr int F52_OFF_MULT[F52_NOFF] = {2, 4, 6, 10, 14};

template <int ND>
struct Field52 {
    double q[ND];                 // 52-bit limbs of q
    double qinv;                  // -q^-1 mod 2^52 (used when q[0] != 1)
    long long qi[ND];             // limbs of q (integers)
    long long off[F52_NOFF][ND];  // limbs of 2q, 4q, 6q, 10q, 14q (normalised; the top limb takes the excess)
    unsigned est_magic;           // quotient estimate of f52_reduce: e = umulhi(V >> P, est_magic) >> est_mshift,
    int est_mshift;               //   P = bitlen(q) - 25, est_magic = floor(2^(32 + est_mshift) / ((q >> P) + 1))
    int est_sh;                   // P - 52 (ND - 2): position of P relative to limb ND-2
    int q0one;                    // q == 1 (mod 2^52)
    int max_lazy;                 // floor(2^(52 ND) / q): lazy values must stay below max_lazy * q
};

// ---- representation changes -------------------------------------------------------------------
// NL 32-bit words (little endian, value < 2^(32 NL)) -> ND limbs of 52 bits
template <int NL, int ND>
RS_HD void f52_from_words(long long (&l)[ND], const uint32_t *w) {
#pragma unroll
    for (int k = 0; k < ND; k++) {
        const int bit = 52 * k, wi = bit / 32, sh = bit % 32;
        const uint64_t w0 = wi < NL ? w[wi] : 0, w1 = wi + 1 < NL ? w[wi + 1] : 0, w2 = wi + 2 < NL ? w[wi + 2] : 0;
        uint64_t v = (w0 | (w1 << 32)) >> sh;
        if (sh) v |= w2 << (64 - sh);
        l[k] = (long long)(v & F52_MASK);
    }
}
// normalised limbs (0 <= l[k] < 2^52, value < 2^(32 NL)) -> NL words
template <int NL, int ND>
RS_HD void f52_to_words(uint32_t *w, const long long (&l)[ND]) {
#pragma unroll
    for (int j = 0; j < NL; j++) {
        const int bit = 32 * j, k = bit / 52, o = bit % 52;
        uint64_t v = (uint64_t)l[k] >> o;
        if (o > 20 && k + 1 < ND) v |= (uint64_t)l[k + 1] << (52 - o);
        w[j] = (uint32_t)v;
    }
}

// carry propagation on signed limbs; the value must be in [0, 2^(52 ND)).  The top limb absorbs the rest.
template <int ND>
RS_HD void f52_normalize(long long (&l)[ND]) {
#pragma unroll
    for (int k = 0; k < ND - 1; k++) {
        const long long c = l[k] >> 52;
        l[k] &= (long long)F52_MASK;
        l[k + 1] += c;
    }
}

template <int ND>
RS_HD void f52_to_doubles(double (&d)[ND], const long long (&l)[ND]) {
#pragma unroll
    for (int k = 0; k < ND; k++) d[k] = f52_to_double((uint64_t)l[k]);
}

// ---- Montgomery product ------------------------------------------------------------------------
// number of (i, j) with i + j == k, 0 <= i < ND, j0 <= j < ND
RS_HD constexpr int f52_pairs(int k, int ND, int j0) {
    int n = 0;
    for (int i = 0; i < ND; i++) { const int j = k - i; if (j >= j0 && j < ND) n++; }
    return n;
}
// initial value of column k: minus the exponent fields of every pattern that will be added to it
template <int ND, bool Q0ONE>
RS_HD constexpr uint64_t f52_col_init(int k) {
    const int j0 = Q0ONE ? 1 : 0;
    uint64_t s = 0;
    const int n_lo = f52_pairs(k, ND, 0) + f52_pairs(k, ND, j0);
    const int n_hi = (k >= 1) ? f52_pairs(k - 1, ND, 0) + f52_pairs(k - 1, ND, j0) : 0;
    s -= (uint64_t)n_lo * F52_EXP_LO;
    s -= (uint64_t)n_hi * F52_EXP_HI;
    return s;
}

// r = a * b / 2^(52 ND) (mod q) as column sums (unnormalised limbs, each < 2^58), value < q + a b / 2^(52 ND) <= 2q
//   a : normalised limbs as doubles, any value < 2^(52 ND);  b : normalised limbs as doubles, value < q
//   Q0ONE : q == 1 (mod 2^52)  =>  m = -t (mod 2^52) and the products by q[0] vanish
template <int ND, bool Q0ONE>
RS_HD void mont52(long long (&r)[ND], const double (&a)[ND], const double (&b)[ND], const Field52<ND> &F) {
    uint64_t c[2 * ND];
#pragma unroll
    for (int k = 0; k < 2 * ND; k++) c[k] = f52_col_init<ND, Q0ONE>(k);
#pragma unroll
    for (int i = 0; i < ND; i++) {
#pragma unroll
        for (int j = 0; j < ND; j++) {
            uint64_t hi, lo;
            f52_split(a[i], b[j], hi, lo);
            c[i + j] += lo;
            c[i + j + 1] += hi;
        }
    }
#pragma unroll
    for (int i = 0; i < ND; i++) {
        // every contribution to column i except this round's own has been added: c[i] is its true value
        uint64_t m;
        if (Q0ONE) {
            m = (0 - c[i]) & F52_MASK;
        } else {
            // the generic column initial value already contains -EXP_LO for the m*q[0] low half of this round
            // (and -EXP_HI in column i+1); take them out of the value we read
            const uint64_t t = (c[i] + F52_EXP_LO) & F52_MASK;
            uint64_t hi, lo;
            f52_split(f52_to_double(t), F.qinv, hi, lo);
            m = lo & F52_MASK;
        }
        const double md = f52_to_double(m);
#pragma unroll
        for (int j = Q0ONE ? 1 : 0; j < ND; j++) {
            uint64_t hi, lo;
            f52_split(md, F.q[j], hi, lo);
            c[i + j] += lo;
            c[i + j + 1] += hi;
        }
        if (Q0ONE) c[i] += m;
        c[i + 1] += c[i] >> 52;     // c[i] is a multiple of 2^52 now
    }
#pragma unroll
    for (int k = 0; k < ND; k++) r[k] = (long long)c[ND + k];
}

// ---- lazy reduction -------------------------------------------------------------------------------
// l: signed unnormalised limbs, |l[k]| < 2^62, value V in [0, 64 q).  On return the limbs are normalised and the
// value is V - e q in [0, 2.0001 q): e = umulhi(th, magic) >> s never exceeds floor(V / q) (th <= V >> P and
// (q >> P) + 1 > q / 2^P) and falls short of it by at most 1 (th >= (V >> P) - 3, q >> P >= 2^24).
template <int ND>
RS_HD void f52_reduce(long long (&l)[ND], const Field52<ND> &F) {
    long long t;
    if (F.est_sh >= 52) t = (l[ND - 1] + (l[ND - 2] >> 52)) >> (F.est_sh - 52);
    else t = (l[ND - 1] << (52 - F.est_sh)) + (l[ND - 2] >> F.est_sh);
    t -= 2;                                              // the limbs below ND-2 and the floors: at most 2 units
    const uint32_t th = t < 0 ? 0u : (uint32_t)t;
    const long long e = (long long)(mul_hi(th, F.est_magic) >> F.est_mshift);
#pragma unroll
    for (int k = 0; k < ND; k++) l[k] -= e * F.qi[k];
    f52_normalize<ND>(l);
}

// ---- host-side set-up (plain C++; used by capi.cu and by the host tests) -----------------------------------
// q: NL 32-bit words.  Returns false when the field does not fit the scheme (2^(52 ND) < 36 q).
template <int NL, int ND>
inline bool f52_setup(Field52<ND> &F, const uint32_t *q) {
    long long ql[ND];
    f52_from_words<NL, ND>(ql, q);
    for (int k = 0; k < ND; k++) { F.qi[k] = ql[k]; F.q[k] = (double)ql[k]; }
    F.q0one = (ql[0] == 1);
    // -q^-1 mod 2^52 by Newton iteration on the low limb
    uint64_t inv = 1;
    for (int i = 0; i < 6; i++) inv *= 2 - (uint64_t)ql[0] * inv;
    F.qinv = (double)((0 - inv) & F52_MASK);
    for (int o = 0; o < F52_NOFF; o++) {
        unsigned __int128 c = 0;
        for (int k = 0; k < ND; k++) {
            c += (unsigned __int128)ql[k] * F52_OFF_MULT[o];
            F.off[o][k] = (k == ND - 1) ? (long long)c : (long long)((uint64_t)c & F52_MASK);
            c >>= 52;
        }
    }
    int bits = 0;
    for (int i = NL - 1; i >= 0; i--) if (q[i]) { bits = 32 * i + (32 - __builtin_clz(q[i])); break; }
    const int P = bits - 25;
    if (P < 52 * (ND - 2)) return false;
    // (q >> P) + 1
    uint64_t qh = 0;
    for (int b = 0; b < 25; b++) { const int bit = P + b; if ((q[bit / 32] >> (bit % 32)) & 1) qh |= 1ull << b; }
    qh += 1;
    int s = 63 - __builtin_clzll(qh);
    unsigned __int128 mg = ((unsigned __int128)1 << (32 + s)) / qh;
    if (mg > 0xffffffffull) mg = 0xffffffffull;
    F.est_magic = (unsigned)mg;
    F.est_mshift = s;
    F.est_sh = P - 52 * (ND - 2);
    // floor(2^(52 ND) / q), capped
    const int slack = 52 * ND - bits;                    // 2^(52 ND) / q is in (2^slack, 2^(slack+1))
    if (slack >= 20) F.max_lazy = 1 << 20;
    else {
        // exact for the small case: compare k q with 2^(52 ND) using long arithmetic on words
        int k = 1 << slack;
        for (;; k++) {
            unsigned __int128 c = 0; bool over = false;
            uint32_t w[NL + 2];
            for (int i = 0; i < NL + 2; i++) { c += (unsigned __int128)(i < NL ? q[i] : 0) * (unsigned)(k + 1); w[i] = (uint32_t)c; c >>= 32; }
            // (k+1) q >= 2^(52 ND) ?
            const int tb = 52 * ND;
            for (int i = NL + 1; i >= 0 && !over; i--) {
                for (int b = 31; b >= 0; b--) if ((w[i] >> b) & 1) { if (32 * i + b >= tb) over = true; break; }
                if (w[i]) break;
            }
            if (over) break;
        }
        F.max_lazy = k;
    }
    return F.max_lazy >= 36;
}

}  // namespace rs
