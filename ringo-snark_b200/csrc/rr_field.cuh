// rr_field.cuh -- reduced-radix ("RR") lazy Montgomery arithmetic for the forward big-field NTT.
//
// Why: on sm_100 the carry-in multiply-add IMAD.WIDE.U32.X that a full-radix (32-bit limb) CIOS
// product is built from issues at half the rate of the carry-free IMAD.WIDE.U32 (9.3 vs 16.7 T/s,
// profiles/), and the ~130 carry/select ALU instructions per butterfly saturate the ALU pipe.  With
// 29-bit limbs held in 32-bit registers every partial product is < 2^58, so a 64-bit accumulator
// absorbs a whole column of the product (<= 2 W terms) with plain mad.wide -- no carry chain at
// all -- and additions need no carry propagation either.
//
// Representation: an element is the same integer as the reference's Montgomery value (x * 2^(64 L)
// mod q), sliced into W limbs of B = 29 bits, kept LAZY: any representative in [0, bound * q), limbs
// possibly above 2^29 between normalisations.  Twiddles are stored as omega * 2^(B W) mod q so that
//     mul(a, w') = a * w' * 2^-(B W)  ==  (x * omega) * 2^(64 L)   (mod q)
// i.e. products land back in the reference's Montgomery form.  The final stage canonicalises to
// [0, q), so results are bit-identical to the full-radix kernels and to the reference.
//
// Requires q == 1 (mod 2^29) (true for every jindo modulus: b^k with 2^31 | b^k); QZ = 1 additionally
// uses q == 1 (mod 2^58) (limb 1 of q is zero).  Plain C++ (no PTX): also compiles for the host, where
// tests/test_device_field_host.py checks it against Python integers.
#pragma once
#include <stdint.h>

#include "field.cuh"

namespace rs {

constexpr int RR_B = 29;
constexpr uint32_t RR_MASK = (1u << RR_B) - 1;

template <int W>
struct RRParams {
    uint32_t q[W];        // q in 29-bit limbs (q[0] == 1)
    uint32_t c3q[W];      // 3q in limb-dominating form: every limb >= 2^29 - 1 (see rr_sub)
    uint32_t q2[W];       // 2q, normalised limbs (final canonicalisation)
    uint32_t one[W];      // 2^(29 W) mod q, normalised ("1" of the RR Montgomery domain)
};

// z = a * b * 2^-(29 W) mod q, z in [0, 2q) with normalised limbs (< 2^29, top limb holds the rest).
// a: lazy (value < 2^(29W) / q * q, limbs < 2^31.4), b: normalised limbs, value < q.
template <int W, int QZ>
RS_HD void rr_mul(uint32_t *z, const uint32_t *a, const uint32_t *b, const RRParams<W> &P) {
    uint64_t acc[W];
#pragma unroll
    for (int j = 0; j < W; j++) acc[j] = 0;
#pragma unroll
    for (int i = 0; i < W; i++) {
        // acc[j] holds column i + j
#pragma unroll
        for (int j = 0; j < W; j++) acc[j] += (uint64_t)a[j] * b[i];
        // Montgomery step in radix 2^29 with q[0] == 1: m = -acc[0] mod 2^29
        const uint32_t m = (0u - (uint32_t)acc[0]) & RR_MASK;
        acc[0] += m;                                   // low 29 bits become zero
#pragma unroll
        for (int j = (QZ ? 2 : 1); j < W; j++) acc[j] += (uint64_t)m * P.q[j];
        // shift the window by one limb
        const uint64_t carry = acc[0] >> RR_B;
#pragma unroll
        for (int j = 0; j < W - 1; j++) acc[j] = acc[j + 1];
        acc[W - 1] = 0;
        acc[0] += carry;
    }
    // normalise
    uint64_t carry = 0;
#pragma unroll
    for (int j = 0; j < W - 1; j++) {
        const uint64_t t = acc[j] + carry;
        z[j] = (uint32_t)t & RR_MASK;
        carry = t >> RR_B;
    }
    z[W - 1] = (uint32_t)(acc[W - 1] + carry);
}

// lazy add: limb-wise, no carries
template <int W>
RS_HD void rr_add(uint32_t *z, const uint32_t *x, const uint32_t *y) {
#pragma unroll
    for (int j = 0; j < W; j++) z[j] = x[j] + y[j];
}
// lazy subtract of a NORMALISED y in [0, 2q): z = x + 3q - y, limb-wise without borrows
// (c3q is 3q rewritten so that every limb dominates a normalised limb of y)
template <int W>
RS_HD void rr_sub(uint32_t *z, const uint32_t *x, const uint32_t *y, const RRParams<W> &P) {
#pragma unroll
    for (int j = 0; j < W; j++) z[j] = x[j] + P.c3q[j] - y[j];
}
// carry propagation: limbs back below 2^29 (top limb keeps the excess)
template <int W>
RS_HD void rr_normalise(uint32_t *x) {
    uint32_t carry = 0;
#pragma unroll
    for (int j = 0; j < W - 1; j++) {
        const uint32_t t = x[j] + carry;
        x[j] = t & RR_MASK;
        carry = t >> RR_B;
    }
    x[W - 1] += carry;
}

// x (normalised, value < 2^(29 W)) >= y (normalised)?  then x -= y.  Returns nothing; branch-free.
template <int W>
RS_HD void rr_cond_sub(uint32_t *x, const uint32_t *y) {
    uint32_t d[W];
    int32_t borrow = 0;
#pragma unroll
    for (int j = 0; j < W - 1; j++) {
        const int32_t t = (int32_t)x[j] - (int32_t)y[j] + borrow;     // in (-2^29 - 1, 2^29)
        d[j] = (uint32_t)t & RR_MASK;
        borrow = t >> RR_B;                                            // 0 or -1 (arithmetic shift)
    }
    const int32_t top = (int32_t)x[W - 1] - (int32_t)y[W - 1] + borrow;
    d[W - 1] = (uint32_t)top;
    const bool ge = top >= 0;
#pragma unroll
    for (int j = 0; j < W; j++) x[j] = ge ? d[j] : x[j];
}

// full-radix (NL x 32 bit, canonical) -> RR limbs: pure bit slicing
template <int NL, int W>
RS_HD void rr_from_words(uint32_t *r, const uint32_t *w) {
#pragma unroll
    for (int j = 0; j < W; j++) {
        const int bit = RR_B * j, wi = bit / 32, sh = bit % 32;
        uint32_t lo = (wi < NL) ? w[wi] : 0u, hi = (wi + 1 < NL) ? w[wi + 1] : 0u;
        const uint32_t v = sh ? ((lo >> sh) | (hi << (32 - sh))) : lo;
        r[j] = (j == W - 1) ? v : (v & RR_MASK);
    }
}
// RR limbs (normalised, value < 2^(32 NL)) -> full radix words
template <int NL, int W>
RS_HD void rr_to_words(uint32_t *w, const uint32_t *r) {
#pragma unroll
    for (int i = 0; i < NL; i++) w[i] = 0;
#pragma unroll
    for (int j = 0; j < W; j++) {
        const int bit = RR_B * j, wi = bit / 32, sh = bit % 32;
        if (wi < NL) w[wi] |= r[j] << sh;
        if (sh && wi + 1 < NL) w[wi + 1] |= r[j] >> (32 - sh);
    }
}

// final canonicalisation of a lazy value: one multiplication by "1" brings it to [0, 2q),
// one conditional subtraction to [0, q)
template <int W, int QZ>
RS_HD void rr_canonical(uint32_t *x, const RRParams<W> &P) {
    uint32_t t[W];
    rr_mul<W, QZ>(t, x, P.one, P);
#pragma unroll
    for (int j = 0; j < W; j++) x[j] = t[j];
    rr_cond_sub<W>(x, P.q);
}

}  // namespace rs
