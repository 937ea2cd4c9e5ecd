// field.cuh -- multi-limb Montgomery arithmetic for the jindo-modulus prime fields
// (reference semantics: gnark-crypto generated zp packages,
//  jindo/internal/zp/element.go:397-466 Add/Sub/Neg, :481-602 Mul, :604-653 fromMont).
//
// Elements are NL 32-bit limbs (NL = 2 * Limbs of the reference's [L]uint64,
// same little-endian byte layout), Montgomery form with R = 2^(32 NL) = 2^(64 L),
// always fully reduced to [0, q) exactly like the reference.
//
// The multiplier is a CIOS Montgomery product on 32x32->64 multiply-adds arranged
// as two carry chains ("even" limbs and "odd" limbs) so that ptxas can fuse each
// mad.lo.cc/madc.hi.cc pair into one IMAD.WIDE.U32(.X): 2*NL^2 + NL wide MADs for a
// generic modulus.  Every jindo modulus is q = b^k + 1 with a large power of two
// dividing b^k, so the low limbs of q are (1, 0, ...): QS = 1 (q == 1 mod 2^32) and
// QS = 2 (q == 1 mod 2^64) drop the multiplications by those limbs.
//
// The same instruction sequence compiles for the host (carry flag emulated) so that
// tests/ can check it against Python integers without a GPU (tests/host_field_test.cpp).
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define RS_HD __host__ __device__ __forceinline__
#define RS_D __device__ __forceinline__
#else
#define RS_HD inline
#define RS_D inline
#endif

namespace rs {

// ---------------------------------------------------------------------------
// carry-chain primitives: PTX on the device, emulated on the host
// ---------------------------------------------------------------------------
#ifdef __CUDA_ARCH__
struct CC {
    RS_D void add_cc(uint32_t &d, uint32_t a, uint32_t b) { asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); }
    RS_D void addc_cc(uint32_t &d, uint32_t a, uint32_t b) { asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); }
    RS_D void addc(uint32_t &d, uint32_t a, uint32_t b) { asm volatile("addc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); }
    RS_D void sub_cc(uint32_t &d, uint32_t a, uint32_t b) { asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); }
    RS_D void subc_cc(uint32_t &d, uint32_t a, uint32_t b) { asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); }
    RS_D void subc(uint32_t &d, uint32_t a, uint32_t b) { asm volatile("subc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); }
    RS_D void mad_lo_cc(uint32_t &d, uint32_t a, uint32_t b, uint32_t c) { asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); }
    RS_D void madc_lo_cc(uint32_t &d, uint32_t a, uint32_t b, uint32_t c) { asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); }
    RS_D void madc_hi_cc(uint32_t &d, uint32_t a, uint32_t b, uint32_t c) { asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); }
};
#else
struct CC {
    uint32_t cf = 0;
    void add_cc(uint32_t &d, uint32_t a, uint32_t b) { uint64_t s = (uint64_t)a + b; d = (uint32_t)s; cf = (uint32_t)(s >> 32); }
    void addc_cc(uint32_t &d, uint32_t a, uint32_t b) { uint64_t s = (uint64_t)a + b + cf; d = (uint32_t)s; cf = (uint32_t)(s >> 32); }
    void addc(uint32_t &d, uint32_t a, uint32_t b) { d = a + b + cf; }
    void sub_cc(uint32_t &d, uint32_t a, uint32_t b) { uint64_t s = (uint64_t)a - b; d = (uint32_t)s; cf = (uint32_t)(s >> 63); }
    void subc_cc(uint32_t &d, uint32_t a, uint32_t b) { uint64_t s = (uint64_t)a - b - cf; d = (uint32_t)s; cf = (uint32_t)(s >> 63); }
    void subc(uint32_t &d, uint32_t a, uint32_t b) { d = a - b - cf; }
    void mad_lo_cc(uint32_t &d, uint32_t a, uint32_t b, uint32_t c) { uint64_t s = (uint64_t)(uint32_t)((uint64_t)a * b) + c; d = (uint32_t)s; cf = (uint32_t)(s >> 32); }
    void madc_lo_cc(uint32_t &d, uint32_t a, uint32_t b, uint32_t c) { uint64_t s = (uint64_t)(uint32_t)((uint64_t)a * b) + c + cf; d = (uint32_t)s; cf = (uint32_t)(s >> 32); }
    void madc_hi_cc(uint32_t &d, uint32_t a, uint32_t b, uint32_t c) { uint64_t s = (((uint64_t)a * b) >> 32) + c + cf; d = (uint32_t)s; cf = (uint32_t)(s >> 32); }
};
#endif

RS_HD uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
RS_HD uint32_t mul_hi(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

// ---------------------------------------------------------------------------
// field parameters (passed to kernels by value: they live in the constant bank,
// so q limbs are constant-bank operands of the IMADs and cost no registers)
// ---------------------------------------------------------------------------
template <int NL>
struct FieldParams {
    uint32_t q[NL];
    uint32_t qinv;        // -q^-1 mod 2^32
    uint32_t one[NL];     // R mod q
    uint32_t r2[NL];      // R^2 mod q
    uint32_t q2[NL];      // 2q (lazy [0, 2q) arithmetic; only meaningful when 3q < 2^(32 NL))
};

// ---------------------------------------------------------------------------
// modular add / sub / neg, results canonical in [0, q)
// ---------------------------------------------------------------------------
template <int NL>
RS_HD void mod_add(uint32_t *z, const uint32_t *x, const uint32_t *y, const FieldParams<NL> &F) {
    CC c;
    uint32_t s[NL], d[NL], carry, borrow;
    c.add_cc(s[0], x[0], y[0]);
#pragma unroll
    for (int i = 1; i < NL; i++) c.addc_cc(s[i], x[i], y[i]);
    c.addc(carry, 0, 0);
    c.sub_cc(d[0], s[0], F.q[0]);
#pragma unroll
    for (int i = 1; i < NL; i++) c.subc_cc(d[i], s[i], F.q[i]);
    c.subc(borrow, carry, 0);             // 0xffffffff iff (carry:s) < q
    const bool keep = (borrow == 0xffffffffu);
#pragma unroll
    for (int i = 0; i < NL; i++) z[i] = keep ? s[i] : d[i];
}

template <int NL>
RS_HD void mod_sub(uint32_t *z, const uint32_t *x, const uint32_t *y, const FieldParams<NL> &F) {
    CC c;
    uint32_t d[NL], mask;
    c.sub_cc(d[0], x[0], y[0]);
#pragma unroll
    for (int i = 1; i < NL; i++) c.subc_cc(d[i], x[i], y[i]);
    c.subc(mask, 0, 0);                   // 0xffffffff if x < y
    c.add_cc(z[0], d[0], F.q[0] & mask);
#pragma unroll
    for (int i = 1; i < NL - 1; i++) c.addc_cc(z[i], d[i], F.q[i] & mask);
    c.addc(z[NL - 1], d[NL - 1], F.q[NL - 1] & mask);
}

template <int NL>
RS_HD void mod_neg(uint32_t *z, const uint32_t *x, const FieldParams<NL> &F) {
    uint32_t any = 0;
#pragma unroll
    for (int i = 0; i < NL; i++) any |= x[i];
    const uint32_t mask = any ? 0xffffffffu : 0u;   // -0 = 0 (element.go:448-466)
    CC c;
    c.sub_cc(z[0], F.q[0] & mask, x[0]);
#pragma unroll
    for (int i = 1; i < NL - 1; i++) c.subc_cc(z[i], F.q[i] & mask, x[i]);
    c.subc(z[NL - 1], F.q[NL - 1] & mask, x[NL - 1]);
}

// ---------------------------------------------------------------------------
// Montgomery product z = x * y * R^-1 mod q, canonical.
//
// Running sum T is kept as two limb arrays with 64-bit-aligned products:
//   T = sum A[k] 2^(32k) + sum B[k] 2^(32(k+1)) + top 2^(32(NL+1))
// Row i:   B <- (previous A >> 64) + odd-limb products of x*y_i   (in place, shifted)
//          A <- previous B        + even-limb products of x*y_i
//          m = A[0] * qinv ;  B += odd limbs of m*q ; A += even limbs of m*q  (A[0] becomes 0)
// and the two arrays swap roles for the next row (T >>= 32).
// ---------------------------------------------------------------------------
template <int NL, int QS>
RS_HD void mont_row(uint32_t *A, uint32_t *B, uint32_t &top, const uint32_t *x, uint32_t yi,
                    const FieldParams<NL> &F, bool first) {
    // On entry (not first): `A` is the array aligned at limb 0 for this row (previous B),
    // `B` holds the previous A whose limb 0 is zero; B[1] folds into A[0] and B[k+2]
    // becomes the new B[k].  `top` carries weight 2^(32 NL) into the new B[NL-1].
    CC c;
    if (first) {
#pragma unroll
        for (int j = 0; j < NL; j += 2) {
            A[j] = mul_lo(x[j], yi);
            A[j + 1] = mul_hi(x[j], yi);
            B[j] = mul_lo(x[j + 1], yi);
            B[j + 1] = mul_hi(x[j + 1], yi);
        }
        top = 0;
    } else {
        c.add_cc(A[0], A[0], B[1]);
#pragma unroll
        for (int j = 0; j < NL - 2; j += 2) {
            c.madc_lo_cc(B[j], x[j + 1], yi, B[j + 2]);
            c.madc_hi_cc(B[j + 1], x[j + 1], yi, B[j + 3]);
        }
        c.madc_lo_cc(B[NL - 2], x[NL - 1], yi, 0);
        c.madc_hi_cc(B[NL - 1], x[NL - 1], yi, top);
        c.addc(top, 0, 0);
        c.mad_lo_cc(A[0], x[0], yi, A[0]);
        c.madc_hi_cc(A[1], x[0], yi, A[1]);
#pragma unroll
        for (int j = 2; j < NL; j += 2) {
            c.madc_lo_cc(A[j], x[j], yi, A[j]);
            c.madc_hi_cc(A[j + 1], x[j], yi, A[j + 1]);
        }
        c.addc_cc(B[NL - 1], B[NL - 1], 0);
        c.addc(top, top, 0);
    }
    // reduction step
    // (for QS >= 1 qinv == 0xffffffff and m == -A[0]; the multiply form is kept because ptxas
    //  only fuses the following mad.lo/madc.hi pairs into IMAD.WIDE when m comes from an IMAD)
    const uint32_t m = mul_lo(A[0], F.qinv);
    if (QS >= 2) {
        // q[1] == 0: the first odd product vanishes
        if (NL > 2) {
            c.mad_lo_cc(B[2], m, F.q[3], B[2]);
            c.madc_hi_cc(B[3], m, F.q[3], B[3]);
#pragma unroll
            for (int j = 5; j < NL; j += 2) {
                c.madc_lo_cc(B[j - 1], m, F.q[j], B[j - 1]);
                c.madc_hi_cc(B[j], m, F.q[j], B[j]);
            }
            c.addc(top, top, 0);
        }
    } else {
        c.mad_lo_cc(B[0], m, F.q[1], B[0]);
        c.madc_hi_cc(B[1], m, F.q[1], B[1]);
#pragma unroll
        for (int j = 3; j < NL; j += 2) {
            c.madc_lo_cc(B[j - 1], m, F.q[j], B[j - 1]);
            c.madc_hi_cc(B[j], m, F.q[j], B[j]);
        }
        c.addc(top, top, 0);
    }
    if (QS >= 1) {
        // q[0] == 1, m == -A[0]: A[0] + m == 0 with carry iff A[0] != 0
        c.add_cc(A[0], A[0], m);
        c.addc_cc(A[1], A[1], 0);
    } else {
        c.mad_lo_cc(A[0], m, F.q[0], A[0]);
        c.madc_hi_cc(A[1], m, F.q[0], A[1]);
    }
#pragma unroll
    for (int j = 2; j < NL; j += 2) {
        c.madc_lo_cc(A[j], m, F.q[j], A[j]);
        c.madc_hi_cc(A[j + 1], m, F.q[j], A[j + 1]);
    }
    c.addc_cc(B[NL - 1], B[NL - 1], 0);
    c.addc(top, top, 0);
}

template <int NL, int QS>
RS_HD void mont_mul(uint32_t *z, const uint32_t *x, const uint32_t *y, const FieldParams<NL> &F) {
    static_assert(NL % 2 == 0 && NL >= 2, "NL must be even");
    uint32_t P[NL], Q[NL], top;
    // rows alternate the roles of P and Q
#pragma unroll
    for (int i = 0; i < NL; i += 2) {
        mont_row<NL, QS>(P, Q, top, x, y[i], F, i == 0);
        mont_row<NL, QS>(Q, P, top, x, y[i + 1], F, false);
    }
    // after an even number of rows: aligned-at-0 array for the *next* row would be P's
    // partner: last call used A=Q, B=P  =>  T = (Q >> 32, Q[0]==0) + P * 2^32 ... in the
    // shifted frame: r = P + (Q >> 64 shifted) i.e. r[k] = P[k] + Q[k+1]
    CC c;
    uint32_t r[NL], rt;
    c.add_cc(r[0], P[0], Q[1]);
#pragma unroll
    for (int k = 1; k < NL - 1; k++) c.addc_cc(r[k], P[k], Q[k + 1]);
    c.addc_cc(r[NL - 1], P[NL - 1], 0);
    c.addc(rt, top, 0);
    // conditional subtraction: r (NL+1 limbs, rt in {0,1}) < 2q
    uint32_t d[NL], borrow;
    c.sub_cc(d[0], r[0], F.q[0]);
#pragma unroll
    for (int k = 1; k < NL; k++) c.subc_cc(d[k], r[k], F.q[k]);
    c.subc(borrow, rt, 0);
    const bool keep = (borrow == 0xffffffffu);   // r < q
#pragma unroll
    for (int k = 0; k < NL; k++) z[k] = keep ? r[k] : d[k];
}

// ---------------------------------------------------------------------------
// Lazy [0, 2q) variants, for moduli with q < 2^(32 NL) / 3 (every jindo modulus except the
// 128-bit one).  The NTT keeps all intermediate values in [0, 2q): the Montgomery product then
// needs neither the overflow word `top` nor the final conditional subtraction (x < 2q, y < q give
// x*y*R^-1 + q < 2q), which removes ~50 of the ~300 instructions of a butterfly.  Values are brought
// back to [0, q) once, in the last stage.
// ---------------------------------------------------------------------------
template <int NL>
RS_HD void mod_add_m(uint32_t *z, const uint32_t *x, const uint32_t *y, const uint32_t *mod) {
    CC c;
    uint32_t s[NL], d[NL], carry, borrow;
    c.add_cc(s[0], x[0], y[0]);
#pragma unroll
    for (int i = 1; i < NL; i++) c.addc_cc(s[i], x[i], y[i]);
    c.addc(carry, 0, 0);
    c.sub_cc(d[0], s[0], mod[0]);
#pragma unroll
    for (int i = 1; i < NL; i++) c.subc_cc(d[i], s[i], mod[i]);
    c.subc(borrow, carry, 0);
    const bool keep = (borrow == 0xffffffffu);
#pragma unroll
    for (int i = 0; i < NL; i++) z[i] = keep ? s[i] : d[i];
}

template <int NL>
RS_HD void mod_sub_m(uint32_t *z, const uint32_t *x, const uint32_t *y, const uint32_t *mod) {
    CC c;
    uint32_t d[NL], mask;
    c.sub_cc(d[0], x[0], y[0]);
#pragma unroll
    for (int i = 1; i < NL; i++) c.subc_cc(d[i], x[i], y[i]);
    c.subc(mask, 0, 0);
    c.add_cc(z[0], d[0], mod[0] & mask);
#pragma unroll
    for (int i = 1; i < NL - 1; i++) c.addc_cc(z[i], d[i], mod[i] & mask);
    c.addc(z[NL - 1], d[NL - 1], mod[NL - 1] & mask);
}

// x in [0, 2 mod) -> [0, mod)
template <int NL>
RS_HD void cond_sub(uint32_t *x, const uint32_t *mod) {
    CC c;
    uint32_t d[NL], borrow;
    c.sub_cc(d[0], x[0], mod[0]);
#pragma unroll
    for (int i = 1; i < NL; i++) c.subc_cc(d[i], x[i], mod[i]);
    c.subc(borrow, 0, 0);
    const bool keep = (borrow == 0xffffffffu);
#pragma unroll
    for (int i = 0; i < NL; i++) x[i] = keep ? x[i] : d[i];
}

// mont_row without the overflow word (valid while T stays below 2^(32 (NL+1)): 3q 2^32 + 2q < 2^(32(NL+1)))
template <int NL, int QS>
RS_HD void mont_row_nt(uint32_t *A, uint32_t *B, const uint32_t *x, uint32_t yi, const FieldParams<NL> &F, bool first) {
    CC c;
    if (first) {
#pragma unroll
        for (int j = 0; j < NL; j += 2) {
            A[j] = mul_lo(x[j], yi);
            A[j + 1] = mul_hi(x[j], yi);
            B[j] = mul_lo(x[j + 1], yi);
            B[j + 1] = mul_hi(x[j + 1], yi);
        }
    } else {
        c.add_cc(A[0], A[0], B[1]);
#pragma unroll
        for (int j = 0; j < NL - 2; j += 2) {
            c.madc_lo_cc(B[j], x[j + 1], yi, B[j + 2]);
            c.madc_hi_cc(B[j + 1], x[j + 1], yi, B[j + 3]);
        }
        c.madc_lo_cc(B[NL - 2], x[NL - 1], yi, 0);
        c.madc_hi_cc(B[NL - 1], x[NL - 1], yi, 0);
        c.mad_lo_cc(A[0], x[0], yi, A[0]);
        c.madc_hi_cc(A[1], x[0], yi, A[1]);
#pragma unroll
        for (int j = 2; j < NL; j += 2) {
            c.madc_lo_cc(A[j], x[j], yi, A[j]);
            c.madc_hi_cc(A[j + 1], x[j], yi, A[j + 1]);
        }
        c.addc(B[NL - 1], B[NL - 1], 0);
    }
    const uint32_t m = mul_lo(A[0], F.qinv);
    if (QS >= 2) {
        if (NL > 2) {
            c.mad_lo_cc(B[2], m, F.q[3], B[2]);
            c.madc_hi_cc(B[3], m, F.q[3], B[3]);
#pragma unroll
            for (int j = 5; j < NL; j += 2) {
                c.madc_lo_cc(B[j - 1], m, F.q[j], B[j - 1]);
                c.madc_hi_cc(B[j], m, F.q[j], B[j]);
            }
        }
    } else {
        c.mad_lo_cc(B[0], m, F.q[1], B[0]);
        c.madc_hi_cc(B[1], m, F.q[1], B[1]);
#pragma unroll
        for (int j = 3; j < NL; j += 2) {
            c.madc_lo_cc(B[j - 1], m, F.q[j], B[j - 1]);
            c.madc_hi_cc(B[j], m, F.q[j], B[j]);
        }
    }
    if (QS >= 1) {
        c.add_cc(A[0], A[0], m);
        c.addc_cc(A[1], A[1], 0);
    } else {
        c.mad_lo_cc(A[0], m, F.q[0], A[0]);
        c.madc_hi_cc(A[1], m, F.q[0], A[1]);
    }
#pragma unroll
    for (int j = 2; j < NL; j += 2) {
        c.madc_lo_cc(A[j], m, F.q[j], A[j]);
        c.madc_hi_cc(A[j + 1], m, F.q[j], A[j + 1]);
    }
    c.addc(B[NL - 1], B[NL - 1], 0);
}

// z = x * y * R^-1 (mod q), x in [0, 2q), y in [0, q)  ->  z in [0, 2q); no final subtraction
template <int NL, int QS>
RS_HD void mont_mul_lazy(uint32_t *z, const uint32_t *x, const uint32_t *y, const FieldParams<NL> &F) {
    static_assert(NL % 2 == 0 && NL >= 2, "NL must be even");
    uint32_t P[NL], Q[NL];
#pragma unroll
    for (int i = 0; i < NL; i += 2) {
        mont_row_nt<NL, QS>(P, Q, x, y[i], F, i == 0);
        mont_row_nt<NL, QS>(Q, P, x, y[i + 1], F, false);
    }
    CC c;
    c.add_cc(z[0], P[0], Q[1]);
#pragma unroll
    for (int k = 1; k < NL - 1; k++) c.addc_cc(z[k], P[k], Q[k + 1]);
    c.addc(z[NL - 1], P[NL - 1], 0);
}

}  // namespace rs
