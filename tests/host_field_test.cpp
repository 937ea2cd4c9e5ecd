// Host build of ringo-snark_b200/csrc/field.cuh (carry flag emulated) so the exact
// device instruction sequence can be checked against Python ints without a GPU.
// Test infrastructure only.
#include <cstddef>
#include <cstring>
#include "../ringo-snark_b200/csrc/field.cuh"

using namespace rs;

template <int NL, int QS>
static void run(int op, const uint32_t *q, uint32_t qinv, const uint32_t *x, const uint32_t *y, uint32_t *z, size_t n) {
    FieldParams<NL> F;
    memcpy(F.q, q, 4 * NL);
    F.qinv = qinv;
    for (size_t i = 0; i < n; i++) {
        const uint32_t *a = x + i * NL, *b = y + i * NL;
        uint32_t *o = z + i * NL;
        switch (op) {
            case 0: mont_mul<NL, QS>(o, a, b, F); break;
            case 1: mod_add<NL>(o, a, b, F); break;
            case 2: mod_sub<NL>(o, a, b, F); break;
            case 3: mod_neg<NL>(o, a, F); break;
        }
    }
}

template <int NL>
static int dispatch_qs(int qs, int op, const uint32_t *q, uint32_t qinv, const uint32_t *x, const uint32_t *y, uint32_t *z, size_t n) {
    switch (qs) {
        case 0: run<NL, 0>(op, q, qinv, x, y, z, n); return 0;
        case 1: run<NL, 1>(op, q, qinv, x, y, z, n); return 0;
        case 2: run<NL, 2>(op, q, qinv, x, y, z, n); return 0;
    }
    return -1;
}

extern "C" int hf_op(int NL, int qs, int op, const uint32_t *q, uint32_t qinv, const uint32_t *x, const uint32_t *y, uint32_t *z, size_t n) {
    switch (NL) {
        case 2: return dispatch_qs<2>(qs, op, q, qinv, x, y, z, n);
        case 4: return dispatch_qs<4>(qs, op, q, qinv, x, y, z, n);
        case 8: return dispatch_qs<8>(qs, op, q, qinv, x, y, z, n);
        case 14: return dispatch_qs<14>(qs, op, q, qinv, x, y, z, n);
        case 28: return dispatch_qs<28>(qs, op, q, qinv, x, y, z, n);
    }
    return -2;
}

// ---- reduced-radix lazy arithmetic (rr_field.cuh) -------------------------------------------
#include "../ringo-snark_b200/csrc/rr_field.cuh"

template <int NL, int W, int QZ>
static void rr_run(int op, const uint32_t *params /* q, c3q, q2, one: 4*W */, const uint32_t *x, const uint32_t *y,
                   uint32_t *z, size_t n) {
    RRParams<W> P;
    memcpy(P.q, params, 4 * W); memcpy(P.c3q, params + W, 4 * W); memcpy(P.q2, params + 2 * W, 4 * W); memcpy(P.one, params + 3 * W, 4 * W);
    for (size_t i = 0; i < n; i++) {
        switch (op) {
            case 0: rr_mul<W, QZ>(z + i * W, x + i * W, y + i * W, P); break;          // RR limbs in, RR limbs out
            case 1: rr_from_words<NL, W>(z + i * W, x + i * NL); break;                // words -> RR
            case 2: rr_to_words<NL, W>(z + i * NL, x + i * W); break;                  // RR -> words
            case 3: {                                                                  // canonicalise a lazy value -> words
                uint32_t t[W];
                memcpy(t, x + i * W, 4 * W);
                rr_canonical<W, QZ>(t, P);
                rr_to_words<NL, W>(z + i * NL, t);
            } break;
            case 4: {   // three lazy butterfly stages on (u, v) with twiddles y (3 per item), then canonical:
                        // out = words(u3), words(v3)
                uint32_t u[W], v[W], t[W], s[W];
                rr_from_words<NL, W>(u, x + i * 2 * NL);
                rr_from_words<NL, W>(v, x + i * 2 * NL + NL);
                for (int st = 0; st < 3; st++) {
                    rr_mul<W, QZ>(t, v, y + (i * 3 + st) * W, P);
                    rr_sub<W>(s, u, t, P);
                    rr_add<W>(u, u, t);
                    // feed the difference back as the next multiplicand (worst case for limb growth)
                    memcpy(v, s, 4 * W);
                }
                rr_canonical<W, QZ>(u, P);
                rr_canonical<W, QZ>(v, P);
                rr_to_words<NL, W>(z + i * 2 * NL, u);
                rr_to_words<NL, W>(z + i * 2 * NL + NL, v);
            } break;
        }
    }
}

extern "C" int hf_rr(int NL, int W, int QZ, int op, const uint32_t *params, const uint32_t *x, const uint32_t *y, uint32_t *z, size_t n) {
    if (NL == 8 && W == 9 && QZ == 1) { rr_run<8, 9, 1>(op, params, x, y, z, n); return 0; }
    if (NL == 8 && W == 9 && QZ == 0) { rr_run<8, 9, 0>(op, params, x, y, z, n); return 0; }
    if (NL == 8 && W == 8 && QZ == 1) { rr_run<8, 8, 1>(op, params, x, y, z, n); return 0; }
    if (NL == 8 && W == 8 && QZ == 0) { rr_run<8, 8, 0>(op, params, x, y, z, n); return 0; }
    return -1;
}

// ---- lazy [0, 2q) variants of field.cuh -----------------------------------------------------
template <int NL, int QS>
static void lazy_run(int op, const uint32_t *q, uint32_t qinv, const uint32_t *q2, const uint32_t *x, const uint32_t *y, uint32_t *z, size_t n) {
    FieldParams<NL> F;
    memcpy(F.q, q, 4 * NL); memcpy(F.q2, q2, 4 * NL);
    F.qinv = qinv;
    for (size_t i = 0; i < n; i++) {
        const uint32_t *a = x + i * NL, *b = y + i * NL;
        uint32_t *o = z + i * NL;
        switch (op) {
            case 0: mont_mul_lazy<NL, QS>(o, a, b, F); break;
            case 1: mod_add_m<NL>(o, a, b, F.q2); break;
            case 2: mod_sub_m<NL>(o, a, b, F.q2); break;
            case 3: memcpy(o, a, 4 * NL); cond_sub<NL>(o, F.q); break;
        }
    }
}
extern "C" int hf_lazy(int NL, int qs, int op, const uint32_t *q, uint32_t qinv, const uint32_t *q2, const uint32_t *x, const uint32_t *y, uint32_t *z, size_t n) {
#define LZ(N, Q) if (NL == N && qs == Q) { lazy_run<N, Q>(op, q, qinv, q2, x, y, z, n); return 0; }
    LZ(4, 0) LZ(4, 1) LZ(8, 0) LZ(8, 1) LZ(8, 2)
#undef LZ
    return -1;
}
