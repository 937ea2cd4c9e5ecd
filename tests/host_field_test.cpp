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
