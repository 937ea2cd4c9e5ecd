// Host build of ringo-snark_b200/csrc/shoup.cuh (carry flag emulated): the exact multiply-add sequence of the
// Shoup butterfly product and the predicated lazy add / sub, checked against Python integers without a GPU.
// Test infrastructure only.
#include <cstddef>
#include <cstring>
#include "../ringo-snark_b200/csrc/shoup.cuh"

using namespace rs;

// op 0: z = shoup_mul_lazy(x, w, wp)    (y holds w then wp, 2 NL words per item)
// op 1: z = lazy_add_p(x, y)   op 3: z = lazy_sub_p(x, y)
// op 4: frac / h of tri_hi (z gets NL words of h; frac returned through fr)
template <int NL, int QS>
static void run(int op, const uint32_t *q, const uint32_t *q2, const uint32_t *x, const uint32_t *y, uint32_t *z, uint32_t *fr, size_t n) {
    FieldParams<NL> F;
    memcpy(F.q, q, 4 * NL); memcpy(F.q2, q2, 4 * NL);
    F.qinv = 0;
    for (size_t i = 0; i < n; i++) {
        const uint32_t *a = x + i * NL;
        uint32_t *o = z + i * NL;
        switch (op) {
            case 0: shoup_mul_lazy<NL, QS>(o, a, y + i * 2 * NL, y + i * 2 * NL + NL, F); break;
            case 1: lazy_add_p<NL>(o, a, y + i * NL, F.q2); break;
            case 3: lazy_sub_p<NL>(o, a, y + i * NL, F.q2); break;
            case 4: tri_hi<NL>(o, fr[i], a, y + i * 2 * NL + NL); break;
        }
    }
}

extern "C" int hs_op(int NL, int qs, int op, const uint32_t *q, const uint32_t *q2, const uint32_t *x, const uint32_t *y, uint32_t *z,
                     uint32_t *fr, size_t n) {
#define SH(N, Q) if (NL == N && qs == Q) { run<N, Q>(op, q, q2, x, y, z, fr, n); return 0; }
    SH(4, 0) SH(4, 1) SH(8, 0) SH(8, 1) SH(8, 2)
#undef SH
    return -1;
}
