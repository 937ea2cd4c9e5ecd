// ntt_l8.cu -- NTT kernels for 8 32-bit limbs (4 x uint64), modulus shapes QS in {0 1 2}.
#include "ntt_launch.cuh"

namespace rs {
int ntt_run_l8(const rs_ntt *t, uint32_t *out, const uint32_t *in, size_t n_polys, bool inverse) {
    switch (t->f->qs) {
        case 1: return ntt_run<8, 1>(t, out, in, n_polys, inverse);
        case 2: return ntt_run<8, 2>(t, out, in, n_polys, inverse);
        default: return ntt_run<8, 0>(t, out, in, n_polys, inverse);
    }
}
}  // namespace rs
