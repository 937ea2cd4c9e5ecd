// ntt_l4.cu -- NTT kernels for 4 32-bit limbs (2 x uint64), modulus shapes QS in {0 1}.
#include "ntt_launch.cuh"

namespace rs {
int ntt_run_l4(const rs_ntt *t, uint32_t *out, const uint32_t *in, size_t n_polys, bool inverse) {
    switch (t->f->qs) {
        case 1: return ntt_run<4, 1>(t, out, in, n_polys, inverse);
        default: return ntt_run<4, 0>(t, out, in, n_polys, inverse);
    }
}
}  // namespace rs
