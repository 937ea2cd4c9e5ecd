// ntt_l14.cu -- NTT kernels for 14 32-bit limbs (7 x uint64), modulus shapes QS in {0 2}.
#include "ntt_launch.cuh"

namespace rs {
int ntt_run_l14(const rs_ntt *t, uint32_t *out, const uint32_t *in, size_t n_polys, bool inverse) {
    switch (t->f->qs) {
        case 2: return ntt_run<14, 2>(t, out, in, n_polys, inverse);
        default: return ntt_run<14, 0>(t, out, in, n_polys, inverse);
    }
}
}  // namespace rs
