// ntt_sh_l8.cu -- forward NTT kernels for 8 32-bit limbs with Shoup twiddle products (shoup.cuh), modulus shapes QS in {1 2}
// (every jindo modulus is q == 1 mod 2^32; the generic shape keeps the Montgomery kernels of ntt_l8.cu).
#include "ntt_launch.cuh"

namespace rs {
int ntt_run_sh_l8(const rs_ntt *t, uint32_t *out, const uint32_t *in, size_t n_polys) {
    // the 255-bit jindo modulus has its own kernels with q compiled in (ntt.cuh FixedField)
    if (t->f->qs == 2 && fixed_field_matches<8, FID_ZP255>(t->f->q) && !getenv("RS_NTT_NO_FIXED_Q"))
        return ntt_run<8, 2, true, FID_ZP255>(t, out, in, n_polys, false);
    switch (t->f->qs) {
        case 1: return ntt_run<8, 1, true>(t, out, in, n_polys, false);
        case 2: return ntt_run<8, 2, true>(t, out, in, n_polys, false);
        default: return set_error(RS_ERR_UNSUPPORTED, "ntt: no Shoup kernels for this modulus shape");
    }
}
}  // namespace rs
