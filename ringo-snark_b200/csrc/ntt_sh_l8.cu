// ntt_sh_l8.cu -- NTT kernels for 8 32-bit limbs with Shoup twiddle products (shoup.cuh): forward kernels for the modulus
// shapes QS in {1 2} (every jindo modulus is q == 1 mod 2^32; the generic shape keeps the Montgomery kernels of ntt_l8.cu),
// forward and inverse kernels with q compiled in for the 255-bit modulus (ntt.cuh FixedField).
#include "ntt_launch.cuh"

namespace rs {
static bool fixed_zp255(const rs_field *f) {
    static const bool off = getenv("RS_NTT_NO_FIXED_Q") != nullptr;
    return !off && f->NL == 8 && f->qs == 2 && fixed_field_matches<8, FID_ZP255>(f->q);
}
bool shoup_inverse_available(const rs_field *f) { return fixed_zp255(f); }

int ntt_run_sh_l8(const rs_ntt *t, uint32_t *out, const uint32_t *in, size_t n_polys, bool inverse) {
    if (fixed_zp255(t->f)) return ntt_run<8, 2, true, FID_ZP255>(t, out, in, n_polys, inverse);
    switch (t->f->qs) {
        case 1: return ntt_run<8, 1, true>(t, out, in, n_polys, inverse);
        case 2: return ntt_run<8, 2, true>(t, out, in, n_polys, inverse);
        default: return set_error(RS_ERR_UNSUPPORTED, "ntt: no Shoup kernels for this modulus shape");
    }
}
}  // namespace rs
