// ntt52_v2.cu -- FP64-pipe NTT kernels, tuning variant 2: radix-2^2 rounds, 256 threads, 2 / 2 CTAs per SM (forward / inverse).
#include "ntt52_launch.cuh"

namespace rs {
int ntt52_run_v2(const rs_ntt *t, uint32_t *out, const uint32_t *in, size_t n_polys, bool inverse) {
    return ntt52_run_variant<Ntt52Variant<2, 8, 2, 2>>(t, out, in, n_polys, inverse);
}
}  // namespace rs
